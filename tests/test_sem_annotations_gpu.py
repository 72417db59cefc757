"""spectral-element kernels that are NOT one of the hand-written families get the
one-block-per-element schedule from the annotations script libnomp_b200/python/nomp_sem.py (element_loop / dof_loop
clauses), with per-element temporaries in shared memory.  Checked against the kernel string itself compiled by gcc."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "libnomp_b200" / "python"))

from libnomp_b200 import capi  # noqa: E402
from nomp_bridge.families import AX_KERNEL_SOURCE  # noqa: E402
from oracle import ffi  # noqa: E402
from tests.kernel_oracle import run_kernel  # noqa: E402

P, I = capi.NOMP_PTR, capi.NOMP_INT
POINT = "e * n * n * n + k * n * n + j * n + i"
HELMHOLTZ = (AX_KERNEL_SOURCE.replace("const double *D, int E, int n)", "const double *D, const double *h, int E, int n)")
             .replace("nomp_ax(", "helmholtz(").replace(f"w[{POINT}] = acc;", f"w[{POINT}] = acc + h[{POINT}] * u[{POINT}];"))
SEM_CLAUSES = (("annotate", "element_loop", "e"), ("annotate", "dof_loop", "i"), ("annotate", "dof_loop", "j"),
               ("annotate", "dof_loop", "k"))


@pytest.fixture(scope="module", autouse=True)
def runtime():
    capi.check(capi.init(backend="cuda", device=0, verbose=0, annotations_script="nomp_sem"))
    yield capi.nomp()
    assert capi.nomp().nomp_finalize_excluding_interpreter() == 0


def helmholtz_case(n, E, seed):
    n3 = n ** 3
    u = ffi.fill_uniform_f64(E * n3, seed, -1.0, 1.0)
    g = ffi.fill_uniform_f64(E * 6 * n3, seed + 1, 0.5, 1.5)
    h = ffi.fill_uniform_f64(E * n3, seed + 2, 0.0, 2.0)
    D = np.ascontiguousarray(ffi.gll_derivative(n)[0].ravel())
    return u, g, D, h


@pytest.mark.parametrize("n", [4, 6, 8, 10])
def test_annotated_operator_matches_its_own_c_loop(n):
    assert "acc + h[" in HELMHOLTZ
    E = 37
    u, g, D, h = helmholtz_case(n, E, 40 + n)
    want = np.zeros_like(u)
    run_kernel(HELMHOLTZ, want, u, g, D, h, E, n)
    w = np.zeros_like(u)
    arrays = (w, u, g, D, h)
    for a in arrays:
        capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_TO))
    err, kid = capi.jit(HELMHOLTZ, capi.clauses(*SEM_CLAUSES),
                        [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("h", 8, P), ("E", 4, I),
                         ("n", 4, I | capi.NOMP_JIT, C.c_int(n))])
    capi.check(err)
    info = capi.nomp().nomp_b200_prog_info(kid).decode()
    assert "kind=nvrtc" in info and "family=generic" in info
    capi.check(capi.run(kid, w.ctypes.data, u.ctypes.data, g.ctypes.data, D.ctypes.data, h.ctypes.data, C.c_int(E)))
    capi.check(capi.update(w.ctypes.data, 0, w.size, 8, capi.NOMP_FROM))
    for a in arrays:
        capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_FREE))
    # same operations in the same order, no FMA contraction on either side
    assert np.array_equal(w, want)


def test_annotated_operator_throughput():
    """Not a gate on speed of light -- this is the generic path -- but the schedule must be a real GPU schedule."""
    n, E = 8, 32768
    u, g, D, h = helmholtz_case(n, E, 77)
    w = np.zeros_like(u)
    arrays = (w, u, g, D, h)
    for a in arrays:
        capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_TO))
    err, kid = capi.jit(HELMHOLTZ, capi.clauses(*SEM_CLAUSES),
                        [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("h", 8, P), ("E", 4, I),
                         ("n", 4, I | capi.NOMP_JIT, C.c_int(n))])
    capi.check(err)
    args = (w.ctypes.data, u.ctypes.data, g.ctypes.data, D.ctypes.data, h.ctypes.data, C.c_int(E))
    for _ in range(5):
        capi.check(capi.run(kid, *args))
    capi.check(capi.nomp().nomp_sync())
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        capi.check(capi.run(kid, *args))
    capi.check(capi.nomp().nomp_sync())
    dt = (time.perf_counter() - t0) / reps
    gdofs = E * n ** 3 / dt / 1e9
    print(f"annotated Helmholtz operator, generic path: {dt * 1e3:.3f} ms, {gdofs:.1f} GDOF/s, {72 * gdofs:.0f} GB/s of 72 B/DOF")
    for a in arrays:
        capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_FREE))
    assert gdofs > 5.0
