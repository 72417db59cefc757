"""Device-resident reduction results (include/nomp-b200.h: nomp_b200_device_reductions) and kernels that read their
scalars from device memory: a conjugate-gradient iteration enqueued without a single host round trip.

The CPU tier runs this file on the CUDA test double as well (tests/test_hostdev_cpu.py).
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "libnomp_b200" / "python"))

from libnomp_b200 import capi  # noqa: E402
from nomp_bridge.families import AX_DOT_KERNEL_SOURCE  # noqa: E402
from oracle import ffi  # noqa: E402

pytestmark = [pytest.mark.gpu]

P, I, F = capi.NOMP_PTR, capi.NOMP_INT, capi.NOMP_FLOAT


@pytest.fixture(scope="module", autouse=True)
def runtime():
    capi.check(capi.init(backend="cuda", device=0, verbose=0))
    yield capi.nomp()
    assert capi.nomp().nomp_finalize_excluding_interpreter() == 0


def jit(src, clauses, args):
    err, kid = capi.jit(src, clauses, args)
    capi.check(err)
    return kid


def to_device(*arrays):
    for a in arrays:
        capi.check(capi.update(a.ctypes.data, 0, a.size, a.itemsize, capi.NOMP_TO))


def from_device(*arrays):
    for a in arrays:
        capi.check(capi.update(a.ctypes.data, 0, a.size, a.itemsize, capi.NOMP_FROM))


def free(*arrays):
    for a in arrays:
        capi.check(capi.update(a.ctypes.data, 0, a.size, a.itemsize, capi.NOMP_FREE))


def test_mapped_reduction_variable_keeps_its_result_on_the_device():
    lib = capi.nomp()
    n = 10007
    a = ffi.fill_int_f64(n, 3, 0, 7)
    s, t = np.array([-1.0]), np.array([-2.0])
    to_device(a, s)
    kid = jit("void sum(const double *a, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i]; }",
              capi.clauses(("reduce", "s", "+")), [("a", 8, P), ("N", 4, I), ("s", 8, F)])
    # default: the reference's semantics even for a mapped variable -- the host copy holds the sum when nomp_run returns
    assert lib.nomp_b200_device_reductions(-1) == 0
    capi.check(capi.run(kid, a.ctypes.data, C.c_int(n), s.ctypes.data))
    assert s[0] == a.sum()
    # on: the host copy is left alone, the device copy holds the result
    assert lib.nomp_b200_device_reductions(1) == 0 and lib.nomp_b200_device_reductions(-1) == 1
    s[0] = -1.0
    capi.check(capi.run(kid, a.ctypes.data, C.c_int(n), s.ctypes.data))
    assert s[0] == -1.0
    from_device(s)
    assert s[0] == a.sum()
    # a variable that is not mapped still gets its value on the host at once, and later waits are not confused by the
    # launches that did not wait (sequence numbers)
    capi.check(capi.run(kid, a.ctypes.data, C.c_int(n), t.ctypes.data))
    assert t[0] == a.sum()
    # integer type, generated reduction (condition -> skeleton), 4-byte result next to other data
    b = (np.arange(n) % 11).astype(np.int32)
    cnt = np.array([7, -5, 9], dtype=np.int32)
    to_device(b, cnt)
    kid2 = jit("void count(const int *b, int N, int *c) { for (int i = 0; i < N; i++) if (b[i] > 4) c[0] += b[i]; }",
               capi.clauses(("reduce", "c", "+")), [("b", 4, P), ("N", 4, I), ("c", 4, I)])
    capi.check(capi.run(kid2, b.ctypes.data, C.c_int(n), cnt.ctypes.data))
    assert list(cnt) == [7, -5, 9]
    from_device(cnt)
    assert list(cnt) == [int(b[b > 4].sum()), -5, 9]
    assert lib.nomp_b200_device_reductions(0) == 1
    free(a, s, b, cnt)


def test_cg_iterations_without_host_round_trips():
    """CG on the local Poisson operator with every scalar in device memory: per iteration Ax + p.Ap (native), alpha,
    the fused update + r.r (reduce skeleton reading alpha[0]), beta, the new direction (map skeleton reading beta[0]) --
    six launches, no host wait; the host looks at the residual after the last iteration.  Against the same CG on the
    host (oracle Ax, numpy dots)."""
    lib = capi.nomp()
    n, E, iters = 8, 6, 5
    n3 = n ** 3
    N = E * n3
    xt = ffi.fill_uniform_f64(N, 11, 0.0, 1.0) - 0.5
    v = ffi.fill_uniform_f64(6 * N, 13, 0.0, 1.0).reshape(E, 6, n3)
    g = 0.2 * (v - 0.5)
    for f in (0, 3, 5):
        g[:, f, :] = 1.0 + 0.5 * v[:, f, :]
    g = np.ascontiguousarray(g.ravel())
    D = np.ascontiguousarray(ffi.gll_derivative(n)[0].ravel())
    b = ffi.ax(n, xt, g, D)
    # host CG
    x, r, p = np.zeros(N), b.copy(), b.copy()
    rr = float(r @ r)
    hist = []
    for _ in range(iters):
        w = ffi.ax(n, p, g, D)
        pap = float(p @ w)
        alpha = rr / pap
        x += alpha * p
        r -= alpha * w
        rr_new = float(r @ r)
        p = r + (rr_new / rr) * p
        rr = rr_new
        hist.append((pap, alpha, rr))
    # device CG
    dx, dr, dp, dw = np.zeros(N), b.copy(), b.copy(), np.zeros(N)
    sc = {k: np.zeros(1) for k in ("pap", "alpha", "beta", "rr", "rr_new")}
    sc["rr"][0] = float(b @ b)
    trace = np.zeros(3 * iters)
    to_device(dx, dr, dp, dw, g, D, trace, *sc.values())
    k_ax = jit(AX_DOT_KERNEL_SOURCE, capi.clauses(("reduce", "pap", "+")),
               [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I), ("n", 4, I | capi.NOMP_JIT, C.c_int(n)), ("pap", 8, F)])
    k_alpha = jit("void cg_alpha(double *alpha, const double *rr, const double *pap, double *trace, int it) {"
                  " for (int i = 0; i < 1; i++) { alpha[i] = rr[i] / pap[i]; trace[3 * it] = pap[i]; trace[3 * it + 1] = alpha[i]; } }",
                  capi.clauses(), [("alpha", 8, P), ("rr", 8, P), ("pap", 8, P), ("trace", 8, P), ("it", 4, I)])
    k_upd = jit("void cg_update(double *x, double *r, const double *p, const double *w, const double *alpha, int N, double *rr_new) {"
                " for (int i = 0; i < N; i++) { x[i] += alpha[0] * p[i]; r[i] -= alpha[0] * w[i]; rr_new[0] += r[i] * r[i]; } }",
                capi.clauses(("reduce", "rr_new", "+")),
                [("x", 8, P), ("r", 8, P), ("p", 8, P), ("w", 8, P), ("alpha", 8, P), ("N", 4, I), ("rr_new", 8, F)])
    k_beta = jit("void cg_beta(double *beta, double *rr, const double *rr_new, double *trace, int it) {"
                 " for (int i = 0; i < 1; i++) { beta[i] = rr_new[i] / rr[i]; rr[i] = rr_new[i]; trace[3 * it + 2] = rr_new[i]; } }",
                 capi.clauses(), [("beta", 8, P), ("rr", 8, P), ("rr_new", 8, P), ("trace", 8, P), ("it", 4, I)])
    k_dir = jit("void cg_direction(double *p, const double *r, const double *beta, int N) {"
                " for (int i = 0; i < N; i++) p[i] = r[i] + beta[0] * p[i]; }",
                capi.clauses(), [("p", 8, P), ("r", 8, P), ("beta", 8, P), ("N", 4, I)])
    info = lambda k: lib.nomp_b200_prog_info(k).decode()  # noqa: E731
    assert "family=axdot" in info(k_ax) and "family=reduce" in info(k_upd) and "family=map" in info(k_dir)
    lib.nomp_b200_device_reductions(1)
    launches = lib.nomp_b200_launch_count()
    ptr = lambda a: a.ctypes.data  # noqa: E731
    for it in range(iters):
        capi.check(capi.run(k_ax, ptr(dw), ptr(dp), ptr(g), ptr(D), C.c_int(E), ptr(sc["pap"])))
        capi.check(capi.run(k_alpha, ptr(sc["alpha"]), ptr(sc["rr"]), ptr(sc["pap"]), ptr(trace), C.c_int(it)))
        capi.check(capi.run(k_upd, ptr(dx), ptr(dr), ptr(dp), ptr(dw), ptr(sc["alpha"]), C.c_int(N), ptr(sc["rr_new"])))
        capi.check(capi.run(k_beta, ptr(sc["beta"]), ptr(sc["rr"]), ptr(sc["rr_new"]), ptr(trace), C.c_int(it)))
        capi.check(capi.run(k_dir, ptr(dp), ptr(dr), ptr(sc["beta"]), C.c_int(N)))
    assert lib.nomp_b200_launch_count() - launches == 5 * iters
    assert sc["pap"][0] == 0.0 and sc["rr_new"][0] == 0.0            # nothing came back to the host on its own
    lib.nomp_b200_device_reductions(0)
    from_device(dx, trace, sc["rr"])
    for it, (pap, alpha, rr_it) in enumerate(hist):
        for got, want in zip(trace[3 * it: 3 * it + 3], (pap, alpha, rr_it)):
            assert abs(got - want) <= 1e-10 * abs(want), (it, got, want)
    assert sc["rr"][0] == trace[-1]
    assert np.max(np.abs(dx - x)) <= 1e-9 * np.max(np.abs(x))
    free(dx, dr, dp, dw, g, D, trace, *sc.values())


def test_cg_example_gives_the_same_iterates_with_device_scalars():
    """examples/cg_poisson.c, "host" against "device": the same kernels on the same data in the same order -- the
    iterates agree to the last bit, and so do the iteration counts when the check interval divides them."""
    import json
    import subprocess
    exe = ROOT / "libnomp_b200" / "build" / "cg_poisson"
    if not exe.exists():
        pytest.skip("examples/cg_poisson was not built")
    env = dict(os.environ, NOMP_INSTALL_DIR=str(ROOT / "libnomp_b200"))
    runs = {}
    for mode in ("host", "device", "device3", "fused", "device_fused"):
        r = subprocess.run([str(exe), "6", "8", "400", "1e-9", mode, "1", "--nomp-backend", "cuda", "--nomp-device", "0",
                            "--nomp-verbose", "1"], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        runs[mode] = [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
    host, dev = runs["host"], runs["device"]
    assert host[0] == dev[0] and host[1:6] == dev[1:6]
    assert dev[-1]["scalars"] == "device" and host[-1]["scalars"] == "host"
    assert dev[-1]["iterations"] == host[-1]["iterations"] and dev[-1]["rr_final"] == host[-1]["rr_final"]
    assert dev[-1]["true_residual_rel"] < 1e-7
    # three launches per iteration (alpha and beta folded into their consumers, the residual scalars swapped by name)
    fused = runs["fused"]          # the direction update inside the operator kernel: two launches per iteration
    assert fused[:6] == host[:6] and fused[-1]["scalars"] == "fused" and fused[-1]["bytes_per_dof"] == 128
    assert fused[-1]["iterations"] == host[-1]["iterations"] and fused[-1]["rr_final"] == host[-1]["rr_final"]
    both = runs["device_fused"]    # ... and with the scalars in device memory: three launches, 128 B/DOF, no round trip
    assert both[0] == host[0] and both[-1]["scalars"] == "device_fused" and both[-1]["bytes_per_dof"] == 128
    assert both[-1]["iterations"] == host[-1]["iterations"] and both[-1]["rr_final"] == host[-1]["rr_final"]
    dev3 = runs["device3"]
    assert dev3[-1]["scalars"] == "device3" and dev3[0] == host[0]
    assert dev3[-1]["iterations"] == host[-1]["iterations"] and dev3[-1]["rr_final"] == host[-1]["rr_final"]


def test_cg_iterations_replayed_from_a_cuda_graph():
    """nomp_b200_graph_*: two CG iterations in the three-launch form (the residual scalars swap names every iteration,
    so two iterations make one period) are captured once and replayed; the iterates equal those of the same launches
    issued one by one, bit for bit.  What a capturing stream cannot do is refused."""
    lib = capi.nomp()
    n, E, periods = 8, 5, 3
    N = E * n ** 3
    xt = ffi.fill_uniform_f64(N, 21, 0.0, 1.0) - 0.5
    v = ffi.fill_uniform_f64(6 * N, 23, 0.0, 1.0).reshape(E, 6, n ** 3)
    g = 0.2 * (v - 0.5)
    for f in (0, 3, 5):
        g[:, f, :] = 1.0 + 0.5 * v[:, f, :]
    g = np.ascontiguousarray(g.ravel())
    D = np.ascontiguousarray(ffi.gll_derivative(n)[0].ravel())
    b = ffi.ax(n, xt, g, D)
    args_ax = [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I), ("n", 4, I | capi.NOMP_JIT, C.c_int(n)), ("pap", 8, F)]
    k_ax = jit(AX_DOT_KERNEL_SOURCE, capi.clauses(("reduce", "pap", "+")), args_ax)
    k_upd = jit("void cg_update_3(double *x, double *r, const double *p, const double *w, const double *rr, const double *pap, int N,"
                " double *rr_new) { for (int i = 0; i < N; i++) { x[i] += (rr[0] / pap[0]) * p[i]; r[i] -= (rr[0] / pap[0]) * w[i];"
                " rr_new[0] += r[i] * r[i]; } }", capi.clauses(("reduce", "rr_new", "+")),
                [("x", 8, P), ("r", 8, P), ("p", 8, P), ("w", 8, P), ("rr", 8, P), ("pap", 8, P), ("N", 4, I), ("rr_new", 8, F)])
    k_dir = jit("void cg_direction_3(double *p, const double *r, const double *rr_new, const double *rr, int N) {"
                " for (int i = 0; i < N; i++) p[i] = r[i] + (rr_new[0] / rr[0]) * p[i]; }", capi.clauses(),
                [("p", 8, P), ("r", 8, P), ("rr_new", 8, P), ("rr", 8, P), ("N", 4, I)])
    ptr = lambda a: a.ctypes.data  # noqa: E731

    def solve(use_graph):
        x, r, p, w = np.zeros(N), b.copy(), b.copy(), np.zeros(N)
        pap, rr, rrn = np.zeros(1), np.array([float(b @ b)]), np.zeros(1)
        state = (x, r, p, w, pap, rr, rrn)
        to_device(*state)

        def iteration(cur, new):
            capi.check(capi.run(k_ax, ptr(w), ptr(p), ptr(g), ptr(D), C.c_int(E), ptr(pap)))
            capi.check(capi.run(k_upd, ptr(x), ptr(r), ptr(p), ptr(w), ptr(cur), ptr(pap), C.c_int(N), ptr(new)))
            capi.check(capi.run(k_dir, ptr(p), ptr(r), ptr(new), ptr(cur), C.c_int(N)))

        def period():
            iteration(rr, rrn)
            iteration(rrn, rr)

        lib.nomp_b200_device_reductions(1)
        period()                                   # also loads every kernel before anything is captured
        if use_graph:
            graph = C.c_int(-1)
            capi.check(lib.nomp_b200_graph_begin())
            period()
            # a capturing stream cannot be waited for, and a reduction cannot deliver to the host
            assert capi.err_info(lib.nomp_sync())[0] == capi.NOMP_USER_INPUT_IS_INVALID
            assert capi.err_info(capi.update(ptr(x), 0, N, 8, capi.NOMP_FROM))[0] == capi.NOMP_USER_INPUT_IS_INVALID
            host_scalar = C.c_double(0.0)
            assert capi.err_info(capi.run(k_ax, ptr(w), ptr(p), ptr(g), ptr(D), C.c_int(E), host_scalar))[0] == capi.NOMP_USER_INPUT_IS_INVALID
            capi.check(lib.nomp_b200_graph_end(C.byref(graph)))
            assert graph.value >= 0
            for _ in range(periods):               # the capture itself executed nothing: `periods` replays do the work
                capi.check(lib.nomp_b200_graph_launch(graph.value))
            capi.check(lib.nomp_sync())
            capi.check(lib.nomp_b200_graph_free(graph.value))
            assert capi.err_info(lib.nomp_b200_graph_launch(graph.value))[0] == capi.NOMP_USER_INPUT_IS_INVALID
        else:
            for _ in range(periods):
                period()
        lib.nomp_b200_device_reductions(0)
        from_device(x, rr)
        out = x.copy(), rr[0]
        free(*state)
        return out

    to_device(g, D)
    x_plain, rr_plain = solve(False)
    x_graph, rr_graph = solve(True)
    free(g, D)
    assert rr_graph == rr_plain and np.array_equal(x_graph, x_plain)
    assert rr_plain < 1e-2 * float(b @ b)            # and the iteration does reduce the residual


def test_staged_derivative_matrix_survives_graph_replays_and_remapping():
    """The Ax family stages D into __constant__ memory and skips the copy while the same device image is reused.  A
    replayed graph re-stages ITS matrix behind the runtime's back, and a mapping freed and created again can come back
    at the same device address: in both cases the next launch must stage its own D again."""
    lib = capi.nomp()
    n, E = 8, 9
    u = ffi.fill_int_f64(E * n ** 3, 5, -4, 4)
    g = ffi.fill_int_f64(E * 6 * n ** 3, 6, 0, 3)
    D1, D2 = ffi.fill_int_f64(n * n, 7, -2, 2), ffi.fill_int_f64(n * n, 8, -2, 2)
    w, pap = np.zeros_like(u), np.zeros(1)
    from nomp_bridge.families import AX_KERNEL_SOURCE
    k_ax = jit(AX_KERNEL_SOURCE, capi.clauses(), [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I),
                                                  ("n", 4, I | capi.NOMP_JIT, C.c_int(n))])
    to_device(u, g, D1, D2, w, pap)
    ptr = lambda a: a.ctypes.data  # noqa: E731
    ax = lambda D: capi.check(capi.run(k_ax, ptr(w), ptr(u), ptr(g), ptr(D), C.c_int(E)))  # noqa: E731
    ax(D1)                                             # loads the kernel
    graph = C.c_int(-1)
    capi.check(lib.nomp_b200_graph_begin())
    ax(D1)
    capi.check(lib.nomp_b200_graph_end(C.byref(graph)))
    ax(D2)                                             # caches D2 ...
    capi.check(lib.nomp_b200_graph_launch(graph.value))  # ... the replay stages D1 ...
    from_device(w)
    assert np.array_equal(w, ffi.ax(n, u, g, D1))
    ax(D2)                                             # ... so this launch must not trust its cache
    from_device(w)
    assert np.array_equal(w, ffi.ax(n, u, g, D2))
    capi.check(lib.nomp_b200_graph_free(graph.value))
    # free D2 and map a matrix with other values at the same host address (and, usually, the same device address)
    free(D2)
    D2[:] = ffi.fill_int_f64(n * n, 9, -2, 2)
    to_device(D2)
    ax(D2)
    from_device(w)
    assert np.array_equal(w, ffi.ax(n, u, g, D2))
    free(u, g, D1, D2, w, pap)


@pytest.mark.parametrize("n", [6, 8, 10, 12])
def test_direction_update_fused_into_the_operator(n):
    """The canonical xpay + Ax + dot kernel strings (beta as a scalar argument, and as beta[0] in device memory) -> one
    launch of nompk_ax_xpay_dot_peers_f64, against the two-kernel sequence p = r + beta p; w = A p, p.w through the same
    API: bitwise on exact data, p updated in place."""
    from nomp_bridge.families import AX_XPAY_DOT_DEV_KERNEL_SOURCE, AX_XPAY_DOT_KERNEL_SOURCE
    lib = capi.nomp()
    E = 13
    p0 = ffi.fill_int_f64(E * n ** 3, 31, -2, 2)
    r = ffi.fill_int_f64(E * n ** 3, 32, -2, 2)
    g = ffi.fill_int_f64(E * 6 * n ** 3, 33, 0, 3)
    D = ffi.fill_int_f64(n * n, 34, -2, 2)
    beta = 3.0
    # the two-kernel sequence
    k_dir = jit("void dir(double *p, const double *r, double beta, int N) { for (int i = 0; i < N; i++) p[i] = r[i] + beta * p[i]; }",
                capi.clauses(), [("p", 8, P), ("r", 8, P), ("beta", 8, F), ("N", 4, I)])
    k_ax = jit(AX_DOT_KERNEL_SOURCE, capi.clauses(("reduce", "pap", "+")),
               [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I), ("n", 4, I | capi.NOMP_JIT, C.c_int(n)), ("pap", 8, F)])
    p, w = p0.copy(), np.zeros_like(p0)
    to_device(p, r, g, D, w)
    pap = C.c_double(0.0)
    capi.check(capi.run(k_dir, p.ctypes.data, r.ctypes.data, C.c_double(beta), C.c_int(p.size)))
    capi.check(capi.run(k_ax, w.ctypes.data, p.ctypes.data, g.ctypes.data, D.ctypes.data, C.c_int(E), pap))
    from_device(p, w)
    want_p, want_w, want_pap = p.copy(), w.copy(), pap.value
    assert np.array_equal(want_p, r + beta * p0) and want_pap == float(want_p @ want_w)
    # one launch, beta by value
    args = [("w", 8, P), ("p", 8, P), ("res", 8, P), ("g", 8, P), ("D", 8, P), ("beta", 8, F), ("E", 4, I),
            ("n", 4, I | capi.NOMP_JIT, C.c_int(n)), ("pap", 8, F)]
    k_fused = jit(AX_XPAY_DOT_KERNEL_SOURCE, capi.clauses(("reduce", "pap", "+")), args)
    assert "family=axxpaydot" in lib.nomp_b200_prog_info(k_fused).decode()
    p[:], w[:] = p0, np.nan
    to_device(p, w)
    launches = lib.nomp_b200_launch_count()
    pap = C.c_double(0.0)
    capi.check(capi.run(k_fused, w.ctypes.data, p.ctypes.data, r.ctypes.data, g.ctypes.data, D.ctypes.data, C.c_double(beta), C.c_int(E), pap))
    assert lib.nomp_b200_launch_count() - launches == 1
    from_device(p, w)
    assert np.array_equal(p, want_p) and np.array_equal(w, want_w) and pap.value == want_pap
    # one launch, beta in device memory
    args[5] = ("beta", 8, P)
    k_fused_dev = jit(AX_XPAY_DOT_DEV_KERNEL_SOURCE, capi.clauses(("reduce", "pap", "+")), args)
    beta_d = np.array([beta])
    p[:], w[:] = p0, np.nan
    to_device(p, w, beta_d)
    pap = C.c_double(0.0)
    capi.check(capi.run(k_fused_dev, w.ctypes.data, p.ctypes.data, r.ctypes.data, g.ctypes.data, D.ctypes.data, beta_d.ctypes.data, C.c_int(E), pap))
    from_device(p, w)
    assert np.array_equal(p, want_p) and np.array_equal(w, want_w) and pap.value == want_pap
    free(p, r, g, D, w, beta_d)
