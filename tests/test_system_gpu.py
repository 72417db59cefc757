"""System-level checks on the GPU box: the reference's own nomp-api programs (when they were built into oracle/_ref in
the build container), the smoke entry point, the bench.py contract, init/finalize cycles of a C-style host, and -- with
two or more GPUs -- the NCCL allreduce of the reduce clause."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = Path(__file__).resolve().parent.parent


def test_reference_nomp_api_suite_unmodified():
    tests = ROOT / "oracle" / "_ref" / "tests"
    if not (tests / "nomp-api-000").exists():
        pytest.skip("reference test programs were not built (needs /root/reference at build time)")
    r = subprocess.run([str(ROOT / "tools" / "run_ref_tests.sh")], capture_output=True, text=True, timeout=1800)
    tail = "\n".join(r.stdout.splitlines()[-60:])
    assert r.returncode == 0 and "reference suite failures: 0" in r.stdout, tail
    assert r.stdout.count(": Passed") >= 16, tail


def test_smoke_entry_point():
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "match the oracle bit for bit" in r.stdout, r.stdout + r.stderr


def test_bench_contract():
    r = subprocess.run([sys.executable, "bench.py", "--steps", "20", "--warmup", "3"], cwd=ROOT, capture_output=True, text=True,
                       timeout=1800)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["gpu_launches"] == 20 and line["dtype"] == "f64" and line["unit"] == "GDOF/s"
    assert line["roofline"]["bound"] == "hbm" and 0.3 < line["roofline"]["frac"] < 1.3
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["value"] < line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert "family=ax" in line["config"]["kernel"]


NCCL_WORKER = r"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, {root!r})
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(NOMP_COMM_SIZE=str(world), NOMP_COMM_RANK=str(rank), NOMP_COMM_ID_FILE=sys.argv[3], NOMP_COMM_ALLREDUCE=sys.argv[4])
from libnomp_b200 import capi
from oracle import ffi
capi.check(capi.init(backend="cuda", device=rank, verbose=1))
lib = capi.nomp()
assert lib.nomp_b200_comm_size() == world and lib.nomp_b200_comm_rank() == rank
if sys.argv[4] == "nccl":
    assert lib.nomp_b200_comm_uses_nvlink_kernel() == 0
P, I, F = capi.NOMP_PTR, capi.NOMP_INT, capi.NOMP_FLOAT
n = 1 << 20
x = ffi.fill_int_f64(n, 11, 0, 7); y = ffi.fill_int_f64(n, 12, 0, 7); xi = ffi.fill_i64(n, 5)
lo, hi = n * rank // world, n * (rank + 1) // world
xl, yl, xil = x[lo:hi].copy(), y[lo:hi].copy(), xi[lo:hi].copy()
for a in (xl, yl, xil):
    capi.check(capi.update(a.ctypes.data, 0, a.size, 8, capi.NOMP_TO))
red = capi.clauses(("reduce", "s", "+"))
err, kd = capi.jit("void f(const double *a, const double *b, int N, double *s) {{ for (int i = 0; i < N; i++) s[0] += a[i] * b[i]; }}", red, [("a", 8, P), ("b", 8, P), ("N", 4, I), ("s", 8, F)]); capi.check(err)
err, ki = capi.jit("void f(const long *a, int N, long *s) {{ for (int i = 0; i < N; i++) s[0] += a[i]; }}", red, [("a", 8, P), ("N", 4, I), ("s", 8, I)]); capi.check(err)
err, kc = capi.jit("void f(const double *a, int N, double *s) {{ for (int i = 0; i < N; i++) {{ if (a[i] > 3) s[0] += 1; }} }}", red, [("a", 8, P), ("N", 4, I), ("s", 8, F)]); capi.check(err)
err, km = capi.jit("void f(const double *a, int N, double *m) {{ for (int i = 0; i < N; i++) m[0] = (a[i] > m[0]) ? a[i] : m[0]; }}", capi.clauses(("reduce", "m", "max")), [("a", 8, P), ("N", 4, I), ("m", 8, F)]); capi.check(err)
s, si = C.c_double(), C.c_long()
for _ in range(3):
    capi.check(capi.run(kd, xl.ctypes.data, yl.ctypes.data, C.c_int(hi - lo), s))
    assert s.value == ffi.reduce_(0, ffi.F64, x, y), (s.value, "dot over all ranks")
capi.check(capi.run(ki, xil.ctypes.data, C.c_int(hi - lo), si)); assert si.value == ffi.reduce_(0, ffi.I64, xi)
capi.check(capi.run(kc, xl.ctypes.data, C.c_int(hi - lo), s)); assert s.value == float((x > 3).sum())
capi.check(capi.run(km, xl.ctypes.data, C.c_int(hi - lo), s)); assert s.value == x.max()
path = "nvlink-kernel" if lib.nomp_b200_comm_uses_nvlink_kernel() else "nccl"
assert lib.nomp_finalize_excluding_interpreter() == 0
print("rank", rank, "ok", path)
"""


@pytest.mark.parametrize("mode", ["auto", "nccl"])
def test_allreduce_of_reduce_clause_across_gpus(tmp_path, mode):
    """auto = the NVLink one-shot kernel when peers can be mapped (else NCCL); nccl = forced ncclAllReduce."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(NCCL_WORKER.format(root=str(ROOT)))
    idfile = f"/dev/shm/nomp-test-nccl-{os.getpid()}-{mode}"
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), idfile, mode], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for f in [idfile] + [f"{idfile}.ipc.{r}" for r in range(world)]:
        try:
            os.unlink(f)
        except OSError:
            pass
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in o, o[-3000:]
    print(outs[0].strip().splitlines()[-1])
