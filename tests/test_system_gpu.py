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
    assert line["ms_per_step_min"] <= line["ms_per_step_median"] and line["roofline"]["protocol_50_launches"]["reps"] == 50
    # the run checks itself: p.Ap against the oracle on a sampled slab, BASELINE configs[1] values against closed forms
    sc, ex = line["selfcheck"], line["extras"]
    assert sc["ok"] and sc["pAp_rel_diff_vs_oracle_sample"] <= 1e-12 and sc["pAp_rel_diff_fused_vs_unfused"] <= 1e-12, sc
    for key in ("sum_f64", "dot_f64", "sum_i64", "dot_i64"):
        assert ex[key]["checked"] and ex[key]["n_total"] == 1 << 28 and ex[key]["value"] == ex[key]["expected"], ex[key]
    assert ex["dot_f64_device_resident"]["checked"]
    for key in ("cg_step", "cg_step_fused", "cg_step_fused_graph"):     # device-resident scalars / graph replay: same p.Ap
        assert ex[key]["pAp_equals_host_semantics"], ex[key]
    n10 = ex["n10"]                                                    # BASELINE configs[3] at N = 9
    assert "family=ax n=10" in n10["ax"]["kernel"] and n10["ax"]["roofline"]["frac"] > 0.5 and n10["selfcheck"]["ok"], n10
    assert n10["cg_step_fused_graph"]["pAp_equals_host_semantics"]


def _cg_reference(E_total, n, iters):
    """The example's CG on the host with the oracle's Ax and plain numpy dots (same data generator)."""
    import numpy as np
    from oracle import ffi
    n3 = n ** 3
    N = E_total * n3
    xt = ffi.fill_uniform_f64(N, 11, 0.0, 1.0) - 0.5
    v = ffi.fill_uniform_f64(6 * N, 13, 0.0, 1.0).reshape(E_total, 6, n3)
    g = 0.2 * (v - 0.5)
    for f in (0, 3, 5):
        g[:, f, :] = 1.0 + 0.5 * v[:, f, :]
    g = np.ascontiguousarray(g.ravel())
    D = np.ascontiguousarray(ffi.gll_derivative(n)[0].ravel())
    b = ffi.ax(n, xt, g, D)
    x, r, p = np.zeros(N), b.copy(), b.copy()
    rr = float(r @ r)
    out = [dict(rr0=rr)]
    for it in range(iters):
        w = ffi.ax(n, p, g, D)
        pap = float(p @ w)
        alpha = rr / pap
        x += alpha * p
        r -= alpha * w
        rr_new = float(r @ r)
        p = r + (rr_new / rr) * p
        out.append(dict(pAp=pap, alpha=alpha, rr=rr_new))
        rr = rr_new
    return out


def _run_cg(world, E, n, tmp_path, mode=None, iters="400", tol="1e-9"):
    exe = ROOT / "libnomp_b200" / "build" / "cg_poisson"
    if not exe.exists():
        pytest.skip("examples/cg_poisson was not built")
    idfile = f"/dev/shm/nomp-test-cg-{os.getpid()}-{world}"
    procs = []
    for r in range(world):
        env = dict(os.environ, NOMP_INSTALL_DIR=str(ROOT / "libnomp_b200"))
        if world > 1:
            env.update(NOMP_COMM_SIZE=str(world), NOMP_COMM_RANK=str(r), NOMP_COMM_ID_FILE=idfile)
        procs.append(subprocess.Popen([str(exe), str(E), str(n), iters, tol, *([mode, "10"] if mode else []), "--nomp-backend", "cuda",
                                       "--nomp-device", str(r), "--nomp-verbose", "1"],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for f in [idfile] + [f"{idfile}.ipc.{r}" for r in range(world)]:
        try:
            os.unlink(f)
        except OSError:
            pass
    for p, o in zip(procs, outs):
        assert p.returncode in ((0, 2) if mode else (0,)), o[-3000:]     # 2: not converged within a fixed iteration count
    return [[json.loads(l) for l in o.splitlines() if l.startswith("{")] for o in outs]


def test_cg_example_matches_host_cg(tmp_path):
    """examples/cg_poisson.c (fused Ax+dot, fused update+dot, xpay through the public API) against the same CG on the
    host: the first iterations agree to 1e-10 relative, the solver converges, the true residual is small."""
    E, n = 48, 8
    lines = _run_cg(1, E, n, tmp_path)[0]
    ref = _cg_reference(E, n, 5)
    assert abs(lines[0]["rr0"] - ref[0]["rr0"]) <= 1e-12 * ref[0]["rr0"]
    for it in range(5):
        for key in ("pAp", "alpha", "rr"):
            assert abs(lines[1 + it][key] - ref[1 + it][key]) <= 1e-10 * abs(ref[1 + it][key]), (it, key)
    final = lines[-1]
    assert final["iterations"] < 400 and final["true_residual_rel"] < 1e-7


def test_cg_example_on_two_gpus(tmp_path):
    """Two ranks, each with its own block of elements of a 2E-element mesh: every rank sees the global scalars."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    E, n = 24, 8
    per_rank = _run_cg(2, E, n, tmp_path)
    ref = _cg_reference(2 * E, n, 5)
    for lines in per_rank:
        assert abs(lines[0]["rr0"] - ref[0]["rr0"]) <= 1e-12 * ref[0]["rr0"]
        for it in range(5):
            for key in ("pAp", "alpha", "rr"):
                assert abs(lines[1 + it][key] - ref[1 + it][key]) <= 1e-10 * abs(ref[1 + it][key]), (it, key)
    assert per_rank[0][1:6] == per_rank[1][1:6], "ranks must see bit-identical reduction results"


@pytest.mark.parametrize("world", [1, 2])
def test_cg_with_device_scalars_and_graph_replay(tmp_path, world):
    """examples/cg_poisson.c with its scalars in device memory ("device3": three launches per iteration, no host round
    trip) and the same iterations replayed from a CUDA graph ("graph"), on one rank and -- the collective call number
    being a device counter -- on two: 41 iterations give the same residual bit for bit either way, equal to the
    host-scalar run up to rounding, and all ranks see the same bits."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    E, n = 24, 8
    runs = {mode: _run_cg(world, E, n, tmp_path, mode, iters="41", tol="1e-30") for mode in ("host", "device3", "graph")}
    finals = {mode: [lines[-1] for lines in per_rank] for mode, per_rank in runs.items()}
    for mode, per_rank in finals.items():
        assert all(f["iterations"] == 41 and f["scalars"] == mode for f in per_rank), (mode, per_rank)
        assert all(f["rr_final"] == per_rank[0]["rr_final"] for f in per_rank), "ranks must see bit-identical scalars"
    assert finals["graph"][0]["rr_final"] == finals["device3"][0]["rr_final"]
    host = finals["host"][0]["rr_final"]
    assert abs(finals["graph"][0]["rr_final"] - host) <= 1e-6 * abs(host)      # 41 iterations of rounding differences
    assert finals["graph"][0]["true_residual_rel"] == finals["device3"][0]["true_residual_rel"]


def _poisson_reference(ex, ey, ez, n, iters):
    """examples/poisson_box.c on the host: oracle Ax, oracle gather-scatter, numpy dots."""
    import numpy as np
    from oracle import ffi
    D, xi = ffi.gll_derivative(n)
    N = n - 1
    P = np.polynomial.legendre.Legendre.basis(N)(xi)
    wt = 2.0 / (N * (N + 1.0) * P * P)
    hx, hy, hz = 1.0 / ex, 1.0 / ey, 1.0 / ez
    J = hx * hy * hz / 8.0
    ids = ffi.box_ids(n, ex, ey, ez)
    E = ex * ey * ez
    e = np.arange(E)
    e_x, e_y, e_z = e % ex, (e // ex) % ey, e // (ex * ey)
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    X = (e_x[:, None, None, None] + 0.5 * (xi[i] + 1.0)[None]) * hx
    Y = (e_y[:, None, None, None] + 0.5 * (xi[j] + 1.0)[None]) * hy
    Z = (e_z[:, None, None, None] + 0.5 * (xi[k] + 1.0)[None]) * hz
    B = np.broadcast_to((wt[i] * wt[j] * wt[k] * J)[None], X.shape)
    ue = (np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z)).reshape(-1)
    px, py, pz = N * ex + 1, N * ey + 1, N * ez + 1
    gid = ids - 1
    gx, gy, gz = gid % px, (gid // px) % py, gid // (px * py)
    mask = ((gx > 0) & (gx < px - 1) & (gy > 0) & (gy < py - 1) & (gz > 0) & (gz < pz - 1)).astype(np.float64)
    g = np.zeros((E, 6, n ** 3))
    Bf = np.ascontiguousarray(B).reshape(E, -1)
    g[:, 0], g[:, 3], g[:, 5] = Bf * 4 / hx ** 2, Bf * 4 / hy ** 2, Bf * 4 / hz ** 2
    g = np.ascontiguousarray(g.ravel())
    Dm = np.ascontiguousarray(D.ravel())
    c = 1.0 / ffi.gs(0, ffi.F64, ids, np.ones(ids.size))
    r = ffi.gs(0, ffi.F64, ids, (Bf.reshape(-1) * 3 * np.pi ** 2 * ue).copy()) * mask
    x, p = np.zeros_like(r), r.copy()
    rr = float((r * r * c).sum())
    out = [dict(rr0=rr)]
    for _ in range(iters):
        w = ffi.ax(n, p, g, Dm)
        pap = float(p @ w)
        w = ffi.gs(0, ffi.F64, ids, w)
        alpha = rr / pap
        x += alpha * p
        r -= alpha * (mask * w)
        rr_new = float((r * r * c).sum())
        p = r + (rr_new / rr) * p
        out.append(dict(pAp=pap, alpha=alpha, rr=rr_new))
        rr = rr_new
    return out


def _run_poisson(world, dims, tmp_path, extra=()):
    exe = ROOT / "libnomp_b200" / "build" / "poisson_box"
    if not exe.exists():
        pytest.skip("examples/poisson_box was not built")
    idfile = f"/dev/shm/nomp-test-poisson-{os.getpid()}-{world}"
    procs = []
    for r in range(world):
        env = dict(os.environ, NOMP_INSTALL_DIR=str(ROOT / "libnomp_b200"))
        if world > 1:
            env.update(NOMP_COMM_SIZE=str(world), NOMP_COMM_RANK=str(r), NOMP_COMM_ID_FILE=idfile)
        procs.append(subprocess.Popen([str(exe), *[str(d) for d in dims], *extra, "--nomp-backend", "cuda", "--nomp-device", str(r),
                                       "--nomp-verbose", "1"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for f in [idfile] + [f"{idfile}.ipc.{r}" for r in range(world)]:
        try:
            os.unlink(f)
        except OSError:
            pass
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    return [[json.loads(l) for l in o.splitlines() if l.startswith("{")] for o in outs]


def test_poisson_box_solves_the_pde(tmp_path):
    """examples/poisson_box.c: CG on the assembled operator (Ax + gather-scatter + mask).  The first iterations agree
    with the same algorithm on the host (oracle Ax, oracle gather-scatter) and the converged solution is the exact
    sin(pi x) sin(pi y) sin(pi z) up to the spectral discretisation error."""
    dims = (3, 4, 2, 8)
    lines = _run_poisson(1, dims, tmp_path)[0]
    ref = _poisson_reference(*dims, 5)
    assert abs(lines[0]["rr0"] - ref[0]["rr0"]) <= 1e-12 * ref[0]["rr0"]
    for it in range(5):
        for key in ("pAp", "alpha", "rr"):
            assert abs(lines[1 + it][key] - ref[1 + it][key]) <= 1e-10 * abs(ref[1 + it][key]), (it, key)
    final = lines[-1]
    assert final["iterations"] < 500 and final["residual_rel"] <= 1e-10 and final["max_error"] < 1e-6, final
    # refinement in the polynomial degree: the error of the discrete solution falls spectrally
    coarse = _run_poisson(1, (2, 2, 2, 6), tmp_path)[0][-1]["max_error"]
    fine = _run_poisson(1, (2, 2, 2, 10), tmp_path)[0][-1]["max_error"]
    assert fine < 1e-3 * coarse, (coarse, fine)


def test_poisson_box_on_two_gpus(tmp_path):
    """Slab partition over two ranks: same scalars as the single-rank run (the interface plane is summed through
    NVLink peer memory), same solution."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    dims = (3, 4, 4, 8)
    one = _run_poisson(1, dims, tmp_path)[0]
    two = _run_poisson(2, dims, tmp_path)
    assert two[0][0]["gs_shared_with_ranks"] == (7 * 3 + 1) * (7 * 4 + 1)
    for rank_lines in two:
        assert abs(rank_lines[0]["rr0"] - one[0]["rr0"]) <= 1e-12 * one[0]["rr0"]
        for it in range(5):
            for key in ("pAp", "alpha", "rr"):
                assert abs(rank_lines[1 + it][key] - one[1 + it][key]) <= 1e-10 * abs(one[1 + it][key]), (it, key)
        assert rank_lines[-1]["max_error"] < 1e-6 and rank_lines[-1]["iterations"] == two[0][-1]["iterations"]


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
