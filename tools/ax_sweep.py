#!/usr/bin/env python3
"""Device-time sweep of the libnompk kernels (run on the GPU box): Ax variants, map and reduce bandwidth.
Prints one JSON line per measurement to stdout and appends them to gpurun_out/sweep.jsonl."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from libnomp_b200 import capi  # noqa: E402

PEAK = 6548.5
try:
    PEAK = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
log = open(out / "sweep.jsonl", "a")


def emit(**kw):
    s = json.dumps(kw)
    print(s, flush=True)
    log.write(s + "\n")
    log.flush()


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def spin_up(seconds=0.4):
    """Bring the SM clocks up before anything is timed (the first kernels after an idle period run at low clocks)."""
    a = torch.rand(1 << 26, dtype=torch.float64, device="cuda")
    t0 = __import__("time").perf_counter()
    while __import__("time").perf_counter() - t0 < seconds:
        a.mul_(1.0000001)
        torch.cuda.synchronize()


def main():
    spin_up()
    lib = capi.nompk()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    which = sys.argv[1:] or ["ax", "map", "reduce"]
    if "ax" in which:
        for n, E in ((8, 32768), (8, 262144), (10, 16384), (10, 131072), (12, 65536), (6, 524288)):
            n3 = n ** 3
            u = torch.rand(E * n3, dtype=torch.float64, device="cuda")
            g = torch.rand(E * 6 * n3, dtype=torch.float64, device="cuda")
            D = torch.rand(n * n, dtype=torch.float64, device="cuda")
            w = torch.empty_like(u)
            for variant in range(18):
                lib.nompk_ax_set_variant(variant)

                def run():
                    capi.nompk_check(lib.nompk_ax_f64(n, E, u.data_ptr(), g.data_ptr(), D.data_ptr(), w.data_ptr(),
                                                      0, st))
                med, best = timeit(run)
                gb = E * n3 * 64 / 1e9
                emit(kernel="ax", n=n, E=E, variant=variant, ms=med, ms_min=best, gdofs=E * n3 / med / 1e6,
                     gbs=gb / med * 1e3, frac=gb / med * 1e3 / PEAK)
            lib.nompk_ax_set_variant(0)
            del u, g, w
    if "axrobust" in which:
        # Interleaved rounds: on a power-capped board the same kernel moves by several per cent between one timing and
        # the next, so every variant is timed once per round, round after round, and the median over rounds is kept.
        variants = [int(v) for v in os.environ.get("AX_VARIANTS", "0,1,7,8,11,21,22,23").split(",")]
        rounds = int(os.environ.get("AX_ROUNDS", "7"))
        shapes = [tuple(int(x) for x in sh.split(":")) for sh in
                  os.environ.get("AX_SHAPES", "10:131072,12:65536,8:262144,6:524288").split(",")]
        for n, E in shapes:
            n3 = n ** 3
            u = torch.rand(E * n3, dtype=torch.float64, device="cuda")
            g = torch.rand(E * 6 * n3, dtype=torch.float64, device="cuda")
            D = torch.rand(n * n, dtype=torch.float64, device="cuda")
            w = torch.empty_like(u)
            samples = {v: [] for v in variants}

            def run():
                capi.nompk_check(lib.nompk_ax_f64(n, E, u.data_ptr(), g.data_ptr(), D.data_ptr(), w.data_ptr(), 0, st))
            for _ in range(rounds):
                for v in variants:
                    lib.nompk_ax_set_variant(v)
                    samples[v].append(timeit(run, reps=10, warm=3)[0])
            lib.nompk_ax_set_variant(0)
            gb = E * n3 * 64 / 1e9
            for v in variants:
                ts = sorted(samples[v])
                med = ts[len(ts) // 2]
                emit(kernel="ax_interleaved", n=n, E=E, variant=v, rounds=rounds, ms=med, ms_min=ts[0], ms_max=ts[-1],
                     gdofs=E * n3 / med / 1e6, frac=gb / med * 1e3 / PEAK)
            del u, g, w
    if "axdot" in which:
        # the fused forms next to the plain operator, interleaved: Ax, Ax + p.Ap, (p <- r + beta p) + Ax + p.Ap
        rounds = int(os.environ.get("AX_ROUNDS", "7"))
        shapes = [tuple(int(x) for x in sh.split(":")) for sh in
                  os.environ.get("AX_SHAPES", "10:262144,12:65536,8:262144,6:524288").split(",")]
        ws = torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
        res = torch.zeros(1, dtype=torch.float64, device="cuda")
        for n, E in shapes:
            n3 = n ** 3
            u = torch.rand(E * n3, dtype=torch.float64, device="cuda")
            r = torch.rand(E * n3, dtype=torch.float64, device="cuda")
            g = torch.rand(E * 6 * n3, dtype=torch.float64, device="cuda")
            D = torch.rand(n * n, dtype=torch.float64, device="cuda")
            w = torch.empty_like(u)

            FL = [0]     # 2 = NOMPK_AX_D_ANTISYMMETRIC (timing only: the random D is not antisymmetric)

            def ax():
                capi.nompk_check(lib.nompk_ax_f64(n, E, u.data_ptr(), g.data_ptr(), D.data_ptr(), w.data_ptr(), FL[0], st))

            def axdot():
                capi.nompk_check(lib.nompk_ax_dot_f64(n, E, u.data_ptr(), g.data_ptr(), D.data_ptr(), w.data_ptr(),
                                                      res.data_ptr(), None, 0, ws.data_ptr(), FL[0], st))

            def axxpay():
                capi.nompk_check(lib.nompk_ax_xpay_dot_peers_f64(n, E, u.data_ptr(), r.data_ptr(), C.c_double(1e-9), None,
                                                                 g.data_ptr(), D.data_ptr(), w.data_ptr(), res.data_ptr(),
                                                                 None, 0, ws.data_ptr(), None, FL[0], st))

            def with_flags(fn, fl):
                def go():
                    FL[0] = fl
                    fn()
                    FL[0] = 0
                return go
            def variant_of(fn, v):
                def go():
                    lib.nompk_ax_set_variant(v)
                    fn()
                return go
            dot_variants = [int(v) for v in os.environ.get("AX_DOT_VARIANTS", "60,61,63,64,67").split(",") if v]
            kinds = [("ax", variant_of(ax, 0), 64), ("ax_dot", variant_of(axdot, 0), 64)] + \
                    [(f"ax_dot_v{v}", variant_of(axdot, v), 64) for v in dot_variants] + [("ax_xpay_dot", variant_of(axxpay, 0), 80)] + \
                    [("ax_eo", with_flags(variant_of(ax, 0), 2), 64), ("ax_dot_eo", with_flags(variant_of(axdot, 0), 2), 64),
                     ("ax_xpay_dot_eo", with_flags(variant_of(axxpay, 0), 2), 80)] + \
                    [(f"ax_dot_v{v}_eo", with_flags(variant_of(axdot, v), 2), 64) for v in dot_variants]
            samples = {k: [] for k, _, _ in kinds}
            for _ in range(rounds):
                for k, fn, _ in kinds:
                    samples[k].append(timeit(fn, reps=10, warm=3)[0])
            lib.nompk_ax_set_variant(0)
            for k, _, bpd in kinds:
                ts = sorted(samples[k])
                med = ts[len(ts) // 2]
                emit(kernel=k, n=n, E=E, rounds=rounds, ms=med, ms_min=ts[0], ms_max=ts[-1], gdofs=E * n3 / med / 1e6,
                     bytes_per_dof=bpd, frac=E * n3 * bpd / med / 1e6 / PEAK)
            del u, r, g, w
    if "map" in which:
        for lg in (20, 24, 26, 28):
            n = 1 << lg
            x = torch.rand(n, dtype=torch.float64, device="cuda")
            y = torch.rand(n, dtype=torch.float64, device="cuda")
            alpha = C.c_double(0.5)
            for op, name in ((capi.MAP_ADD, "add"), (capi.MAP_AXPY, "axpy")):
                def run():
                    capi.nompk_check(lib.nompk_map(op, capi.F64, n, y.data_ptr(), x.data_ptr(), None,
                                                   C.addressof(alpha), None, st))
                med, best = timeit(run)
                emit(kernel="map_" + name, n=n, ms=med, ms_min=best, gbs=n * 24 / med / 1e6, frac=n * 24 / med / 1e6 / PEAK)
            def cp():
                y.copy_(x)
            med, best = timeit(cp)
            emit(kernel="torch_copy", n=n, ms=med, ms_min=best, gbs=n * 16 / med / 1e6, frac=n * 16 / med / 1e6 / PEAK)
    if "reduce" in which:
        ws = torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
        res = torch.zeros(1, dtype=torch.float64, device="cuda")
        for lg in (20, 24, 28):
            n = 1 << lg
            x = torch.rand(n, dtype=torch.float64, device="cuda")
            y = torch.rand(n, dtype=torch.float64, device="cuda")
            xi = torch.randint(-2 ** 62, 2 ** 62, (n,), dtype=torch.int64, device="cuda")
            for name, dt, a, b, bytes_ in (("sum_f64", capi.F64, x, None, 8), ("dot_f64", capi.F64, x, y, 16),
                                            ("sum_i64", capi.I64, xi, None, 8)):
                def run():
                    capi.nompk_check(lib.nompk_reduce(capi.RED_SUM, dt, n, a.data_ptr(),
                                                      b.data_ptr() if b is not None else None, res.data_ptr(), None, 0,
                                                      ws.data_ptr(), st))
                med, best = timeit(run)
                emit(kernel=name, n=n, ms=med, ms_min=best, gbs=n * bytes_ / med / 1e6,
                     frac=n * bytes_ / med / 1e6 / PEAK)


if __name__ == "__main__":
    main()
