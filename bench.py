#!/usr/bin/env python3
"""Headline benchmark of libnomp_b200 (driver contract: one JSON line on stdout from rank 0).

Metric (BASELINE.json): SEM Ax GDOF/s, N = 7 (n = 8 points per direction), fp64.
Workload: the local Poisson operator on E = 262144 hexahedral elements (BASELINE.json configs[3], the size the
metric's target is quoted on; it fits one B200: u 1 GiB + w 1 GiB + six geometric factors 6 GiB), element-
partitioned over the ranks (E / N elements per GPU, strong scaling, no collective on the Ax path).

A "step" is one application w = A u through the reference-facing API: nomp_run() of the jitted canonical Ax kernel
string (libnomp_b200/python/nomp_bridge/families.py), i.e. exactly one launch of the hand-written ax_kernel.

  value     GDOF/s with u, g, D, w resident in HBM, timed with CUDA events on the backend's own stream;
  e2e       the same step with HOST buffers: nomp_update(u, NOMP_TO) from pinned memory, nomp_run, and
            nomp_update(w, NOMP_FROM) inside the timed region (the geometric factors stay mapped, as in a solver);
  roofline  64 algorithmic bytes per DOF x DOFs per launch / average launch duration, against the measured copy
            bandwidth in MEASURED_PEAKS.json;
  cpu_baseline  the oracle's serial C loop nest (oracle/nomp_oracle.c) on a bounded sample, 1 core;
  extras    axpy / sum / dot bandwidth (configs[0], [1]) and a CG-style step (Ax + dot + axpy, configs[3]) that
            exercises the NCCL allreduce of the dot product when N > 1.

`--impl reference` times the reference's CPU path.  libnomp itself cannot be built here (needs SymEngine, loopy,
pymbolic, islpy, libclang and a CPU OpenCL platform, none installed and no network), so this arm runs the oracle
port of the same loop nest on all host threads (the stated stand-in for the OpenCL/pocl backend).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_POINTS = 8                     # n = N + 1, N = 7
E_TOTAL = 262144
BYTES_PER_DOF = 64               # u 8 + six geometric factors 48 + w 8
METRIC = "SEM Ax GDOF/s (N=7 fp64)"
UNIT = "GDOF/s"


def ncu_traffic(n, E):
    """DRAM bytes per launch of the Ax kernel from the committed ncu capture of the same shape, else None."""
    try:
        t = json.load(open(ROOT / "profiles" / "ncu_traffic.json"))["ax_kernel"].get(f"n{n}_E{E}")
        return None if t is None else t["read_bytes"] + t["write_bytes"]
    except Exception:
        return None


def measured_peak():
    try:
        return float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the CPU path (oracle port, all host threads)
# ---------------------------------------------------------------------------------------------------------------------

def cpu_ax_sample(E_sample: int, n: int):
    import numpy as np
    from oracle import ffi
    n3 = n ** 3
    u = ffi.fill_uniform_f64(E_sample * n3, 1234, 0.5, 1.5)
    g = ffi.fill_uniform_f64(E_sample * 6 * n3, 99, 0.5, 1.5)
    D, _ = ffi.gll_derivative(n)
    D = np.ascontiguousarray(D.ravel())
    w = np.empty_like(u)
    return ffi, u, g, D, w


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = N_POINTS
    E_sample = 32768
    ffi, u, g, D, w = cpu_ax_sample(E_sample, n)
    lib = ffi.lib()
    cores = lib.oracle_num_threads()

    def step():
        lib.oracle_ax_f64_mt(n, E_sample, u.ctypes.data, g.ctypes.data, D.ctypes.data, w.ctypes.data)

    for _ in range(max(1, min(args.warmup, 3))):
        step()
    steps = max(1, args.steps)
    budget_s = 120.0
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    value = E_sample * n ** 3 * done / dt / 1e9
    sample = (f"Ax N=7 on E={E_sample} elements per step (1/{E_TOTAL // E_sample} of the workload), {done} steps, "
              f"oracle C port of the loop nest on {cores} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"SEM local Poisson Ax, N=7 (n=8), fp64, E={E_TOTAL} total; CPU arm timed on a bounded "
                               f"sample of E={E_sample} elements per step", "E_total": E_TOTAL, "n": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML, ~2 ms period)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------

def box_slab_ids(n, ex, ey, ez, z0, nz):
    """Global ids (1-based, lexicographic over the (n-1)*e + 1 points per direction) of the degrees of freedom of the
    element layers [z0, z0 + nz) of an ex x ey x ez box, element-major with x fastest (the layout of u and w)."""
    import numpy as np
    N = n - 1
    px, py = N * ex + 1, N * ey + 1
    pt = np.arange(n, dtype=np.int64)
    gx = (np.arange(ex, dtype=np.int64)[:, None] * N + pt[None]).reshape(1, 1, ex, 1, 1, n)
    gy = (np.arange(ey, dtype=np.int64)[:, None] * N + pt[None]).reshape(1, ey, 1, 1, n, 1)
    gz = (np.arange(z0, z0 + nz, dtype=np.int64)[:, None] * N + pt[None]).reshape(nz, 1, 1, n, 1, 1)
    out = np.empty((nz, ey, ex, n, n, n), dtype=np.int64)
    np.multiply(gz, py, out=out)
    out += gy
    out *= px
    out += gx
    out += 1
    return out.reshape(-1)


def gll_derivative_matrix(n: int):
    """D[a][l] = l_l'(x_a) on the n Gauss-Lobatto-Legendre nodes (numpy; the product path does not touch oracle/)."""
    import numpy as np
    from numpy.polynomial import legendre as L
    N = n - 1
    cN = np.zeros(N + 1)
    cN[N] = 1.0
    x = np.concatenate(([-1.0], np.sort(L.legroots(L.legder(cN))), [1.0]))
    PN = L.legval(x, cN)
    D = np.zeros((n, n))
    for a in range(n):
        for l in range(n):
            if a != l:
                D[a, l] = PN[a] / (PN[l] * (x[a] - x[l]))
    D[0, 0] = -N * (N + 1) / 4.0
    D[N, N] = N * (N + 1) / 4.0
    return D

FILL_KERNEL = """
void nomp_fill(double *a, int n, int seed) {
  for (int i = 0; i < n; i++)
    a[i] = 0.5 + (double)((((unsigned)i + (unsigned)seed) * 2654435761u >> 9) & 1023u) * 0.0009765625;
}
"""
AXPY_KERNEL = "void nomp_axpy(double *y, const double *x, double alpha, int n) { for (int i = 0; i < n; i++) y[i] += alpha * x[i]; }"
ADD_KERNEL = "void nomp_add(double *a, const double *b, int n) { for (int i = 0; i < n; i++) a[i] += b[i]; }"
SUM_KERNEL = "void nomp_sum(const double *a, int n, double *s) { for (int i = 0; i < n; i++) s[0] += a[i]; }"
DOT_KERNEL = "void nomp_dot(const double *a, const double *b, int n, double *s) { for (int i = 0; i < n; i++) s[0] += a[i] * b[i]; }"
ISUM_KERNEL = "void nomp_isum(const long *a, int n, long *s) { for (int i = 0; i < n; i++) s[0] += a[i]; }"


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from libnomp_b200 import capi
    sys.path.insert(0, str(ROOT / "libnomp_b200" / "python"))
    from nomp_bridge.families import AX_KERNEL_SOURCE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libnomp_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        token = [f"{os.getpid()}-{time.time_ns()}" if rank == 0 else None]
        dist.broadcast_object_list(token, src=0)
        os.environ["NOMP_COMM_SIZE"] = str(world)
        os.environ["NOMP_COMM_RANK"] = str(rank)
        os.environ["NOMP_COMM_ID_FILE"] = f"/dev/shm/nomp-nccl-{os.environ.get('MASTER_PORT', '0')}-{token[0]}"

    def barrier():
        if world > 1:
            dist.barrier()

    lib = capi.nomp()
    capi.check(capi.init(backend="cuda", device=local_rank, verbose=1))
    stream = torch.cuda.ExternalStream(lib.nomp_b200_stream(), device=torch.device("cuda", local_rank))

    n = N_POINTS
    n3 = n ** 3
    if E_TOTAL % world:
        raise SystemExit(f"E={E_TOTAL} is not divisible by {world} ranks")
    E = E_TOTAL // world
    ndof = E * n3

    # ---- host buffers: u and w are real pinned buffers (the e2e leg copies them), g is only a key: its device image
    # is generated on the device, the host pages are never touched
    u_host = torch.empty(ndof, dtype=torch.float64).pin_memory()
    w_host = torch.empty(ndof, dtype=torch.float64).pin_memory()
    rng = np.random.default_rng(1234 + rank)
    u_host.numpy()[:] = rng.uniform(0.5, 1.5, ndof)
    g_key = np.empty(6 * ndof, dtype=np.float64)
    D_host = np.ascontiguousarray(gll_derivative_matrix(n).ravel())
    up, wp, gp, Dp = u_host.data_ptr(), w_host.data_ptr(), g_key.ctypes.data, D_host.ctypes.data

    capi.check(capi.update(up, 0, ndof, 8, capi.NOMP_TO))
    capi.check(capi.update(wp, 0, ndof, 8, capi.NOMP_ALLOC))
    capi.check(capi.update(gp, 0, 6 * ndof, 8, capi.NOMP_ALLOC))
    capi.check(capi.update(Dp, 0, n * n, 8, capi.NOMP_TO))

    no_clause = capi.clauses()
    err, fill_id = capi.jit(FILL_KERNEL, no_clause, [("a", 8, capi.NOMP_PTR), ("n", 4, capi.NOMP_INT), ("seed", 4, capi.NOMP_INT)])
    capi.check(err)
    capi.check(capi.run(fill_id, gp, C.c_int(6 * ndof), C.c_int(17 + rank)))

    err, ax_id = capi.jit(AX_KERNEL_SOURCE, no_clause,
                          [("w", 8, capi.NOMP_PTR), ("u", 8, capi.NOMP_PTR), ("g", 8, capi.NOMP_PTR),
                           ("D", 8, capi.NOMP_PTR), ("E", 4, capi.NOMP_INT), ("n", 4, capi.NOMP_INT | capi.NOMP_JIT, C.c_int(n))])
    capi.check(err)
    info = lib.nomp_b200_prog_info(ax_id).decode()
    if "kind=native family=ax" not in info:
        raise SystemExit(f"the Ax kernel string was not routed to the hand-written kernel: {info}")
    E_c = C.c_int(E)

    def ax_step():
        capi.check(capi.run(ax_id, wp, up, gp, Dp, E_c))

    def sync():
        capi.check(lib.nomp_sync())

    # ---- device-resident timing ----------------------------------------------------------------------------------
    # warm-up: at least W steps, and at least ~0.3 s of work so that the SM clocks have left their idle state
    warm = max(args.warmup, 3)
    t_warm = time.perf_counter()
    done_warm = 0
    while done_warm < warm or time.perf_counter() - t_warm < 0.3:
        ax_step()
        done_warm += 1
        if done_warm % 8 == 0:
            sync()
    sync()
    barrier()
    launches0 = lib.nomp_b200_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            ax_step()
        ev1.record(stream)
        sync()
    barrier()
    launches = lib.nomp_b200_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    ms_per_step = ms_total_max / args.steps
    value = E_TOTAL * n3 / (ms_per_step * 1e-3) / 1e9
    # roofline of the dominant (only) kernel, this rank's launches
    peak, peak_kind = measured_peak()
    kernel_ms = ms_total / args.steps
    achieved = ndof * BYTES_PER_DOF / (kernel_ms * 1e-3) / 1e9

    # ---- end to end: host buffers in, host buffers out ----------------------------------------------------------------
    # (a) blocking, reference API only: nomp_update(u, TO); nomp_run; nomp_update(w, FROM)
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_blocking_step():
        capi.check(capi.update(up, 0, ndof, 8, capi.NOMP_TO))
        ax_step()
        capi.check(capi.update(wp, 0, ndof, 8, capi.NOMP_FROM))

    def timed_wall(fn, reps):
        fn()
        sync()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        sync()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e_blocking = E_TOTAL * n3 * e2e_steps / timed_wall(e2e_blocking_step, e2e_steps) / 1e9
    checksum = float(w_host[:: max(1, ndof // 4096)].sum())

    # (b) pipelined: the arrays are mapped in NCHUNK element blocks; per block nomp_b200_update_async(u_c, TO),
    # nomp_run on the block, nomp_b200_update_async(w_c, FROM); one nomp_sync per step.  Both PCIe directions and the
    # kernel overlap.  The big mappings of u, w, g are released first (same host buffers, new device images).
    NCHUNK = int(os.environ.get("NOMP_BENCH_E2E_BLOCKS", "8"))
    Ec = E // NCHUNK
    e2e_value, e2e_note = e2e_blocking, "blocking only (E per GPU not divisible into 8 blocks)"
    if Ec * NCHUNK == E and Ec > 0:
        for ptr, cnt in ((up, ndof), (wp, ndof), (gp, 6 * ndof)):
            capi.check(capi.update(ptr, 0, cnt, 8, capi.NOMP_FREE))
        cdof = Ec * n3
        Ec_c = C.c_int(Ec)
        blocks = []
        for c in range(NCHUNK):
            uc, wc, gc = up + c * cdof * 8, wp + c * cdof * 8, gp + c * 6 * cdof * 8
            capi.check(capi.update(uc, 0, cdof, 8, capi.NOMP_TO))
            capi.check(capi.update(wc, 0, cdof, 8, capi.NOMP_ALLOC))
            capi.check(capi.update(gc, 0, 6 * cdof, 8, capi.NOMP_ALLOC))
            capi.check(capi.run(fill_id, gc, C.c_int(6 * cdof), C.c_int(17 + rank + 6 * c * cdof)))
            blocks.append((uc, wc, gc))

        def e2e_pipelined_step():
            for uc, wc, gc in blocks:
                capi.check(capi.update_async(uc, 0, cdof, 8, capi.NOMP_TO))
                capi.check(capi.run(ax_id, wc, uc, gc, Dp, Ec_c))
                capi.check(capi.update_async(wc, 0, cdof, 8, capi.NOMP_FROM))
            sync()

        w_host.zero_()
        e2e_value = E_TOTAL * n3 * e2e_steps / timed_wall(e2e_pipelined_step, e2e_steps) / 1e9
        checksum_p = float(w_host[:: max(1, ndof // 4096)].sum())
        if abs(checksum_p - checksum) > 1e-9 * abs(checksum):
            raise SystemExit(f"pipelined e2e result differs from the blocking one: {checksum_p} vs {checksum}")
        e2e_note = (f"per step and per block of {Ec} elements ({NCHUNK} blocks): nomp_b200_update_async(u, TO) from pinned "
                    "host memory, nomp_run(Ax), nomp_b200_update_async(w, FROM); one nomp_sync per step; geometric "
                    "factors and D stay mapped.  blocking_value: the same step with the blocking reference calls "
                    "nomp_update(TO) / nomp_run / nomp_update(FROM) on the whole arrays")
        # hand the extras their big mappings back
        for uc, wc, gc in blocks:
            for ptr, cnt in ((uc, cdof), (wc, cdof), (gc, 6 * cdof)):
                capi.check(capi.update(ptr, 0, cnt, 8, capi.NOMP_FREE))
        capi.check(capi.update(up, 0, ndof, 8, capi.NOMP_TO))
        capi.check(capi.update(wp, 0, ndof, 8, capi.NOMP_ALLOC))
        capi.check(capi.update(gp, 0, 6 * ndof, 8, capi.NOMP_ALLOC))
        capi.check(capi.run(fill_id, gp, C.c_int(6 * ndof), C.c_int(17 + rank)))

    # ---- extras: maps, reductions, CG-style step ---------------------------------------------------------------------------
    extras = {}
    try:
        extras = run_extras(args, capi, lib, torch, dist, world, rank, stream, up, wp, gp, Dp, ax_id, E_c, ndof, peak, barrier)
    except Exception as exc:  # extras never invalidate the headline
        extras = {"error": repr(exc)}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle's serial loop nest on a bounded sample ------------------------------
    cpu = cpu_all = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        E_sample = 32768
        ffi, cu, cg_, cD, cw = cpu_ax_sample(E_sample, n)
        fn = ffi.lib().oracle_ax_f64
        fn(n, 256, cu.ctypes.data, cg_.ctypes.data, cD.ctypes.data, cw.ctypes.data)
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 200):   # about 10 s of CPU work
            fn(n, E_sample, cu.ctypes.data, cg_.ctypes.data, cD.ctypes.data, cw.ctypes.data)
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": E_sample * n3 * reps / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"Ax N=7 on E={E_sample} elements (1/8 of the workload) x {reps} repetitions ({dt:.1f} s), serial C "
                         "loop nest of oracle/nomp_oracle.c (gcc -O2 -march=native -ffp-contract=off)"}
        # the same loop nest on all host cores: stand-in for the reference's OpenCL backend on a CPU OpenCL platform
        # (pocl), which cannot be installed here
        fn_mt = ffi.lib().oracle_ax_f64_mt
        cores = ffi.lib().oracle_num_threads()
        fn_mt(n, E_sample, cu.ctypes.data, cg_.ctypes.data, cD.ctypes.data, cw.ctypes.data)
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 5.0 and reps < 2000):
            fn_mt(n, E_sample, cu.ctypes.data, cg_.ctypes.data, cD.ctypes.data, cw.ctypes.data)
            reps += 1
        dt = time.perf_counter() - t0
        cpu_all = {"value": E_sample * n3 * reps / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"same loop nest split over {cores} host threads (pthreads), E={E_sample} x {reps} repetitions "
                             f"({dt:.1f} s); stand-in for libnomp's OpenCL backend on pocl, not installable here"}

    allreduce_path = "nvlink-kernel" if lib.nomp_b200_comm_uses_nvlink_kernel() else ("nccl" if world > 1 else "none")
    try:
        capi.check(lib.nomp_finalize_excluding_interpreter())
    except Exception as exc:  # a failure while shutting down must not cost the measured line
        extras = dict(extras, finalize_error=repr(exc))
    if world > 1:
        dist.barrier()
        if rank == 0:   # rendezvous files of this job
            import glob
            for f in glob.glob(os.environ["NOMP_COMM_ID_FILE"] + "*"):
                try:
                    os.unlink(f)
                except OSError:
                    pass
        dist.destroy_process_group()
    if rank != 0:
        return 0

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"SEM local Poisson Ax, N=7 (n=8), fp64, E={E_TOTAL} hexahedral elements total, "
                               f"{E} per GPU (element-partitioned), through nomp_jit/nomp_run",
                   "E_total": E_TOTAL, "E_per_gpu": E, "n": n, "bytes_per_dof": BYTES_PER_DOF,
                   "l2": f"no flush needed: each step streams {ndof * BYTES_PER_DOF / 1e6:.0f} MB per GPU, larger than the 126 MB L2",
                   "kernel": info, "allreduce": allreduce_path},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(n, E), "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                     "kernel": "nompk::ax_kernel<8,...>", "algorithmic_bytes_per_launch": ndof * BYTES_PER_DOF,
                     "avg_launch_ms": kernel_ms},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ndof * 8 * world, "d2h_bytes_per_step": ndof * 8 * world,
                "steps": e2e_steps, "blocking_value": e2e_blocking, "checksum": checksum, "note": e2e_note},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "cpu_baseline": cpu,
        "cpu_baseline_all_cores": cpu_all,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_extras(args, capi, lib, torch, dist, world, rank, stream, up, wp, gp, Dp, ax_id, E_c, ndof, peak, barrier):
    """Map / reduce bandwidth on this rank's slice and a CG-style step (Ax + dot + axpy) over all ranks."""
    out = {}
    none = capi.clauses()
    P, I, F = capi.NOMP_PTR, capi.NOMP_INT, capi.NOMP_FLOAT

    def timed(fn, reps):
        for _ in range(3):
            fn()
        capi.check(lib.nomp_sync())
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        capi.check(lib.nomp_sync())
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nvec = min(ndof, 1 << 28)
    n_c = C.c_int(nvec)
    alpha = C.c_double(0.5)
    err, axpy_id = capi.jit(AXPY_KERNEL, none, [("y", 8, P), ("x", 8, P), ("alpha", 8, F), ("n", 4, I)])
    capi.check(err)
    err, dot_id = capi.jit(DOT_KERNEL, capi.clauses(("reduce", "s", "+")), [("a", 8, P), ("b", 8, P), ("n", 4, I), ("s", 8, F)])
    capi.check(err)
    err, sum_id = capi.jit(SUM_KERNEL, capi.clauses(("reduce", "s", "+")), [("a", 8, P), ("n", 4, I), ("s", 8, F)])
    capi.check(err)
    s = C.c_double(0.0)

    ms = timed(lambda: capi.check(capi.run(axpy_id, wp, up, alpha, n_c)), 20)
    out["axpy_f64"] = {"n_per_gpu": nvec, "ms": ms, "GB/s_per_gpu": nvec * 24 / ms / 1e6, "frac_of_peak": nvec * 24 / ms / 1e6 / peak}
    ms = timed(lambda: capi.check(capi.run(sum_id, up, n_c, s)), 20)
    out["sum_f64"] = {"n_per_gpu": nvec, "ms": ms, "GB/s_per_gpu": nvec * 8 / ms / 1e6, "frac_of_peak": nvec * 8 / ms / 1e6 / peak,
                      "note": "includes the host-visible result (and the allreduce over ranks when n_gpus > 1)"}
    ms = timed(lambda: capi.check(capi.run(dot_id, up, wp, n_c, s)), 20)
    out["dot_f64"] = {"n_per_gpu": nvec, "ms": ms, "GB/s_per_gpu": nvec * 16 / ms / 1e6, "frac_of_peak": nvec * 16 / ms / 1e6 / peak}

    # small vector add, configs[0]: n = 2^20 (24 MiB: L2-resident and launch-latency-bound -- labelled as such)
    n20 = C.c_int(1 << 20)
    err, add_id = capi.jit(ADD_KERNEL, none, [("a", 8, P), ("b", 8, P), ("n", 4, I)])
    capi.check(err)
    ms = timed(lambda: capi.check(capi.run(add_id, wp, up, n20)), 200)
    out["add_f64_2^20"] = {"n": 1 << 20, "us": ms * 1e3, "GB/s": (1 << 20) * 24 / ms / 1e6, "note": "L2-resident, launch-bound"}

    # CG-style step, configs[3]: w = A p ; pAp = p.w (allreduce over ranks) ; x += alpha p   -> 104 B/DOF
    nd = C.c_int(ndof)

    def cg_step():
        capi.check(capi.run(ax_id, wp, up, gp, Dp, E_c))
        capi.check(capi.run(dot_id, up, wp, nd, s))
        capi.check(capi.run(axpy_id, wp, up, alpha, nd))

    # the same step with p.Ap fused into the Ax kernel (row f): Ax+dot (allreduce) + axpy -> 88 B/DOF
    try:
        from nomp_bridge.families import AX_DOT_KERNEL_SOURCE
        err, axdot_id = capi.jit(AX_DOT_KERNEL_SOURCE, capi.clauses(("reduce", "pap", "+")),
                                 [("w", 8, P), ("u", 8, P), ("g", 8, P), ("D", 8, P), ("E", 4, I),
                                  ("n", 4, I | capi.NOMP_JIT, C.c_int(N_POINTS)), ("pap", 8, F)])
        capi.check(err)
        s2 = C.c_double(0.0)

        def cg_step_fused():
            capi.check(capi.run(axdot_id, wp, up, gp, Dp, E_c, s2))
            capi.check(capi.run(axpy_id, wp, up, alpha, nd))

        # a fresh w for both variants so that the two p.Ap values are comparable
        capi.check(capi.run(ax_id, wp, up, gp, Dp, E_c))
        capi.check(capi.run(dot_id, up, wp, nd, s))
        capi.check(capi.run(axdot_id, wp, up, gp, Dp, E_c, s2))
        rel = abs(s.value - s2.value) / max(abs(s.value), 1e-300)
        ms_f = timed(cg_step_fused, 20)
        out["cg_step_fused"] = {"what": "Ax with p.Ap fused (allreduce over ranks) + axpy, 88 B/DOF", "ms": ms_f,
                                "GDOF/s": ndof * world / ms_f / 1e6, "GB/s_per_gpu": ndof * 88 / ms_f / 1e6,
                                "frac_of_peak": ndof * 88 / ms_f / 1e6 / peak, "pAp_rel_diff_vs_unfused": rel}
    except Exception as exc:
        out["cg_step_fused"] = {"error": repr(exc)}

    # gather-scatter (row f): the mesh seen as a 64 x 64 x (64 / ranks) slab of a 64^3 box of elements, lexicographic
    # global numbering; "min" so that repeated application leaves the data alone.  Every rank reports whether its setup
    # worked and all ranks agree before anything collective is timed.
    h, gs_error = None, None
    try:
        ex = ey = 64
        nz = int(E_c.value) // (ex * ey)
        if nz * ex * ey != int(E_c.value) or nz * world != 64:
            raise ValueError("gather-scatter extra needs E_total = 64^3 split into whole layers")
        ids = box_slab_ids(N_POINTS, ex, ey, 64, rank * nz, nz)
        h = capi.gs_setup(ids)
        del ids
    except Exception as exc:
        gs_error = repr(exc)
    ok = torch.tensor([0 if gs_error else 1], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 1:
        try:
            info = capi.gs_info(h)
            gs_call = lambda: capi.check(lib.nomp_b200_gs(h, wp, 8, capi.NOMP_FLOAT, b"min"))  # noqa: E731
            ms_gs = timed(gs_call, 20)
            alg = info["copies"] * 20 + (info["groups"] + 1) * 4
            out["gather_scatter"] = {"what": "nomp_b200_gs(min) on this rank's slab of a 64^3-element box, N=7", "ms": ms_gs,
                                     "groups": info["groups"], "copies": info["copies"], "ids_shared_with_other_ranks": info["shared_ids"],
                                     "algorithmic_bytes": alg, "GB/s_per_gpu": alg / ms_gs / 1e6, "frac_of_peak": alg / ms_gs / 1e6 / peak,
                                     "note": "algorithmic = 8 B read + 8 B written + 4 B index per shared copy, 4 B per group; the DRAM traffic is about 1.5x that (whole 64-byte lines of the vector are touched)"}
            if "error" not in out.get("cg_step_fused", {"error": 1}):
                def cg_step_assembled():
                    capi.check(capi.run(axdot_id, wp, up, gp, Dp, E_c, s2))
                    capi.check(lib.nomp_b200_gs(h, wp, 8, capi.NOMP_FLOAT, b"min"))
                    capi.check(capi.run(axpy_id, wp, up, alpha, nd))
                ms_a = timed(cg_step_assembled, 20)
                out["cg_step_assembled"] = {"what": "Ax with p.Ap fused + gather-scatter (interface planes over NVLink) + axpy", "ms": ms_a,
                                            "GDOF/s": ndof * world / ms_a / 1e6}
        except Exception as exc:
            out["gather_scatter"] = {"error": repr(exc)}
    else:
        out["gather_scatter"] = {"error": gs_error or "setup failed on another rank"}
    if h is not None:
        try:
            capi.check(lib.nomp_b200_gs_free(h))
        except Exception as exc:
            out.setdefault("gather_scatter", {})["free_error"] = repr(exc)

    ms = timed(cg_step, 20)
    total_dof = ndof * world
    out["cg_step"] = {"what": "Ax + dot (allreduce over ranks) + axpy, 104 B/DOF", "ms": ms, "GDOF/s": total_dof / ms / 1e6,
                      "GB/s_per_gpu": ndof * 104 / ms / 1e6, "frac_of_peak": ndof * 104 / ms / 1e6 / peak, "pAp": s.value}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
