#!/usr/bin/env python3
"""Time nompk_gs_apply on a box of ex*ey*ez elements with n points per direction (one GPU).
usage: gs_bench.py [n] [ex] [ey] [ez] [reps]   ->  one JSON line"""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from libnomp_b200 import capi  # noqa: E402

n, ex, ey, ez = (int(a) for a in (sys.argv[1:5] + ["8", "64", "64", "64"][len(sys.argv[1:5]):]))
reps = int(sys.argv[5]) if len(sys.argv) > 5 and sys.argv[5].isdigit() else 20
lib = capi.nompk()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
# global ids of the box, built on the device (same numbering as oracle/ffi.py: box_ids)
N = n - 1
px, py = N * ex + 1, N * ey + 1
e = torch.arange(ex * ey * ez, device="cuda")
e_x, e_y, e_z = e % ex, (e // ex) % ey, e // (ex * ey)
p = torch.arange(n ** 3, device="cuda")
i, j, k = p % n, (p // n) % n, p // (n * n)
ids = (1 + (e_x[:, None] * N + i[None]) + px * ((e_y[:, None] * N + j[None]) + py * (e_z[:, None] * N + k[None]))).reshape(-1).contiguous()
del e, p
h = C.c_void_p()
torch.cuda.synchronize()
t0 = time.perf_counter()
capi.nompk_check(lib.nompk_gs_create(ids.data_ptr(), ids.numel(), C.byref(h), st))
xb = C.c_size_t()
capi.nompk_check(lib.nompk_gs_finalize_setup(h, 0, 1, C.byref(xb), st))
torch.cuda.synchronize()
setup_s = time.perf_counter() - t0
stats = (C.c_size_t * 8)()
lib.nompk_gs_stats(h, C.byref(stats))
ndof, distinct, groups, copies = (int(stats[q]) for q in range(4))
del ids
v = torch.rand(ndof, dtype=torch.float64, device="cuda")
t_warm = time.perf_counter()
while time.perf_counter() - t_warm < (0.0 if "--no-warmup" in sys.argv else 0.3):    # clocks up before timing
    for _ in range(5):
        capi.nompk_check(lib.nompk_gs_apply(h, capi.RED_MIN, capi.F64, v.data_ptr(), None, st))
    torch.cuda.synchronize()
capi.nompk_check(lib.nompk_gs_apply(h, capi.RED_MIN, capi.F64, v.data_ptr(), None, st))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(reps):
    capi.nompk_check(lib.nompk_gs_apply(h, capi.RED_MIN, capi.F64, v.data_ptr(), None, st))
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
alg = copies * (8 + 8 + 4) + (groups + 1) * 4
kernel = "gs_local_kernel" if os.environ.get("NOMPK_GS_KERNEL") == "group" else "gs_local_warp_kernel"
print(json.dumps({"kernel": kernel, "n": n, "elements": ex * ey * ez, "dofs": ndof, "distinct": distinct,
                  "groups": groups, "copies": copies, "setup_s": round(setup_s, 4), "ms": ms,
                  "algorithmic_bytes": alg, "GB/s": alg / ms / 1e6, "bytes_per_dof": alg / ndof,
                  "ns_per_dof_vs_ax64": {"gs": ms * 1e6 / ndof, "ax_at_6548GB/s": 64 / 6548.5}}))
lib.nompk_gs_destroy(h)
