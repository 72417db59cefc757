#!/usr/bin/env python3
"""Run one libnompk kernel a few times (target for `ncu`).  usage: run_kernel_once.py ax|axdot|axeo|axdoteo|map|reduce [n] [E|len] [variant] [reps]
(axeo / axdoteo: with NOMPK_AX_D_ANTISYMMETRIC -- the even-odd kernels; the random D is not antisymmetric, the numbers
are meaningless, time and traffic are those of the real thing)"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from libnomp_b200 import capi  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "ax"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
size = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
lib = capi.nompk()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flags = 2 if kind in ("axeo", "axdoteo") else 0
if kind in ("ax", "axdot", "axeo", "axdoteo"):
    n3 = n ** 3
    u = torch.rand(size * n3, dtype=torch.float64, device="cuda")
    g = torch.rand(size * 6 * n3, dtype=torch.float64, device="cuda")
    D = torch.rand(n * n, dtype=torch.float64, device="cuda")
    w = torch.empty_like(u)
    lib.nompk_ax_set_variant(variant)
    ws = torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        if kind in ("ax", "axeo"):
            capi.nompk_check(lib.nompk_ax_f64(n, size, u.data_ptr(), g.data_ptr(), D.data_ptr(), w.data_ptr(), flags, st))
        else:
            capi.nompk_check(lib.nompk_ax_dot_f64(n, size, u.data_ptr(), g.data_ptr(), D.data_ptr(), w.data_ptr(),
                                                  res.data_ptr(), None, 0, ws.data_ptr(), flags, st))
elif kind == "map":
    x = torch.rand(size, dtype=torch.float64, device="cuda")
    y = torch.rand(size, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        capi.nompk_check(lib.nompk_map(capi.MAP_ADD, capi.F64, size, y.data_ptr(), x.data_ptr(), None, None, None, st))
else:
    x = torch.rand(size, dtype=torch.float64, device="cuda")
    y = torch.rand(size, dtype=torch.float64, device="cuda")
    ws = torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        capi.nompk_check(lib.nompk_reduce(capi.RED_SUM, capi.F64, size, x.data_ptr(), y.data_ptr() if n else None,
                                          res.data_ptr(), None, 0, ws.data_ptr(), st))
torch.cuda.synchronize()
