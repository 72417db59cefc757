#!/usr/bin/env python3
"""Target for `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:ax_kernel`:
launches every listed Ax variant twice on one shape.  usage: ax_dram_probe.py n E v0,v1,..."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from libnomp_b200 import capi  # noqa: E402

n, E = int(sys.argv[1]), int(sys.argv[2])
variants = [int(v) for v in sys.argv[3].split(",")]
lib = capi.nompk()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
n3 = n ** 3
u = torch.rand(E * n3, dtype=torch.float64, device="cuda")
g = torch.rand(E * 6 * n3, dtype=torch.float64, device="cuda")
D = torch.rand(n * n, dtype=torch.float64, device="cuda")
w = torch.empty_like(u)
for v in variants:
    lib.nompk_ax_set_variant(v)
    for _ in range(2):
        capi.nompk_check(lib.nompk_ax_f64(n, E, u.data_ptr(), g.data_ptr(), D.data_ptr(), w.data_ptr(), 0, st))
    torch.cuda.synchronize()
    print("variant", v, flush=True)
