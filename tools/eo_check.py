#!/usr/bin/env python3
"""Even-odd Ax variants (54, 55) against the extended-precision oracle with the real GLL matrix (run on the GPU box)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from libnomp_b200 import capi
from oracle import ffi
lib = capi.nompk()
import ctypes as C
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for n in (6, 8, 10, 12):
    E = 257
    D = np.ascontiguousarray(ffi.gll_derivative(n)[0].ravel())
    u = ffi.fill_uniform_f64(E * n ** 3, 1234, 0.5, 1.5)
    g = ffi.fill_uniform_f64(E * 6 * n ** 3, 99, 0.5, 1.5)
    ref = ffi.ax(n, u, g, D, "extended")
    tu, tg, tD = (torch.from_numpy(a).cuda() for a in (u, g, D))
    for v in (0, 54, 55):
        tw = torch.full_like(tu, float("nan"))
        lib.nompk_ax_set_variant(v)
        capi.nompk_check(lib.nompk_ax_f64(n, E, tu.data_ptr(), tg.data_ptr(), tD.data_ptr(), tw.data_ptr(), 0, st))
        torch.cuda.synchronize()
        err = np.abs(tw.cpu().numpy() - ref).max() / np.abs(ref).max()
        print(f"n={n} variant={v} rel_err={err:.3e}")
    lib.nompk_ax_set_variant(0)
