#!/usr/bin/env python3
"""Compare the output of Ax variants with variant 0 on the GPU (bitwise; exact-integer data).
usage: ax_variant_check.py <variant> [<variant> ...]"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from libnomp_b200 import capi  # noqa: E402

lib = capi.nompk()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ok = True
for n in (6, 8, 10, 12):
    E = 1237
    n3 = n ** 3
    u = torch.randint(-4, 5, (E * n3,), device="cuda").double()
    g = torch.randint(0, 4, (E * 6 * n3,), device="cuda").double()
    D = torch.randint(-2, 3, (n * n,), device="cuda").double()
    outs = {}
    for v in [0] + [int(a) for a in sys.argv[1:]]:
        w = torch.full_like(u, float("nan"))
        lib.nompk_ax_set_variant(v)
        capi.nompk_check(lib.nompk_ax_f64(n, E, u.data_ptr(), g.data_ptr(), D.data_ptr(), w.data_ptr(), 0, st))
        torch.cuda.synchronize()
        outs[v] = w
    lib.nompk_ax_set_variant(0)
    for v, w in outs.items():
        same = bool(torch.equal(w, outs[0]))
        ok &= same
        print(f"n={n} variant {v}: {'identical to variant 0' if same else 'DIFFERENT'}")
sys.exit(0 if ok else 1)
