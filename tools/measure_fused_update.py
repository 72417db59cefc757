#!/usr/bin/env python3
"""Bandwidth of generated (NVRTC skeleton) kernels through nomp_run: fused CG update, generic maps."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from libnomp_b200 import capi

capi.check(capi.init(backend="cuda", device=0, verbose=1))
lib = capi.nomp()
stream = torch.cuda.ExternalStream(lib.nomp_b200_stream())
P, I, F = capi.NOMP_PTR, capi.NOMP_INT, capi.NOMP_FLOAT
n = 1 << 27
keys = [np.empty(n) for _ in range(4)]
for k in keys:
    capi.check(capi.update(k.ctypes.data, 0, n, 8, capi.NOMP_ALLOC))
x, r, p, w = (k.ctypes.data for k in keys)
err, fill = capi.jit("void f(double *a, int n) { for (int i = 0; i < n; i++) a[i] = 1.0 + (i & 7) * 0.125; }", capi.clauses(),
                     [("a", 8, P), ("n", 4, I)])
capi.check(err)
for k in (x, r, p, w):
    capi.check(capi.run(fill, k, C.c_int(n)))


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    capi.check(lib.nomp_sync())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    capi.check(lib.nomp_sync())
    return e0.elapsed_time(e1) / reps


cases = [
    ("fused_cg_update (48 B)", "void upd(double *x, double *r, const double *p, const double *w, double alpha, int N, double *rr) { for (int i = 0; i < N; i++) { x[i] += alpha * p[i]; r[i] -= alpha * w[i]; rr[0] += r[i] * r[i]; } }",
     capi.clauses(("reduce", "rr", "+")), [("x", 8, P), ("r", 8, P), ("p", 8, P), ("w", 8, P), ("alpha", 8, F), ("N", 4, I), ("rr", 8, F)], 48,
     lambda kid, s: capi.run(kid, x, r, p, w, C.c_double(1e-9), C.c_int(n), s)),
    ("sum of squares skeleton (8 B)", "void ss(const double *p, int N, double *rr) { for (int i = 0; i < N; i++) rr[0] += p[i] * p[i] + 1; }",
     capi.clauses(("reduce", "rr", "+")), [("p", 8, P), ("N", 4, I), ("rr", 8, F)], 8, lambda kid, s: capi.run(kid, p, C.c_int(n), s)),
    ("generic map skeleton a=a*b+i (24 B)", "void m(double *x, const double *p, int N) { for (int i = 0; i < N; i++) x[i] = x[i] * p[i] + i; }",
     capi.clauses(), [("x", 8, P), ("p", 8, P), ("N", 4, I)], 24, lambda kid, s: capi.run(kid, x, p, C.c_int(n))),
    ("native xpay (24 B)", "void m(double *x, const double *p, double b, int N) { for (int i = 0; i < N; i++) x[i] = p[i] + b * x[i]; }",
     capi.clauses(), [("x", 8, P), ("p", 8, P), ("b", 8, F), ("N", 4, I)], 24, lambda kid, s: capi.run(kid, x, p, C.c_double(0.5), C.c_int(n))),
]
for name, src, cl, args, bpe, call in cases:
    err, kid = capi.jit(src, cl, args)
    capi.check(err)
    s = C.c_double()
    ms = timed(lambda: capi.check(call(kid, s)))
    print(f"{name:40s} {lib.nomp_b200_prog_info(kid).decode()[:40]:42s} {ms:.4f} ms  {n * bpe / ms / 1e6:.0f} GB/s")
lib.nomp_finalize_excluding_interpreter()
