#!/usr/bin/env python3
"""Offline check of the Ax kernel's shared-memory layout (libnomp_b200/csrc/kernels/ax.cu).

Every smem access in the Ax kernel is a 128-bit (16 B "chunk") LDS/STS. The hardware serves such a
request 8 lanes per wavefront; the 8 lanes are conflict-free when their chunk addresses are distinct
mod 8 (8 chunk columns x 16 B = 128 B = 32 banks) or identical. This script enumerates the three access
patterns of the kernel for a candidate layout and reports the worst-case wavefront multiplicity.

Layout: chunk address A(k, j, p) = k*SK + j*NP + (p ^ swz(j)), NP = n/2 chunks per row.
Work item t in [0, n*n/2):
  k-column pairs : j = t // NP, p = t % NP, loop over k
  j-line pairs   : k = t // NP, p = t % NP, loop over l (the j index)
  i-line pairs   : rows r = 2t, 2t+1 (r = k*n + j), loop over chunk c
"""
import sys


def layout(n, SK, swz):
    NP = n // 2

    def A(k, j, p):
        return k * SK + j * NP + (p ^ swz(j))

    return A


def worst(groups):
    w = 1
    for g in groups:
        cols = {}
        for a in g:
            cols.setdefault(a % 8, set()).add(a)
        w = max(w, max(len(v) for v in cols.values()))
    return w


def check(n, SK, swz):
    NP = n // 2
    T = n * n // 2
    A = layout(n, SK, swz)
    res = {}
    for name in ("kcol", "jline", "iline"):
        groups = []
        for base in range(0, T, 8):
            lanes = [t for t in range(base, min(base + 8, T))]
            if name == "kcol":
                for k in range(n):
                    groups.append([A(k, t // NP, t % NP) for t in lanes])
            elif name == "jline":
                for l in range(n):
                    groups.append([A(t // NP, l, t % NP) for t in lanes])
            else:
                for which in (0, 1):
                    for c in range(NP):
                        g = []
                        for t in lanes:
                            r = 2 * t + which
                            g.append(A(r // n, r % n, c))
                        groups.append(g)
        res[name] = worst(groups)
    return res


if __name__ == "__main__":
    print("n=8 natural      ", check(8, 32, lambda j: 0))
    print("n=8 swz, SK=36   ", check(8, 36, lambda j: (j >> 1) & 3))
    for sk in range(50, 60):
        print("n=10 SK=%d       " % sk, check(10, sk, lambda j: 0))
    for n in (4, 6, 12):
        NP = n // 2
        for sk in range(n * NP, n * NP + 9):
            print("n=%d SK=%d" % (n, sk), check(n, sk, lambda j: 0))
