#!/usr/bin/env python3
"""Offline model of the Ax kernel's shared-memory accesses (libnomp_b200/csrc/kernels/ax.cu), reconciled with ncu.

Every access is a 128-bit LDS/STS of one 16-byte chunk.  The hardware serves such a request 8 lanes per wavefront
(quarter warps of the launching warp); a wavefront is conflict-free when the 8 chunk addresses are distinct mod 8
(8 x 16 B = 128 B = all 32 banks) or equal.  `L1 Wavefronts Shared` / `Instructions Executed` of the ncu source page is
4.0 for a conflict-free instruction; round 2's captures showed 7.0 (n = 10) and 10.0 (n = 12) on the i-line stages,
which this model reproduces for the round-1 layout ("old") -- round 1's version of this script looked at one element in
isolation and at rows (2t, 2t+1) only from lane 0 of an element, and missed that a group packs several elements onto
consecutive lanes (n = 10: 3 x 50 lanes on 5 warps).

Layout: chunk address of (k, j, p) of element el in a group = el * elem_stride + k * SK + j * NP + (p ^ swz(k, j)).
Patterns (lane L of the group = el * T + tt):
  k-column : (j, p) = (tt // NP, tt % NP), one instruction per k
  j-line   : (k, p) = (tt // NP, tt % NP), one instruction per l (the j index)
  i-line   : rows A(tt), B(tt), one instruction per chunk c and row
"""

SHAPES = {6: (7, 4), 8: (1, 1), 10: (3, 5), 12: (2, 5)}          # n -> (elements per group, warps per group)
SK = {6: 19, 8: 36, 10: 53, 12: 78}


def row_table(n):
    """The constexpr RowTable of ax.cu."""
    T = n * n // 2
    used = [False] * (n * n)
    rows = [0] * (2 * T)
    for which in range(2):
        for tt in range(T):
            want = (tt + which) % 8
            pick = -1
            for p in range(8):
                cls = (want + p) % 8
                for r in range(n * n):
                    if not used[r] and (r // n + r % n) % 8 == cls:
                        pick = r
                        break
                if pick >= 0:
                    break
            used[pick] = True
            rows[2 * tt + which] = pick
    return rows


def model(n, new=True, bufs=3):
    NP, T, sk = n // 2, n * n // 2, SK[n]
    G, W = SHAPES[n]
    chunks = n * sk
    stride = bufs * chunks
    if new and n != 8:
        stride += (T - stride) % 8

    def swz(k, j):
        if n == 8:
            return (j >> 1) & 3
        if n == 12 and new:
            return ((k + j) >> 2) & 1
        return 0

    def at(el, k, j, p):
        return el * stride + k * sk + j * NP + (p ^ swz(k, j))

    tab = row_table(n) if (new and n != 8) else [2 * (i // 2) + i % 2 for i in range(2 * T)]
    lanes = [(L // T, L % T) if L < G * T else ((G * T - 1) // T, (G * T - 1) % T) for L in range(32 * W)]   # surplus lanes mirror
    out = {}
    for name in ("kcol", "jline", "iline"):
        wavefronts = instr = 0
        for q0 in range(0, 32 * W, 8):
            quarter = lanes[q0:q0 + 8]
            accesses = []
            if name == "kcol":
                accesses = [[at(el, k, tt // NP, tt % NP) for el, tt in quarter] for k in range(n)]
            elif name == "jline":
                accesses = [[at(el, tt // NP, l, tt % NP) for el, tt in quarter] for l in range(n)]
            else:
                for which in range(2):
                    for c in range(NP):
                        accesses.append([at(el, tab[2 * tt + which] // n, tab[2 * tt + which] % n, c) for el, tt in quarter])
            for a in accesses:
                cols = {}
                for x in set(a):
                    cols[x % 8] = cols.get(x % 8, 0) + 1
                wavefronts += max(cols.values())
                instr += 1
        out[name] = round(4.0 * wavefronts / instr, 2)      # wavefronts per warp instruction (4 quarters)
    return out


if __name__ == "__main__":
    for n in (6, 8, 10, 12):
        for bufs in (3, 2):
            print(f"n={n:2d} bufs={bufs}  old {model(n, False, bufs)}   new {model(n, True, bufs)}")
