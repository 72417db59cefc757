#!/usr/bin/env python3
"""Generates tests/golden/*.json -- the pins of the CPU oracle.

Two sources, neither of them the oracle itself:
  1. the closed-form known answers the REFERENCE's own tests assert for this path (reference
     tests/nomp-api-200-impl.h:31-185, tests/nomp-api-205-impl.h:30-135, tests/nomp-api-500-impl.h:17-248,
     tests/nomp-api-600-impl.h:33-52; listed in SURVEY.md 8c).  The reference cannot be run here (SymEngine, loopy,
     libclang, pocl missing), so its golden vectors are these formulas, re-evaluated below for the same n and types;
  2. for Ax and the gather-scatter, which the reference does not contain ("parity unpinned" by the reference),
     independent numpy restatements (einsum over the definition in include/nompk.h; bincount / ufunc.at over the
     global ids) on exact-integer data, plus the analytic identities the tests check.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
TYPES = ["int32", "int64", "uint32", "uint64", "float64", "float32"]   # the reference's six-type matrix


def map_cases():
    cases = []
    for n in (10, 50, 70):
        i = np.arange(n)
        # (name, kernel body, a0, b0, c0, expected a) -- formulas of the reference tests
        cases += [
            dict(name="add", body="a[i] += b[i];", n=n, a=(n - i).tolist(), b=i.tolist(), want=[n] * n),
            dict(name="sub", body="a[i] -= b[i] + 1;", n=n, a=(n + i).tolist(), b=i.tolist(), want=[n - 1] * n),
            dict(name="mul_sum", body="a[i] *= b[i] + 1;", n=n, a=(n - i).tolist(), b=i.tolist(), want=((n - i) * (i + 1)).tolist()),
            dict(name="mul", body="a[i] = a[i] * b[i];", n=n, a=(n - i).tolist(), b=i.tolist(), want=((n - i) * i).tolist()),
            dict(name="square", body="a[i] = a[i] * a[i] + b[i] * b[i];", n=n, a=(n - i).tolist(), b=i.tolist(),
                 want=((n - i) ** 2 + i ** 2).tolist()),
            dict(name="linear", body="a[i] = 2 * b[i] + 1;", n=n, a=[0] * n, b=i.tolist(), want=(2 * i + 1).tolist()),
            dict(name="add3", body="a[i] = a[i] + b[i] + c[i];", n=n, a=(n - i).tolist(), b=i.tolist(), c=i.tolist(),
                 want=(n + i).tolist()),
            dict(name="mul3", body="a[i] = a[i] * b[i] * c[i];", n=n, a=(n - i).tolist(), b=i.tolist(), c=[2] * n,
                 want=((n - i) * i * 2).tolist()),
            dict(name="linear3", body="a[i] = a[i] + 3 * b[i] + 2 * c[i];", n=n, a=(n - i).tolist(), b=i.tolist(),
                 c=i.tolist(), want=(n + 4 * i).tolist()),
        ]
    return cases


def reduce_cases():
    cases = []
    for N in (10, 50):
        i = np.arange(N)
        cases += [
            dict(name="sum_const", body="s[0] += 1;", n=N, a=[0] * N, want=N),
            dict(name="sum_index", body="s[0] += i;", n=N, a=[0] * N, want=N * (N - 1) // 2),
            dict(name="sum_array", body="s[0] += a[i];", n=N, a=i.tolist(), want=N * (N - 1) // 2),
            dict(name="dot", body="s[0] += a[i] * b[i];", n=N, a=i.tolist(), b=i.tolist(), want=N * (2 * N - 1) * (N - 1) // 6),
        ]
        for k in (1, 2, 3, 4):
            cases.append(dict(name=f"sum_scaled_{k}", body="s[0] += a[i];", n=N, a=(k * i).tolist(), want=k * N * (N - 1) // 2))
    return cases


def ax_numpy(n, E, u, g, D):
    """Independent restatement of the Ax definition with einsum (fp64; exact for the integer data used here)."""
    u = u.reshape(E, n, n, n)            # [e][k][j][i]
    g = g.reshape(E, 6, n, n, n)
    ur = np.einsum("il,ekjl->ekji", D, u)
    us = np.einsum("jl,ekli->ekji", D, u)
    ut = np.einsum("kl,elji->ekji", D, u)
    wr = g[:, 0] * ur + g[:, 1] * us + g[:, 2] * ut
    ws = g[:, 1] * ur + g[:, 3] * us + g[:, 4] * ut
    wt = g[:, 2] * ur + g[:, 4] * us + g[:, 5] * ut
    w = np.einsum("li,ekjl->ekji", D, wr) + np.einsum("lj,ekli->ekji", D, ws) + np.einsum("lk,elji->ekji", D, wt)
    return w.reshape(-1)


def ax_cases():
    cases = []
    for n, E, seed in ((4, 2, 1), (8, 3, 2), (10, 2, 3)):
        rng = np.random.default_rng(seed)
        u = rng.integers(-4, 5, E * n ** 3).astype(np.float64)
        g = rng.integers(0, 4, E * 6 * n ** 3).astype(np.float64)
        D = rng.integers(-2, 3, (n, n)).astype(np.float64)
        w = ax_numpy(n, E, u, g, D)
        cases.append(dict(n=n, E=E, u=u.tolist(), g=g.tolist(), D=D.ravel().tolist(), w=w.tolist()))
    return cases


def gs_numpy(ids, v, op):
    """Independent restatement of the gather-scatter: combine per global id with ufunc.at, then read back."""
    ids = np.asarray(ids)
    out = v.copy()
    live = ids > 0
    table_size = int(ids.max()) + 1
    if op == "+":
        table = np.zeros(table_size, dtype=v.dtype)
        np.add.at(table, ids[live], v[live])
    elif op == "*":
        table = np.ones(table_size, dtype=v.dtype)
        np.multiply.at(table, ids[live], v[live])
    elif op == "min":
        table = np.full(table_size, v.max(), dtype=v.dtype)
        np.minimum.at(table, ids[live], v[live])
    else:
        table = np.full(table_size, v.min(), dtype=v.dtype)
        np.maximum.at(table, ids[live], v[live])
    out[live] = table[ids[live]]
    return out


def gs_cases():
    """A 3 x 2 x 2 box of elements with 3 points per direction (lexicographic global numbering written out here, not
    taken from the oracle's helper) and a random numbering with non-participating ids; integer-valued data."""
    cases = []
    n, ex, ey, ez = 3, 3, 2, 2
    N = n - 1
    px, py = N * ex + 1, N * ey + 1
    ids = []
    for e in range(ex * ey * ez):
        e_x, e_y, e_z = e % ex, (e // ex) % ey, e // (ex * ey)
        for k in range(n):
            for j in range(n):
                for i in range(n):
                    ids.append(1 + (e_x * N + i) + px * ((e_y * N + j) + py * (e_z * N + k)))
    rng = np.random.default_rng(11)
    numberings = [("box", np.array(ids, dtype=np.int64)), ("random", rng.integers(-1, 40, 300).astype(np.int64))]
    for name, idv in numberings:
        for op in ("+", "*", "min", "max"):
            for dtype in ("float64", "int64", "float32", "int32"):
                v = rng.integers(1, 4, idv.size) if op == "*" else rng.integers(-9, 10, idv.size)
                v = v.astype(dtype)
                if op == "*" and name == "random":
                    v = np.where(rng.random(idv.size) < 0.7, 1, v).astype(dtype)   # keep products small and exact
                cases.append(dict(name=name, op=op, dtype=dtype, ids=idv.tolist(), v=v.tolist(),
                                  want=gs_numpy(idv, v, op).tolist()))
    return cases


def main():
    (HERE / "gs_cases.json").write_text(json.dumps(dict(cases=gs_cases())))
    (HERE / "map_cases.json").write_text(json.dumps(dict(types=TYPES, cases=map_cases())))
    (HERE / "reduce_cases.json").write_text(json.dumps(dict(types=TYPES, cases=reduce_cases())))
    (HERE / "ax_cases.json").write_text(json.dumps(dict(cases=ax_cases())))
    # GLL nodes for N = 7 and N = 9 (Abramowitz & Stegun table 25.6, 10 digits) pin the derivative-matrix generator
    (HERE / "gll_nodes.json").write_text(json.dumps({
        "8": [-1.0, -0.8717401485, -0.5917001814, -0.2092992179, 0.2092992179, 0.5917001814, 0.8717401485, 1.0],
        "10": [-1.0, -0.9195339082, -0.7387738651, -0.4779249498, -0.1652789577, 0.1652789577, 0.4779249498,
               0.7387738651, 0.9195339082, 1.0]}))
    for f in sorted(HERE.glob("*.json")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
