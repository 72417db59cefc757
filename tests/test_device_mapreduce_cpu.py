"""The map and reduce kernels of libnompk (csrc/kernels/map.cu, reduce.cu) compiled for the HOST from their own text and
executed with the cooperative emulator (tests/cuda_emulation.py), against the CPU oracle: every map operator, full and
partial tiles, the scalar tail and the scalar kernel for misaligned operands; sum / product / min / max and dot
products through the vector and the scalar reduction kernel with the grid-wide finish behind them.  Together with
tests/test_device_finish_cpu.py, test_device_gs_cpu.py and test_device_ax_cpu.py every hand-written kernel of the
library is executed in the CPU suite."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from libnomp_b200 import capi
from oracle import ffi
from tests import cuda_emulation as emu

ROOT = Path(__file__).resolve().parent.parent
KERNELS = ROOT / "libnomp_b200" / "csrc" / "kernels"

PRELUDE = r"""
#include <cfloat>
#include <climits>
#include <type_traits>
#define __host__
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
enum { NOMPK_MAP_ADD, NOMPK_MAP_SUB, NOMPK_MAP_MUL, NOMPK_MAP_AXPY, NOMPK_MAP_XPAY, NOMPK_MAP_AXPBY, NOMPK_MAP_SCALE, NOMPK_MAP_COPY,
       NOMPK_MAP_FILL, NOMPK_MAP_ADD3 };
enum { NOMPK_RED_SUM, NOMPK_RED_PROD, NOMPK_RED_MIN, NOMPK_RED_MAX };
namespace nompk {
"""


def common_helpers():
    text = (KERNELS / "nompk_common.cuh").read_text()
    a = text.index("template <typename T> __device__ __forceinline__ T op_add(T a, T b)")
    b = text.index("}  // namespace nompk")
    return text[a:b] + "}\n"


def map_source():
    text = (KERNELS / "map.cu").read_text()
    a, b = text.index("namespace nompk {\nnamespace {"), text.index("template <int OP, typename T>\nint launch_map(")
    body = text[a:b] + "}\n}\n"
    ops = "ADD SUB MUL AXPY XPAY AXPBY SCALE COPY FILL ADD3".split()
    wrappers = []
    for T, tag in (("double", "f64"), ("int", "i32")):
        for i, op in enumerate(ops):
            wrappers.append(f"static void vec_{tag}_{i}({T} *y, const {T} *x, const {T} *z, {T} a, {T} b, unsigned long long nvec, unsigned long long n)"
                            f" {{ nompk::map_vec_kernel<NOMPK_MAP_{op}, {T}, 1, 256>(y, x, z, a, b, nvec, n); }}\n")
            wrappers.append(f"static void sca_{tag}_{i}({T} *y, const {T} *x, const {T} *z, {T} a, {T} b, unsigned long long n)"
                            f" {{ nompk::map_scalar_kernel<NOMPK_MAP_{op}, {T}>(y, x, z, a, b, n); }}\n")
    return PRELUDE + common_helpers() + body + "".join(wrappers)


def reduce_source():
    finish = (KERNELS / "nompk_gridreduce.cuh").read_text().replace('#include "nompk_common.cuh"', "").replace("#pragma once", "")
    finish = re.sub(r'asm volatile\("mov\.u64 %0, %globaltimer;" : "=l"\((\w+)\)\);', r"\1 = nomp_emu_now_ns();", finish)
    text = (KERNELS / "reduce.cu").read_text()
    a, b = text.index("namespace nompk {\nnamespace {"), text.index("template <int OP, typename T>\nint launch_reduce(")
    body = text[a:b] + "}\n}\n"
    wrappers = []
    for T, tag in (("double", "f64"), ("long long", "i64"), ("float", "f32")):
        for op, name in enumerate(("SUM", "PROD", "MIN", "MAX")):
            for dot in (0, 1):
                for vec in (0, 1):
                    wrappers.append(
                        f"static void red_{tag}_{op}_{dot}_{vec}(const {T} *x, const {T} *y, unsigned long long n, void *ws, {T} *res, {T} *pub,"
                        f" unsigned long long seq) {{ nompk::reduce_kernel<NOMPK_RED_{name}, {'true' if dot else 'false'},"
                        f" {'true' if vec else 'false'}, {T}>(x, y, n, ws, res, pub, seq, nompk::PeerExchange()); }}\n")
    return PRELUDE + common_helpers() + finish + body + "".join(wrappers)


def ptr(a):
    return C.c_void_p(a.ctypes.data if a is not None else 0)


@pytest.mark.parametrize("op", range(10))
@pytest.mark.parametrize("tag,dtype", [("f64", ffi.F64), ("i32", ffi.I32)])
def test_map_kernels_on_the_host(op, tag, dtype):
    npdt = ffi.NP_DTYPES[dtype]
    T = {"f64": "double", "i32": "int"}[tag]
    ctype = {"f64": C.c_double, "i32": C.c_int}[tag]
    lanes = 16 // np.dtype(npdt).itemsize
    rng = np.random.default_rng(op)
    alpha, beta = npdt(3), npdt(-2)
    for n, blocks in ((1, 1), (lanes * 256 * 2 + 3, 2), (lanes * 300 + 1, 3)):
        y = rng.integers(-50, 50, n).astype(npdt)
        x = rng.integers(-50, 50, n).astype(npdt)
        z = rng.integers(-50, 50, n).astype(npdt)
        want = ffi.map_(op, dtype, y.copy(), x, z, alpha, beta)
        got = y.copy()
        emu.emulate_cooperative(map_source(), f"vec_{tag}_{op}", (blocks, 1, 1), (256, 1, 1),
                                [f"{T} *", f"const {T} *", f"const {T} *", T, T, "unsigned long long", "unsigned long long"],
                                [ptr(got), ptr(x), ptr(z), ctype(alpha), ctype(beta), C.c_ulonglong(n // lanes), C.c_ulonglong(n)],
                                instance=50)
        assert np.array_equal(got, want), (op, tag, n, blocks)
        got = y.copy()
        emu.emulate_cooperative(map_source(), f"sca_{tag}_{op}", (blocks, 1, 1), (256, 1, 1),
                                [f"{T} *", f"const {T} *", f"const {T} *", T, T, "unsigned long long"],
                                [ptr(got), ptr(x), ptr(z), ctype(alpha), ctype(beta), C.c_ulonglong(n)], instance=50)
        assert np.array_equal(got, want), (op, tag, n, "scalar")


@pytest.mark.parametrize("tag,dtype", [("f64", ffi.F64), ("i64", ffi.I64), ("f32", ffi.F32)])
@pytest.mark.parametrize("op", range(4))
def test_reduce_kernels_on_the_host(tag, dtype, op):
    npdt = ffi.NP_DTYPES[dtype]
    T = {"f64": "double", "i64": "long long", "f32": "float"}[tag]
    lanes = 16 // np.dtype(npdt).itemsize
    rng = np.random.default_rng(10 + op)
    for n, blocks in ((0, 1), (1, 1), (lanes * 256 * 4 * 2 + 5, 3), (9001, 7)):     # > 2048 CTAs: test_device_finish_cpu.py
        if op == capi.RED_PROD:
            x = rng.integers(1, 3, min(n, 40)).astype(npdt)                  # products of 1s and 2s stay exact
            y = rng.integers(1, 3, x.size).astype(npdt)
        else:
            x = rng.integers(-9, 10, n).astype(npdt)
            y = rng.integers(-9, 10, n).astype(npdt)
        for dot in (0, 1):
            want = ffi.reduce_(op, dtype, x, y if dot else None)
            for vec in (0, 1):
                ws = np.zeros(548928 // 8 + 8, dtype=np.uint64)
                res, pub = np.zeros(1, dtype=npdt), np.zeros(24, dtype=np.uint8)
                emu.emulate_cooperative(reduce_source(), f"red_{tag}_{op}_{dot}_{vec}", (blocks, 1, 1), (256, 1, 1),
                                        [f"const {T} *", f"const {T} *", "unsigned long long", "void *", f"{T} *", f"{T} *",
                                         "unsigned long long"],
                                        [ptr(x), ptr(y if dot else None), C.c_ulonglong(x.size), ptr(ws), ptr(res), ptr(pub),
                                         C.c_ulonglong(77)], instance=51)
                assert res[0] == want == pub[:res.itemsize].view(npdt)[0], (tag, op, x.size, blocks, dot, vec)
                assert int(pub[8:16].view(np.uint64)[0]) == 77 and not ws[: (64 + 4 * 2048) // 8].any()
