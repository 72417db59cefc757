"""The hand-written Ax kernel (libnomp_b200/csrc/kernels/ax.cu: ax_kernel, every element-group shape, with and without
the fused p.Ap) compiled for the HOST from its own text and executed with the cooperative emulator: one coroutine per
thread, __syncwarp / named barriers / shuffles / block tickets with their real semantics.  Only the four lines of inline
PTX are swapped (streaming load -> plain load, L2 prefetches -> nothing, bar.sync -> the emulator's named barrier) and
dynamic shared memory becomes a static array.  Checks the kernel's arithmetic and its shared-memory choreography
(mirrored lanes, stage barriers, element groups of several warps, partial last groups) against the oracle bit for bit
without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import ffi
from tests import cuda_emulation as emu

ROOT = Path(__file__).resolve().parent.parent
KERNELS = ROOT / "libnomp_b200" / "csrc" / "kernels"

# <n, elements per group, warps per group, groups per CTA>, as in ax.cu: Shape<N>
SHAPES = {6: (7, 4, 1), 8: (1, 1, 4), 10: (3, 5, 1), 12: (2, 5, 1)}


# what the emulated kernels instantiate: (geometric slabs in flight, fused p.Ap, two shared buffers, kPfMode, kPin, kEO: mask of the even-odd stages)
VARIANTS = {0: (2, False, False, 0, 0, 0),      # three buffers, wrapping prefetch window (n = 8 in production)
            1: (2, True, False, 0, 0, 0),       # ... with the fused p.Ap
            2: (2, False, True, 0, 0, 0),       # two buffers
            3: (1, False, True, 0, 0, 0),       # two buffers, one slab in flight
            4: (2, False, True, 4, 0, 0),       # local prefetch window, last-use demand loads
            5: (2, True, True, 3, 0, 0),        # two buffers, fused p.Ap, local window (production: n = 6, 12)
            6: (2, True, True, 5, 1, 0),        # ... first slabs prefetched too, ring fill pinned in front of S4
            7: (2, False, True, 5, 1, 0),       # the same without the dot product
            8: (2, True, True, 3, 1, 63),         # even-odd contractions (centro-antisymmetric D)
            9: (2, False, False, 0, 0, 63),       # ... on three buffers: n = 8 in production
            10: (3, True, False, 3, 0, 63),       # fused p.Ap, three buffers, three slabs in flight, local window: n = 10
            11: (2, False, True, 4, 0, 25),       # even-odd in S1, S5, S6 only
            12: (2, True, True, 4, 0, 29)}        # even-odd in S1, S2, S5, S6, fused p.Ap: n = 12 in production
DOT_VARIANTS = [v for v, spec in VARIANTS.items() if spec[1]]


def device_source():
    finish = (KERNELS / "nompk_gridreduce.cuh").read_text().replace('#include "nompk_common.cuh"', "").replace("#pragma once", "")
    finish = re.sub(r'asm volatile\("mov\.u64 %0, %globaltimer;" : "=l"\((\w+)\)\);', r"\1 = nomp_emu_now_ns();", finish)
    text = (KERNELS / "ax.cu").read_text()
    # the AxDotArgs structure, then everything the per-n units compile up to the host-side launcher
    s0, s1 = text.index("namespace nompk {\nnamespace {\n\n// What the fused p.Ap finish needs"), text.index("// One entry point per n:")
    a, b = text.index("#if NOMPK_AX_N != 0\n") + len("#if NOMPK_AX_N != 0\n"), text.index("template <int N, int G, int W, int GPC, int GA, int PF")
    body = text[s0:s1] + text[a:b] + "}  // namespace\n}  // namespace nompk\n"
    body, n = re.subn(r'__constant__ double nompk_ax_cD\[12 \* 12\];', "double nompk_ax_cD[12 * 12];", body)
    assert n == 1
    body, n = re.subn(r'__constant__ double nompk_ax_cEO\[4 \* 36\];', "double nompk_ax_cEO[4 * 36];", body)
    assert n == 1
    swaps = [(r'asm volatile\("ld\.global\.nc\.L1::no_allocate\.v2\.f64[^;]*;"[^;]*;', "r = *p;", 1),
             (r'asm volatile\("cp\.async\.bulk\.prefetch\.L2\.global[^;]*;"[^;]*;', ";", 1),
             (r'asm volatile\("prefetch\.global\.L2[^;]*;"[^;]*;', ";", 2),               # normal and evict_last priority
             (r'asm volatile\("createpolicy\.fractional\.L2::evict_first\.b64[^;]*;"[^;]*;', "policy = 0;", 1),
             (r'asm volatile\("ld\.global\.nc\.L2::cache_hint\.v2\.f64[^;]*;"[^;]*;', "r = *p; (void)policy;", 1),
             (r'asm volatile\("ld\.global\.v2\.f64[^;]*;"[^;]*;', "r = *p;", 1),               # the pinned (weak) load
             (r'asm volatile\("bar\.sync %0, %1;"[^;]*;', "nomp_emu_named_barrier(grp + 1, GL);", 1)]
    for pattern, repl, count in swaps:
        body, n = re.subn(pattern, repl, body)
        assert n == count, pattern
    body, n = re.subn(r"extern __shared__ double2 smem\[\];", "static double2 smem[1 << 16];", body)
    assert n == 1
    wrappers = ["#include <utility>\n#include <type_traits>\n#define __constant__ static\n"
                "static inline double __dadd_rn(double a, double b) { return a + b; }   // no contraction: -ffp-contract=off\n"
                "static inline double __dmul_rn(double a, double b) { return a * b; }\n", finish, body,
                ]
    for n_, (G, W, GPC) in SHAPES.items():
        for dot, (ga, with_dot, twobuf, pfmode, pin, eo) in VARIANTS.items():
            wrappers.append(
                f"static void ax{n_}_{dot}(const double *u, const double *g, const double *D, double *w, unsigned long long E, void *ws,"
                f" double *res, unsigned long long stride) {{\n"
                f"  if (threadIdx.x == 0) for (int i = 0; i < {n_ * n_}; i++) nompk::nompk_ax_cD[i] = D[i];   // the __constant__ copy\n"
                f"  if (threadIdx.x == 0) for (int idx = 0; idx < {n_ * n_}; idx++) {{   // what ax_stage_eo writes\n"
                f"    const int n = {n_}, h = n / 2, t = idx / (h * h), a = (idx / h) % h, l = idx % h;\n"
                f"    const double m1 = t >= 2 ? D[l * n + a] : D[a * n + l], m2 = t >= 2 ? D[(n - 1 - l) * n + a] : D[a * n + (n - 1 - l)];\n"
                f"    nompk::nompk_ax_cEO[t * 36 + a * h + l] = 0.5 * ((t & 1) ? m1 - m2 : m1 + m2);\n  }}\n"
                f"  __syncthreads();\n"
                f"  nompk::AxDotArgs d; d.workspace = ws; d.result = res; d.result_host = nullptr; d.host_seq = 0;\n"
                f"  nompk::ax_kernel<{n_}, {G}, {W}, {GPC}, {ga}, 4, false, 1, {'true' if with_dot else 'false'}, true,"
                f" {'true' if twobuf else 'false'}, false, {pfmode}, {pin}, {eo}>(u, g, w, E, d, stride, nompk::AxNoXpay());\n}}\n")
        # p <- r + beta p fused in front of the operator (always with the dot product); p is read and written in place
        wrappers.append(
            f"static void axx{n_}(double *p, const double *r, double beta, const double *beta_dev, const double *g, const double *D,"
            f" double *w, unsigned long long E, void *ws, double *res, unsigned long long stride) {{\n"
            f"  if (threadIdx.x == 0) for (int i = 0; i < {n_ * n_}; i++) nompk::nompk_ax_cD[i] = D[i];\n"
            f"  __syncthreads();\n"
            f"  nompk::AxDotArgs d; d.workspace = ws; d.result = res; d.result_host = nullptr; d.host_seq = 0;\n"
            f"  nompk::AxXpayArgs xp; xp.r = r; xp.p = p; xp.beta_dev = beta_dev; xp.beta = beta;\n"
            f"  nompk::ax_kernel<{n_}, {G}, {W}, {GPC}, {2 if n_ in (8, 12) else 3}, {4 if n_ in (8, 12) else 6}, false, 1, true, true, false, true>"
            f"(p, g, w, E, d, stride, xp);\n}}\n")
    return "".join(wrappers)


def run_ax(n, E, u, g, D, dot, blocks):
    G, W, GPC = SHAPES[n]
    src = device_source()
    ptr = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    w = np.full(E * n ** 3, np.nan)
    ws = np.zeros(548928 // 8 + 8, dtype=np.uint64)
    res = np.zeros(1)
    emu.emulate_cooperative(src, f"ax{n}_{dot}", (blocks, 1, 1), (GPC * W * 32, 1, 1),
                            ["const double *", "const double *", "const double *", "double *", "unsigned long long", "void *", "double *",
                             "unsigned long long"],
                            [ptr(u), ptr(g), ptr(D), ptr(w), C.c_ulonglong(E), ptr(ws), ptr(res), C.c_ulonglong(blocks * GPC * G)], instance=40)
    return w, res[0], ws


# every variant at n = 10 (three elements on five warps: mirrors, element strides, the row table); the other sizes run the
# three-buffer kernels, their own production shapes and the even-odd forms they use
CASES = [(n, v) for v in sorted(VARIANTS) for n in (6, 8, 10, 12)
         if n == 10 or v in {6: (0, 1, 3, 5), 8: (0, 1, 2, 9), 12: (0, 1, 4, 5, 8, 11, 12)}[n]]


@pytest.mark.parametrize("n,dot", CASES)
def test_ax_kernel_on_the_host(n, dot):
    """Exact-integer data: bitwise the oracle's w (and u . w) for element counts that are not multiples of the group
    size, with a persistent grid in which every CTA loops (two CTAs) and with one CTA per group."""
    G, W, GPC = SHAPES[n]
    per_cta = G * GPC
    for E, blocks in ((2 * per_cta + 1, 2), (per_cta + max(1, per_cta // 2), 2), (1, 1)):
        u = ffi.fill_int_f64(E * n ** 3, 5 + n, -4, 4)
        g = ffi.fill_int_f64(E * 6 * n ** 3, 6 + n, 0, 3)
        D = ffi.fill_int_f64(n * n, 7 + n, -2, 2)
        if VARIANTS[dot][5]:      # even-odd: D[a][l] == -D[n-1-a][n-1-l]; its halves are multiples of 1/2 -- still exact
            D = np.ascontiguousarray((D.reshape(n, n) - D.reshape(n, n)[::-1, ::-1]).ravel())
        w, pap, ws = run_ax(n, E, u, g, D, dot, blocks)
        want = ffi.ax(n, u, g, D)
        assert np.array_equal(w, want), (n, E, blocks, int((w != want).sum()))
        if dot in DOT_VARIANTS:
            assert pap == float(u @ want) and not ws[: (64 + 4 * 2048) // 8].any()


@pytest.mark.parametrize("n", [6, 8, 10, 12])
def test_ax_kernel_text_against_the_analytic_known_answer(n):
    """The kernel's own text on the emulator against tests/ax_closed_form.py (harmonic polynomial on a sheared element:
    boundary fluxes, zero inside) with the real GLL matrix -- real-valued data, FMA contraction and all; the fused p.Ap
    equals u . (A u)."""
    from tests.ax_closed_form import sheared_element
    G, W, GPC = SHAPES[n]
    Dm, x = ffi.gll_derivative(n)
    u1, g1, want1, interior1 = sheared_element(n, x)
    E = G * GPC + 1                                   # one full group and a partial one
    u, g, want, interior = np.tile(u1, E), np.tile(g1, E), np.tile(want1, E), np.tile(interior1, E)
    D = np.ascontiguousarray(Dm.ravel())
    scale = np.abs(want).max()
    for dot in (0, 1):
        w, pap, _ = run_ax(n, E, u, g, D, dot, 2)
        assert np.abs(w - want).max() <= 1e-11 * scale, (n, dot)
        assert np.abs(w[interior]).max() <= 1e-11 * scale
        if dot:
            assert abs(pap - float(u @ want)) <= 1e-10 * abs(float(u @ want))


@pytest.mark.parametrize("n", [6, 8, 10, 12])
def test_ax_with_the_direction_update_fused_in_front(n):
    """kXpay: p <- r + beta p, w <- A p, p.w in one kernel -- bitwise the stand-alone sequence (oracle map XPAY, oracle
    Ax, exact dot product) on exact-integer data, with beta as a kernel parameter and read from device memory; p is
    updated in place, partial last group included."""
    G, W, GPC = SHAPES[n]
    per_cta = G * GPC
    src = device_source()
    ptr = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    for E, blocks, beta, from_dev in ((2 * per_cta + 1, 2, 2.0, False), (per_cta + 1, 1, -3.0, True)):
        p0 = ffi.fill_int_f64(E * n ** 3, 15 + n, -2, 2)
        r = ffi.fill_int_f64(E * n ** 3, 16 + n, -2, 2)
        g = ffi.fill_int_f64(E * 6 * n ** 3, 17 + n, 0, 3)
        D = ffi.fill_int_f64(n * n, 18 + n, -2, 2)
        want_p = r + beta * p0
        want_w = ffi.ax(n, want_p, g, D)
        p, w = p0.copy(), np.full(E * n ** 3, np.nan)
        ws = np.zeros(548928 // 8 + 8, dtype=np.uint64)
        res = np.zeros(1)
        beta_dev = np.array([beta])
        emu.emulate_cooperative(src, f"axx{n}", (blocks, 1, 1), (GPC * W * 32, 1, 1),
                                ["double *", "const double *", "double", "const double *", "const double *", "const double *", "double *",
                                 "unsigned long long", "void *", "double *", "unsigned long long"],
                                [ptr(p), ptr(r), C.c_double(0.0 if from_dev else beta), ptr(beta_dev) if from_dev else C.c_void_p(0), ptr(g), ptr(D),
                                 ptr(w), C.c_ulonglong(E), ptr(ws), ptr(res), C.c_ulonglong(blocks * per_cta)], instance=41)
        assert np.array_equal(p, want_p), (n, E)
        assert np.array_equal(w, want_w), (n, E, int((w != want_w).sum()))
        assert res[0] == float(want_p @ want_w)
