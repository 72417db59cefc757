"""Parity of the hand-written sm_100a kernels (libnompk.so, called through the C ABI of include/nompk.h)
against the CPU oracle (oracle/nomp_oracle.c) on the same seeded inputs.

Bars: bit-exact for maps (all dtypes), integer reductions and exact-data (Set X) fp64 reductions / Ax;
1e-12 relative against the compensated / extended-precision oracle for random fp64 data (Set R).
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from libnomp_b200 import capi  # noqa: E402
from oracle import ffi  # noqa: E402

NP = ffi.NP_DTYPES
TORCH_DT = {capi.I32: torch.int32, capi.U32: torch.int32, capi.I64: torch.int64, capi.U64: torch.int64,
            capi.F32: torch.float32, capi.F64: torch.float64}


def dev(a: np.ndarray):
    """numpy -> device tensor holding the same bytes (unsigned types travel as their signed twin)."""
    signed = {np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64}.get(a.dtype)
    t = torch.from_numpy(a.view(signed) if signed else a).cuda()
    return t


def host(t, npdt):
    return t.cpu().numpy().view(npdt)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rand_array(dtype, n, seed):
    rng = np.random.default_rng(seed)
    npdt = NP[dtype]
    if dtype in (capi.F32, capi.F64):
        return rng.uniform(0.5, 1.5, n).astype(npdt)
    info = np.iinfo(npdt)
    return rng.integers(info.min, info.max, n, dtype=npdt, endpoint=True)


MAP_OPS = [capi.MAP_ADD, capi.MAP_SUB, capi.MAP_MUL, capi.MAP_AXPY, capi.MAP_XPAY, capi.MAP_AXPBY, capi.MAP_SCALE,
           capi.MAP_COPY, capi.MAP_FILL, capi.MAP_ADD3]


@pytest.mark.parametrize("dtype", [capi.I32, capi.U32, capi.I64, capi.U64, capi.F32, capi.F64])
@pytest.mark.parametrize("n", [0, 1, 10, 50, 70, 1023, 4099, (1 << 20) + 3])
def test_map_bit_exact(dtype, n):
    lib = capi.nompk()
    npdt = NP[dtype]
    for op in MAP_OPS:
        y = rand_array(dtype, n, 1 + op)
        x = rand_array(dtype, n, 100 + op)
        z = rand_array(dtype, n, 200 + op)
        alpha = npdt(3) if dtype not in (capi.F32, capi.F64) else npdt(0.37)
        beta = npdt(5) if dtype not in (capi.F32, capi.F64) else npdt(-1.25)
        want = ffi.map_(op, dtype, y.copy(), x, z, alpha, beta)
        ty, tx, tz = dev(y), dev(x), dev(z)
        a, b = np.array([alpha], npdt), np.array([beta], npdt)
        rc = lib.nompk_map(op, dtype, n, ty.data_ptr(), tx.data_ptr(), tz.data_ptr(), a.ctypes.data, b.ctypes.data,
                           stream())
        capi.nompk_check(rc, "nompk_map")
        got = host(ty, npdt)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), f"op {op} dtype {dtype} n {n}"


def test_map_misaligned_operands():
    """Sub-range mappings give pointers that are only element-aligned: the scalar path must agree too."""
    lib = capi.nompk()
    n = 5000
    y = rand_array(capi.F64, n + 1, 3)
    x = rand_array(capi.F64, n + 1, 4)
    want = ffi.map_(capi.MAP_ADD, capi.F64, y[1:].copy(), x[1:].copy())
    ty, tx = dev(y), dev(x)
    rc = lib.nompk_map(capi.MAP_ADD, capi.F64, n, ty.data_ptr() + 8, tx.data_ptr() + 8, None, None, None, stream())
    capi.nompk_check(rc)
    assert np.array_equal(host(ty, np.float64)[1:], want)
    assert host(ty, np.float64)[0] == y[0]


def _reduce(op, dtype, x, y=None, mapped=False):
    lib = capi.nompk()
    npdt = NP[dtype]
    ws = torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.int64, device="cuda")
    tx = dev(x)
    ty = dev(y) if y is not None else None
    rc = lib.nompk_reduce(op, dtype, x.size, tx.data_ptr(), ty.data_ptr() if ty is not None else None,
                          res.data_ptr(), None, 0, ws.data_ptr(), stream())
    capi.nompk_check(rc, "nompk_reduce")
    first = res.cpu().numpy().view(npdt)[0]
    # second launch on the same workspace: the ticket must have been reset by the kernel
    rc = lib.nompk_reduce(op, dtype, x.size, tx.data_ptr(), ty.data_ptr() if ty is not None else None,
                          res.data_ptr(), None, 0, ws.data_ptr(), stream())
    capi.nompk_check(rc, "nompk_reduce")
    second = res.cpu().numpy().view(npdt)[0]
    assert first.tobytes() == second.tobytes(), "reduction is not deterministic / workspace not reset"
    return first


@pytest.mark.parametrize("dtype", [capi.I32, capi.U32, capi.I64, capi.U64])
@pytest.mark.parametrize("n", [0, 1, 10, 50, 1000, 4099, (1 << 22) + 5])
def test_reduce_integers_bit_exact(dtype, n):
    x = rand_array(dtype, n, 11)
    y = rand_array(dtype, n, 12)
    for op in (capi.RED_SUM, capi.RED_PROD, capi.RED_MIN, capi.RED_MAX):
        assert _reduce(op, dtype, x) == ffi.reduce_(op, dtype, x), (op, dtype, n)
    assert _reduce(capi.RED_SUM, dtype, x, y) == ffi.reduce_(capi.RED_SUM, dtype, x, y)


@pytest.mark.parametrize("n", [0, 1, 10, 50, 4099, (1 << 22) + 5])
def test_reduce_f64_exact_data(n):
    """Set X: integer-valued doubles -> every summation order is exact -> bitwise equality."""
    x = ffi.fill_int_f64(n, 1, 0, 7)
    y = ffi.fill_int_f64(n, 2, 0, 7)
    assert _reduce(capi.RED_SUM, capi.F64, x) == ffi.reduce_(capi.RED_SUM, capi.F64, x)
    assert _reduce(capi.RED_SUM, capi.F64, x, y) == ffi.reduce_(capi.RED_SUM, capi.F64, x, y)
    assert _reduce(capi.RED_MIN, capi.F64, x) == ffi.reduce_(capi.RED_MIN, capi.F64, x)
    assert _reduce(capi.RED_MAX, capi.F64, x) == ffi.reduce_(capi.RED_MAX, capi.F64, x)


@pytest.mark.parametrize("n", [1000, (1 << 22) + 5, 1 << 25])
def test_reduce_f64_random_data(n):
    """Set R: U[0.5,1.5) -> compare with the compensated oracle, 1e-12 relative (reduction-order tolerance)."""
    x = ffi.fill_uniform_f64(n, 1234, 0.5, 1.5)
    y = ffi.fill_uniform_f64(n, 4321, 0.5, 1.5)
    s = ffi.sum_compensated(x)
    d = ffi.sum_compensated(x, y)
    assert abs(_reduce(capi.RED_SUM, capi.F64, x) - s) <= 1e-12 * abs(s)
    assert abs(_reduce(capi.RED_SUM, capi.F64, x, y) - d) <= 1e-12 * abs(d)


LARGE_N = [(1 << 26) + 5, 1 << 28]   # both take one tile per CTA and the TWO-level ticket finish (reduce.cu: launch_reduce)


@pytest.mark.parametrize("n", LARGE_N)
def test_reduce_two_level_finish_f64(n):
    """BASELINE configs[1] (sum / dot, n = 2^28 fp64) and the smallest ragged size on the same code path.
    Set X (integer-valued doubles): bitwise equal to the oracle's serial loop; Set R: 1e-12 relative against the
    compensated oracle (reduction-order tolerance).  Replaces ref src/reduction.c:33-88 (whose scratch overflows beyond
    n = 2^24); closed forms of ref tests/nomp-api-500-impl.h:172-196 are covered at these sizes in test_nomp_api_gpu.py."""
    x = ffi.fill_int_f64(n, 1, 0, 7)
    y = ffi.fill_int_f64(n, 2, 0, 7)
    assert _reduce(capi.RED_SUM, capi.F64, x) == ffi.reduce_(capi.RED_SUM, capi.F64, x)
    assert _reduce(capi.RED_SUM, capi.F64, x, y) == ffi.reduce_(capi.RED_SUM, capi.F64, x, y)
    x[n - 3], x[n // 2 + 1] = -5.0, 11.0       # extrema in the last (partial) tile and in the middle
    assert _reduce(capi.RED_MIN, capi.F64, x) == -5.0 == ffi.reduce_(capi.RED_MIN, capi.F64, x)
    assert _reduce(capi.RED_MAX, capi.F64, x) == 11.0 == ffi.reduce_(capi.RED_MAX, capi.F64, x)
    del x, y
    x = ffi.fill_uniform_f64(n, 1234, 0.5, 1.5)
    y = ffi.fill_uniform_f64(n, 4321, 0.5, 1.5)
    s, d = ffi.sum_compensated(x), ffi.sum_compensated(x, y)
    assert abs(_reduce(capi.RED_SUM, capi.F64, x) - s) <= 1e-12 * abs(s)
    assert abs(_reduce(capi.RED_SUM, capi.F64, x, y) - d) <= 1e-12 * abs(d)


@pytest.mark.parametrize("n", LARGE_N)
def test_reduce_two_level_finish_i64_bit_exact(n):
    """Full-range int64 (splitmix64): wrap-around sums and dots are associative -> bit-exact in any order."""
    x = ffi.fill_i64(n, 1)
    y = ffi.fill_i64(n, 2)
    assert _reduce(capi.RED_SUM, capi.I64, x) == ffi.reduce_(capi.RED_SUM, capi.I64, x)
    assert _reduce(capi.RED_SUM, capi.I64, x, y) == ffi.reduce_(capi.RED_SUM, capi.I64, x, y)
    assert _reduce(capi.RED_MIN, capi.I64, x) == ffi.reduce_(capi.RED_MIN, capi.I64, x) == x.min()
    assert _reduce(capi.RED_MAX, capi.U64, x.view(np.uint64)) == ffi.reduce_(capi.RED_MAX, capi.U64, x.view(np.uint64))


def test_reduce_f32_small_ints():
    for n in (10, 50):
        x = np.arange(n, dtype=np.float32)
        assert _reduce(capi.RED_SUM, capi.F32, x) == np.float32(n * (n - 1) / 2)
        assert _reduce(capi.RED_SUM, capi.F32, x, x) == np.float32(n * (2 * n - 1) * (n - 1) / 6)


def test_reduce_result_in_mapped_host_memory():
    lib = capi.nompk()
    n = 1 << 20
    x = ffi.fill_int_f64(n, 9, 0, 7)
    tx = dev(x)
    ws = torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.float64, device="cuda")
    pinned = torch.zeros(2, dtype=torch.float64).pin_memory()
    rc = lib.nompk_reduce(capi.RED_SUM, capi.F64, n, tx.data_ptr(), None, res.data_ptr(), pinned.data_ptr(), 41,
                          ws.data_ptr(), stream())
    capi.nompk_check(rc)
    torch.cuda.synchronize()
    assert pinned[0].item() == x.sum() == res.item()
    assert pinned.view(torch.int64)[1].item() == 41        # the sequence number follows the value


def _ax(n, u, g, D, variant=0):
    lib = capi.nompk()
    lib.nompk_ax_set_variant(variant)
    tu, tg, tD = dev(u), dev(g), dev(D)
    tw = torch.full_like(tu, float("nan"))
    E = u.size // n ** 3
    rc = lib.nompk_ax_f64(n, E, tu.data_ptr(), tg.data_ptr(), tD.data_ptr(), tw.data_ptr(), 0, stream())
    capi.nompk_check(rc, "nompk_ax_f64")
    lib.nompk_ax_set_variant(0)
    return host(tw, np.float64)


@pytest.mark.parametrize("n", [6, 8, 10, 12])
@pytest.mark.parametrize("E", [1, 2, 7, 64, 1031])
def test_ax_exact_data_bitwise(n, E):
    """Set X: u in [-4,4], D in [-2,2], g in [0,3] integers -> all partial sums exact -> bitwise equality."""
    u = ffi.fill_int_f64(E * n ** 3, 2, -4, 4)
    g = ffi.fill_int_f64(E * 6 * n ** 3, 3, 0, 3)
    D = ffi.fill_int_f64(n * n, 4, -2, 2)
    want = ffi.ax(n, u, g, D)
    got = _ax(n, u, g, D)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("variant", list(range(18)) + [21, 22, 23] + list(range(30, 49)) + [50, 52])
@pytest.mark.parametrize("n", [6, 8, 10, 12])
def test_ax_variants_agree(variant, n):
    """Every shape / prefetch / buffer variant kept for profiling (ax.cu: dispatch_ax) computes the same bits."""
    E = 333 if n <= 10 else 167
    u = ffi.fill_int_f64(E * n ** 3, 12, -4, 4)
    g = ffi.fill_int_f64(E * 6 * n ** 3, 13, 0, 3)
    D = ffi.fill_int_f64(n * n, 14, -2, 2)
    assert np.array_equal(_ax(n, u, g, D, variant), ffi.ax(n, u, g, D))


@pytest.mark.parametrize("n", [6, 8, 10, 12])
@pytest.mark.parametrize("E", [1, 257, 1400])
def test_ax_random_data_vs_extended_oracle(n, E):
    """Set R with the real GLL derivative matrix: ||w - w_ref||_inf / ||w_ref||_inf <= 1e-12.  E = 1400 is more than one
    wave of groups on 148 SMs for every shape, so the persistent loop and its prefetch window are on the path."""
    D, _ = ffi.gll_derivative(n)
    D = np.ascontiguousarray(D.ravel())
    u = ffi.fill_uniform_f64(E * n ** 3, 1234, 0.5, 1.5)
    g = ffi.fill_uniform_f64(E * 6 * n ** 3, 99, 0.5, 1.5)
    ref = ffi.ax(n, u, g, D, "extended")
    got = _ax(n, u, g, D)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("n,E", [(6, 65536), (8, 32768), (8, 262144), (10, 32768), (10, 262144), (12, 65536)])
def test_ax_properties_at_full_size(n, E):
    """BASELINE sizes (configs[0]: E = 32768; configs[1] and [3]: E = 262144 at n = 8 and 10): size-independent
    properties instead of a full CPU recomputation.
    constant u -> w == 0 (rows of D sum to 0);  linearity A(a u + v) = a Au + Av;  symmetry v.(Au) = u.(Av);
    plus sampled elements against the oracle: 62 bit for bit on exact data, 62 to 1e-12 on the random data."""
    lib = capi.nompk()
    D, _ = ffi.gll_derivative(n)
    tD = torch.from_numpy(np.ascontiguousarray(D.ravel())).cuda()
    gen = torch.Generator(device="cuda").manual_seed(7)
    u = torch.rand(E * n ** 3, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    v = torch.rand(E * n ** 3, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    g = torch.rand(E * 6 * n ** 3, dtype=torch.float64, device="cuda", generator=gen) + 0.5

    def A(x):
        w = torch.empty_like(x)
        capi.nompk_check(lib.nompk_ax_f64(n, E, x.data_ptr(), g.data_ptr(), tD.data_ptr(), w.data_ptr(), 0, stream()))
        return w

    Au, Av = A(u), A(v)
    scale = Au.abs().max().item()
    assert A(torch.ones_like(u)).abs().max().item() <= 1e-10 * scale
    assert (A(2.5 * u + v) - (2.5 * Au + Av)).abs().max().item() <= 1e-12 * scale * 4
    lhs, rhs = torch.dot(v, Au).item(), torch.dot(u, Av).item()
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), abs(rhs))
    n3 = n ** 3
    Dr = np.ascontiguousarray(D.ravel())
    for e in list(range(0, E, E // 61)) + [E - 1]:
        ue = u[e * n3:(e + 1) * n3].cpu().numpy()
        ge = g[e * 6 * n3:(e + 1) * 6 * n3].cpu().numpy()
        ref = ffi.ax(n, ue, ge, Dr, "extended")
        assert np.abs(Au[e * n3:(e + 1) * n3].cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max(), e
    del Au, Av, v

    # exact data, sampled elements
    ui = torch.randint(-4, 5, (E * n ** 3,), device="cuda", generator=gen).double()
    gi = torch.randint(0, 4, (E * 6 * n ** 3,), device="cuda", generator=gen).double()
    Di = ffi.fill_int_f64(n * n, 4, -2, 2)
    tDi = torch.from_numpy(Di).cuda()
    w = torch.empty_like(ui)
    capi.nompk_check(lib.nompk_ax_f64(n, E, ui.data_ptr(), gi.data_ptr(), tDi.data_ptr(), w.data_ptr(), 0, stream()))
    for e in list(range(0, E, E // 61)) + [E - 1]:
        ue = ui[e * n3:(e + 1) * n3].cpu().numpy()
        ge = gi[e * 6 * n3:(e + 1) * 6 * n3].cpu().numpy()
        assert np.array_equal(w[e * n3:(e + 1) * n3].cpu().numpy(), ffi.ax(n, ue, ge, Di)), e


@pytest.mark.parametrize("n", [6, 8, 10, 12])
def test_ax_fused_with_dot_product(n):
    """nompk_ax_dot_f64: w as nompk_ax_f64, and u.(A u) from the energy form; exact data -> both bitwise equal to the
    oracle's w and to sum(u * w); random SPD-like data -> 1e-12 relative."""
    lib = capi.nompk()
    ws = torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.float64, device="cuda")
    pinned = torch.zeros(2, dtype=torch.float64).pin_memory()
    for E in (1, 5, 333, 4099):
        u = ffi.fill_int_f64(E * n ** 3, 21, -4, 4)
        g = ffi.fill_int_f64(E * 6 * n ** 3, 22, 0, 3)
        D = ffi.fill_int_f64(n * n, 23, -2, 2)
        tu, tg, tD = dev(u), dev(g), dev(D)
        tw = torch.full_like(tu, float("nan"))
        for rep in (1, 2):      # twice: the ticket must be reset
            rc = lib.nompk_ax_dot_f64(n, E, tu.data_ptr(), tg.data_ptr(), tD.data_ptr(), tw.data_ptr(), res.data_ptr(),
                                      pinned.data_ptr(), 100 * E + rep, ws.data_ptr(), 0, stream())
            capi.nompk_check(rc, "nompk_ax_dot_f64")
            torch.cuda.synchronize()
            w = ffi.ax(n, u, g, D)
            assert np.array_equal(host(tw, np.float64), w)
            assert res.item() == float(u @ w) == pinned[0].item(), (n, E)
            assert pinned.view(torch.int64)[1].item() == 100 * E + rep
    E = 257
    Dr = np.ascontiguousarray(ffi.gll_derivative(n)[0].ravel())
    u = ffi.fill_uniform_f64(E * n ** 3, 1234, 0.5, 1.5)
    g = ffi.fill_uniform_f64(E * 6 * n ** 3, 99, -0.1, 0.1).reshape(E, 6, n ** 3)
    g[:, (0, 3, 5), :] += 1.0          # G11, G22, G33 dominant: symmetric positive definite metric
    g = np.ascontiguousarray(g.ravel())
    tu, tg, tD = dev(u), dev(g), dev(Dr)
    tw = torch.empty_like(tu)
    capi.nompk_check(lib.nompk_ax_dot_f64(n, E, tu.data_ptr(), tg.data_ptr(), tD.data_ptr(), tw.data_ptr(), res.data_ptr(),
                                          None, 0, ws.data_ptr(), 0, stream()))
    ref_w = ffi.ax(n, u, g, Dr, "extended")
    ref = ffi.sum_compensated(u, ref_w)
    assert np.abs(host(tw, np.float64) - ref_w).max() <= 1e-12 * np.abs(ref_w).max()
    assert abs(res.item() - ref) <= 1e-12 * abs(ref)


@pytest.mark.parametrize("variant", [0, 60, 61, 62, 63, 64, 67, 70, 71])
@pytest.mark.parametrize("n", [6, 8, 10, 12])
def test_ax_dot_variants_agree(variant, n):
    """The shapes of the fused Ax + p.Ap kept for profiling (ax.cu: dispatch_ax_dot): same w, same u . (A u), bit for bit."""
    lib = capi.nompk()
    E = 333 if n <= 10 else 167
    u = ffi.fill_int_f64(E * n ** 3, 31, -4, 4)
    g = ffi.fill_int_f64(E * 6 * n ** 3, 32, 0, 3)
    D = ffi.fill_int_f64(n * n, 33, -2, 2)
    tu, tg, tD = dev(u), dev(g), dev(D)
    tw = torch.full_like(tu, float("nan"))
    ws = torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.float64, device="cuda")
    lib.nompk_ax_set_variant(variant)
    try:
        capi.nompk_check(lib.nompk_ax_dot_f64(n, E, tu.data_ptr(), tg.data_ptr(), tD.data_ptr(), tw.data_ptr(), res.data_ptr(),
                                              None, 0, ws.data_ptr(), 0, stream()))
        torch.cuda.synchronize()
    finally:
        lib.nompk_ax_set_variant(0)
    w = ffi.ax(n, u, g, D)
    assert np.array_equal(host(tw, np.float64), w)
    assert res.item() == float(u @ w)


@pytest.mark.parametrize("n", [6, 8, 10, 12])
def test_ax_even_odd_contractions(n):
    """NOMPK_AX_D_ANTISYMMETRIC (include/nompk.h): with a centro-antisymmetric D the contractions take their even-odd form
    (n = 8, 10, 12; the flag is ignored for n = 6).  Exact data (integer D - flip(D): its halves are multiples of 1/2) ->
    w and u . (A u) bitwise the oracle's, from the plain, the fused-dot and the xpay-fused kernel; random data with the GLL
    matrix -> 1e-12 against the extended-precision oracle, and the three kernels agree with each other bit for bit."""
    lib = capi.nompk()
    EO = 2
    ws = torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    res = torch.zeros(1, dtype=torch.float64, device="cuda")

    def three(u, g, D, E):
        tu, tg, tD = dev(u), dev(g), dev(D)
        out = []
        tw = torch.full_like(tu, float("nan"))
        capi.nompk_check(lib.nompk_ax_f64(n, E, tu.data_ptr(), tg.data_ptr(), tD.data_ptr(), tw.data_ptr(), EO, stream()))
        out.append((host(tw, np.float64), None))
        tw = torch.full_like(tu, float("nan"))
        capi.nompk_check(lib.nompk_ax_dot_f64(n, E, tu.data_ptr(), tg.data_ptr(), tD.data_ptr(), tw.data_ptr(), res.data_ptr(),
                                              None, 0, ws.data_ptr(), EO, stream()))
        torch.cuda.synchronize()
        out.append((host(tw, np.float64), res.item()))
        # p <- r + 0 * p = u, then the operator
        tp, tr = torch.zeros_like(tu), tu.clone()
        tw = torch.full_like(tu, float("nan"))
        capi.nompk_check(lib.nompk_ax_xpay_dot_peers_f64(n, E, tp.data_ptr(), tr.data_ptr(), C.c_double(0.0), None, tg.data_ptr(),
                                                         tD.data_ptr(), tw.data_ptr(), res.data_ptr(), None, 0, ws.data_ptr(), None,
                                                         EO, stream()))
        torch.cuda.synchronize()
        out.append((host(tw, np.float64), res.item()))
        return out

    for E in (1, 333, 1400):
        u = ffi.fill_int_f64(E * n ** 3, 41, -4, 4)
        g = ffi.fill_int_f64(E * 6 * n ** 3, 42, 0, 3)
        R = ffi.fill_int_f64(n * n, 43, -2, 2).reshape(n, n)
        D = np.ascontiguousarray((R - R[::-1, ::-1]).ravel())
        want = ffi.ax(n, u, g, D)
        for w, pap in three(u, g, D, E):
            assert np.array_equal(w, want), (n, E)
            assert pap is None or pap == float(u @ want)
    E = 257
    Dr = np.ascontiguousarray(ffi.gll_derivative(n)[0].ravel())
    u = ffi.fill_uniform_f64(E * n ** 3, 1234, 0.5, 1.5)
    g = ffi.fill_uniform_f64(E * 6 * n ** 3, 99, 0.5, 1.5)
    ref = ffi.ax(n, u, g, Dr, "extended")
    (w0, _), (w1, pap1), (w2, pap2) = three(u, g, Dr, E)
    assert np.abs(w0 - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.array_equal(w0, w1) and np.array_equal(w0, w2) and pap1 == pap2
    assert abs(pap1 - ffi.sum_compensated(u, ref)) <= 1e-12 * abs(pap1)
    # the general path on the same data is as close to the oracle, and (n = 8, 10, 12) not the same bits
    tu, tg, tD = dev(u), dev(g), dev(Dr)
    tw = torch.empty_like(tu)
    capi.nompk_check(lib.nompk_ax_f64(n, E, tu.data_ptr(), tg.data_ptr(), tD.data_ptr(), tw.data_ptr(), 0, stream()))
    wg = host(tw, np.float64)
    assert np.abs(wg - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.array_equal(wg, w0) == (n == 6)


def test_ax_unsupported_n_is_reported():
    lib = capi.nompk()
    t = torch.zeros(9 ** 3 * 6, dtype=torch.float64, device="cuda")
    rc = lib.nompk_ax_f64(9, 1, t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), 0, stream())
    assert rc == -3 and b"n = 9" in lib.nompk_last_error()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_reduction_fused_with_the_all_reduce_on_emulated_ranks(world):
    """nompk_reduce_peers / nompk_ax_dot_peers_f64: the finishing CTA of every rank's kernel exchanges the result through
    the peers' buffers and folds in rank order.  The ranks are emulated on this GPU (one stream, workspace and exchange
    buffer each; the kernels run concurrently and spin on each other's flags exactly as over NVLink).  One rank may use
    the stand-alone path (nompk_reduce + nompk_allreduce_scalar) on the same buffers: the protocols are the same."""
    lib = capi.nompk()
    streams = [torch.cuda.Stream() for _ in range(world)]
    bufs = [torch.zeros(lib.nompk_allreduce_xchg_bytes(world), dtype=torch.uint8, device="cuda") for _ in range(world)]
    table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    wss = [torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda") for _ in range(world)]
    ress = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    pins = [torch.zeros(3, dtype=torch.int64).pin_memory() for _ in range(world)]
    counters = torch.zeros(world, dtype=torch.int64, device="cuda")   # odd ranks: call number in device memory (seq_dev)
    seq = 0

    def peers_of(r):
        if r % 2:
            return capi.NompkPeers(table.data_ptr(), r, world, 0, counters.data_ptr() + 8 * r, pins[r].data_ptr() + 16)
        return capi.NompkPeers(table.data_ptr(), r, world, seq, None, None)

    def fold(op, parts):
        acc = parts[0]
        for p in parts[1:]:
            acc = {capi.RED_SUM: lambda a, b: a + b, capi.RED_PROD: lambda a, b: a * b, capi.RED_MIN: min, capi.RED_MAX: max}[op](acc, p)
        return acc

    cases = [(capi.RED_SUM, capi.F64, 100003, True), (capi.RED_SUM, capi.I64, 5000, False), (capi.RED_MAX, capi.F64, 1, False),
             (capi.RED_SUM, capi.F64, (1 << 22) + 3, False), (capi.RED_MIN, capi.I32, 70001, False), (capi.RED_SUM, capi.F64, 0, False)]
    # Load every stand-alone all-reduce kernel used below before anything spins: with CUDA's lazy module loading the
    # first launch of a kernel waits for the device to drain, which a kernel spinning on a peer of the SAME process
    # never lets happen (a hazard of this one-GPU emulation only: real ranks are separate processes).
    solo = torch.zeros(lib.nompk_allreduce_xchg_bytes(1), dtype=torch.uint8, device="cuda")
    solo_table = torch.tensor([solo.data_ptr()], dtype=torch.int64, device="cuda")
    for k, (op, dtype, _, _) in enumerate(cases):
        capi.nompk_check(lib.nompk_allreduce_scalar(op, dtype, ress[0].data_ptr(), None, 0, solo_table.data_ptr(), 0, 1, k + 1,
                                                    stream()))
    torch.cuda.synchronize()

    for op, dtype, n, dot in cases:
        npdt = NP[dtype]
        seq += 1
        xs = [ffi.fill_int_f64(n + r, 50 + r, 0, 7).astype(npdt) if dtype != capi.I64 else ffi.fill_i64(n + r, 60 + r)
              for r in range(world)]
        ys = [ffi.fill_int_f64(n + r, 70 + r, 0, 7).astype(npdt) if dot else None for r in range(world)]
        txs, tys = [dev(x) for x in xs], [dev(y) if y is not None else None for y in ys]
        local = [ffi.reduce_(op, dtype, xs[r], ys[r]) for r in range(world)]
        with np.errstate(over="ignore"):
            want = fold(op, [npdt(v) for v in local])
        torch.cuda.synchronize()
        for r in range(world):
            st = C.c_void_p(streams[r].cuda_stream)
            ty = tys[r].data_ptr() if tys[r] is not None else None
            if r == world - 1 and seq % 2 == 0:      # this rank takes the two-kernel path
                capi.nompk_check(lib.nompk_reduce(op, dtype, xs[r].size, txs[r].data_ptr(), ty, ress[r].data_ptr(), None, 0,
                                                  wss[r].data_ptr(), st))
                peers = peers_of(r)
                capi.nompk_check(lib.nompk_allreduce_scalar_peers(op, dtype, ress[r].data_ptr(), pins[r].data_ptr(), 1000 + seq,
                                                                  C.byref(peers), st))
            else:
                peers = peers_of(r)
                capi.nompk_check(lib.nompk_reduce_peers(op, dtype, xs[r].size, txs[r].data_ptr(), ty, ress[r].data_ptr(),
                                                        pins[r].data_ptr(), 1000 + seq, wss[r].data_ptr(), C.byref(peers), st))
        torch.cuda.synchronize()
        for r in range(world):
            got = ress[r].cpu().numpy().view(npdt)[0]
            assert got.tobytes() == np.array([want], dtype=npdt).tobytes()[: got.nbytes], (op, dtype, n, r, got, want)
            assert pins[r].numpy().view(npdt)[0].tobytes() == got.tobytes() and pins[r][1].item() == 1000 + seq
            assert pins[r][2].item() == 0            # nobody timed out

    # Ax fused with p.Ap fused with the all-reduce
    n, D = 8, ffi.fill_int_f64(64, 23, -2, 2)
    tD = dev(D)
    us = [ffi.fill_int_f64((3 + r) * 512, 21 + r, -4, 4) for r in range(world)]
    gs_ = [ffi.fill_int_f64((3 + r) * 6 * 512, 31 + r, 0, 3) for r in range(world)]
    tus, tgs = [dev(u) for u in us], [dev(g) for g in gs_]
    tws = [torch.empty_like(t) for t in tus]
    fres = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
    seq += 1
    torch.cuda.synchronize()
    for r in range(world):
        peers = peers_of(r)
        capi.nompk_check(lib.nompk_ax_dot_peers_f64(n, 3 + r, tus[r].data_ptr(), tgs[r].data_ptr(), tD.data_ptr(), tws[r].data_ptr(),
                                                    fres[r].data_ptr(), None, 0, wss[r].data_ptr(), C.byref(peers), 0,
                                                    C.c_void_p(streams[r].cuda_stream)))
    torch.cuda.synchronize()
    parts = [float(us[r] @ ffi.ax(n, us[r], gs_[r], D)) for r in range(world)]
    for r in range(world):
        assert np.array_equal(host(tws[r], np.float64), ffi.ax(n, us[r], gs_[r], D))
        assert fres[r].item() == fold(capi.RED_SUM, parts)
    assert counters.cpu().tolist() == [seq if r % 2 else 0 for r in range(world)]   # the kernels counted the calls themselves


@pytest.mark.parametrize("world", [2, 4])
def test_fused_all_reduce_replayed_from_cuda_graphs(world):
    """The number of a collective call is a counter in device memory that the finishing warp advances
    (nompk_peers_t.seq_dev), so nothing in a launch is specific to one call: each emulated rank captures
    `dot -> all-reduce` ONCE into a CUDA graph and replays it; every replay gives the fold of the ranks' current data in
    rank order, bit for bit, and the two exchange slots keep alternating."""
    lib = capi.nompk()
    n = 70001
    streams = [torch.cuda.Stream() for _ in range(world)]
    bufs = [torch.zeros(lib.nompk_allreduce_xchg_bytes(world), dtype=torch.uint8, device="cuda") for _ in range(world)]
    table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    wss = [torch.zeros(lib.nompk_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda") for _ in range(world)]
    ress = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
    counters = torch.zeros(world, dtype=torch.int64, device="cuda")
    errs = [torch.zeros(1, dtype=torch.int64).pin_memory() for _ in range(world)]
    xs = [dev(ffi.fill_int_f64(n + r, 50 + r, 0, 7)) for r in range(world)]
    ys = [dev(ffi.fill_int_f64(n + r, 70 + r, 0, 7)) for r in range(world)]
    peers = [capi.NompkPeers(table.data_ptr(), r, world, 0, counters.data_ptr() + 8 * r, errs[r].data_ptr()) for r in range(world)]

    def launch(r):
        capi.nompk_check(lib.nompk_reduce_peers(capi.RED_SUM, capi.F64, n + r, xs[r].data_ptr(), ys[r].data_ptr(), ress[r].data_ptr(),
                                                None, 0, wss[r].data_ptr(), C.byref(peers[r]), C.c_void_p(streams[r].cuda_stream)))

    for r in range(world):          # call 1, launched normally (also loads the kernel before anything is captured)
        launch(r)
    torch.cuda.synchronize()
    graphs = []
    for r in range(world):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=streams[r]):
            launch(r)
        graphs.append(g)
    torch.cuda.synchronize()
    assert counters.cpu().tolist() == [1] * world                    # capturing executed nothing
    for replay in range(1, 6):
        for r in range(world):
            xs[r].mul_(2.0)                                            # new data for every replay
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                graphs[r].replay()
        torch.cuda.synchronize()
        want = sum(float(torch.dot(xs[r], ys[r]).item()) for r in range(world))   # small integers times 2^k: exact
        for r in range(world):
            assert ress[r].item() == want, (replay, r)
            assert errs[r].item() == 0
        assert counters.cpu().tolist() == [1 + replay] * world
