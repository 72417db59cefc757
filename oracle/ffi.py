"""ctypes binding of oracle/libnomp_oracle.so (TEST INFRASTRUCTURE, see oracle/nomp_oracle.c)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_lib = None

I32, U32, I64, U64, F32, F64 = range(6)
NP_DTYPES = {I32: np.int32, U32: np.uint32, I64: np.int64, U64: np.uint64, F32: np.float32, F64: np.float64}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        so = HERE / "libnomp_oracle.so"
        if not so.exists() or so.stat().st_mtime < (HERE / "nomp_oracle.c").stat().st_mtime:
            subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
        L = C.CDLL(str(so))
        vp, sz, i, d, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_uint64
        L.oracle_num_threads.restype = i
        L.oracle_set_num_threads.argtypes = [i]
        L.oracle_map.restype = i
        L.oracle_map.argtypes = [i, i, sz, vp, vp, vp, vp, vp]
        L.oracle_reduce.restype = i
        L.oracle_reduce.argtypes = [i, i, sz, vp, vp, vp]
        L.oracle_sum_f64_compensated.restype = d
        L.oracle_sum_f64_compensated.argtypes = [sz, vp, vp]
        L.oracle_sum_f64_mt.restype = d
        L.oracle_sum_f64_mt.argtypes = [sz, vp, vp]
        L.oracle_sum_i64_mt.restype = C.c_int64
        L.oracle_sum_i64_mt.argtypes = [sz, vp, vp]
        L.oracle_axpy_f64_mt.restype = None
        L.oracle_axpy_f64_mt.argtypes = [sz, d, vp, vp]
        for f in ("oracle_ax_f64", "oracle_ax_f64_extended", "oracle_ax_f64_mt"):
            getattr(L, f).restype = i
            getattr(L, f).argtypes = [i, sz, vp, vp, vp, vp]
        L.oracle_fill_int_f64.restype = None
        L.oracle_fill_int_f64.argtypes = [vp, sz, u64, i, i, sz]
        L.oracle_fill_uniform_f64.restype = None
        L.oracle_fill_uniform_f64.argtypes = [vp, sz, u64, d, d, sz]
        L.oracle_fill_i64.restype = None
        L.oracle_fill_i64.argtypes = [vp, sz, u64, sz]
        L.oracle_gs.restype = i
        L.oracle_gs.argtypes = [i, i, vp, sz, vp, i, vp]
        L.oracle_gll_derivative.restype = i
        L.oracle_gll_derivative.argtypes = [i, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def map_(op, dtype, y, x=None, z=None, alpha=None, beta=None):
    """In-place y <- op(...) with the serial C loop; returns y."""
    npdt = NP_DTYPES[dtype]
    a = None if alpha is None else np.array([alpha], dtype=npdt)
    b = None if beta is None else np.array([beta], dtype=npdt)
    rc = lib().oracle_map(op, dtype, y.size, _p(y), _p(x), _p(z), _p(a), _p(b))
    assert rc == 0
    return y


def reduce_(op, dtype, x, y=None):
    out = np.zeros(1, dtype=NP_DTYPES[dtype])
    rc = lib().oracle_reduce(op, dtype, x.size, _p(x), _p(y), _p(out))
    assert rc == 0
    return out[0]


def sum_compensated(x, y=None):
    return lib().oracle_sum_f64_compensated(x.size, _p(x), _p(y))


def ax(n, u, g, D, mode="plain"):
    E = u.size // (n ** 3)
    w = np.empty_like(u)
    fn = {"plain": lib().oracle_ax_f64, "extended": lib().oracle_ax_f64_extended, "mt": lib().oracle_ax_f64_mt}[mode]
    rc = fn(n, E, _p(u), _p(g), _p(D), _p(w))
    assert rc == 0
    return w


def fill_int_f64(n, seed, lo, hi, first=0):
    a = np.empty(n, dtype=np.float64)
    lib().oracle_fill_int_f64(_p(a), n, seed, lo, hi, first)
    return a


def fill_uniform_f64(n, seed, lo, hi, first=0):
    a = np.empty(n, dtype=np.float64)
    lib().oracle_fill_uniform_f64(_p(a), n, seed, lo, hi, first)
    return a


def fill_i64(n, seed, first=0):
    a = np.empty(n, dtype=np.int64)
    lib().oracle_fill_i64(_p(a), n, seed, first)
    return a


def gll_derivative(n):
    D = np.empty((n, n), dtype=np.float64)
    x = np.empty(n, dtype=np.float64)
    rc = lib().oracle_gll_derivative(n, _p(D), _p(x))
    assert rc == 0
    return D, x


def gs(op, dtype, ids, v, segments=None):
    """In-place gather-scatter of v under the global numbering ids (int64); segments = rank boundaries or None."""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    assert v.dtype == NP_DTYPES[dtype] and v.flags.c_contiguous and ids.size == v.size
    seg = None if segments is None else np.ascontiguousarray(segments, dtype=np.uint64)
    rc = lib().oracle_gs(op, dtype, ids.ctypes.data, v.size, v.ctypes.data, 0 if seg is None else seg.size - 1, _p(seg))
    assert rc == 0
    return v


def box_ids(n, ex, ey, ez, z0=0, nz=None):
    """Global ids (1-based, lexicographic over the (n-1)*e + 1 points per direction) of the local degrees of freedom
    of a box of ex x ey x ez hexahedral elements with n points per direction, element-major (x fastest), for the
    elements with z index in [z0, z0 + nz): the slab of one rank."""
    nz = ez if nz is None else nz
    N = n - 1
    px, py = N * ex + 1, N * ey + 1
    e_z, e_y, e_x = np.meshgrid(np.arange(z0, z0 + nz), np.arange(ey), np.arange(ex), indexing="ij")
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    gx = (e_x.reshape(-1, 1, 1, 1) * N + i[None]).astype(np.int64)
    gy = (e_y.reshape(-1, 1, 1, 1) * N + j[None]).astype(np.int64)
    gz = (e_z.reshape(-1, 1, 1, 1) * N + k[None]).astype(np.int64)
    return (1 + gx + px * (gy + py * gz)).reshape(-1)
