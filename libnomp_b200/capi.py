"""ctypes bindings of the two C ABIs (include/nompk.h, include/nomp.h).

These are the calls a reference-side binding would make; tests and bench.py go through them so
that everything measured crosses the same C boundary as a C caller.  Missing libraries raise ImportError-like
errors: there is no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import LIB_DIR, INSTALL_DIR

# ---- enums of include/nompk.h -----------------------------------------------------------------------------
I32, U32, I64, U64, F32, F64 = range(6)
MAP_ADD, MAP_SUB, MAP_MUL, MAP_AXPY, MAP_XPAY, MAP_AXPBY, MAP_SCALE, MAP_COPY, MAP_FILL, MAP_ADD3 = range(10)
RED_SUM, RED_PROD, RED_MIN, RED_MAX = range(4)
AX_D_CACHED = 1

# ---- enums of include/nomp.h ------------------------------------------------------------------------------
NOMP_INT, NOMP_UINT, NOMP_FLOAT, NOMP_PTR = 2048, 4096, 8192, 16384
NOMP_JIT = 1
NOMP_ALLOC, NOMP_TO, NOMP_FROM, NOMP_FREE = 1, 2, 4, 8
NOMP_USER_INPUT_IS_INVALID = -128
NOMP_USER_MAP_PTR_IS_INVALID = -130
NOMP_USER_MAP_OP_IS_INVALID = -132
NOMP_USER_LOG_ID_IS_INVALID = -134
NOMP_INITIALIZE_FAILURE = -256
NOMP_FINALIZE_FAILURE = -258
NOMP_PY_CALL_FAILURE = -384
NOMP_LOOPY_CONVERSION_FAILURE = -386
NOMP_LOOPY_KNL_NAME_NOT_FOUND = -388
NOMP_LOOPY_CODEGEN_FAILURE = -390
NOMP_LOOPY_GRIDSIZE_FAILURE = -392
NOMP_CUDA_FAILURE = -512

NOMPK_SYMBOLS = [
    "nompk_version", "nompk_last_error", "nompk_dtype_size", "nompk_map", "nompk_reduce_workspace_bytes", "nompk_reduce_workspace_layout",
    "nompk_reduce", "nompk_allreduce_xchg_bytes", "nompk_allreduce_scalar", "nompk_allreduce_scalar_peers", "nompk_ax_f64", "nompk_ax_dot_f64", "nompk_ax_supported", "nompk_ax_set_variant", "nompk_launch_count",
    "nompk_reduce_peers", "nompk_ax_dot_peers_f64", "nompk_ax_xpay_dot_peers_f64",
    "nompk_gs_create", "nompk_gs_unique", "nompk_gs_match_peer", "nompk_gs_finalize_setup", "nompk_gs_recv_offsets",
    "nompk_gs_connect", "nompk_gs_apply", "nompk_gs_stats", "nompk_gs_destroy",
]
NOMP_SYMBOLS = [
    "nomp_init", "nomp_update", "nomp_jit", "nomp_run", "nomp_sync", "nomp_get_err_str", "nomp_get_err_no",
    "nomp_finalize", "nomp_finalize_excluding_interpreter", "nomp_copy_env",
    # extensions declared in include/nomp-b200.h
    "nomp_b200_stream", "nomp_b200_update_async", "nomp_b200_device_ptr", "nomp_b200_launch_count", "nomp_b200_comm_rank",
    "nomp_b200_comm_size", "nomp_b200_comm_uses_nvlink_kernel", "nomp_b200_prog_info", "nomp_b200_exchange_blob",
    "nomp_b200_jit_cache_stats", "nomp_b200_sha256_hex", "nomp_b200_gs_setup", "nomp_b200_gs", "nomp_b200_gs_info",
    "nomp_b200_gs_free", "nomp_b200_jit_cache_dir", "nomp_b200_jit_cache_put", "nomp_b200_jit_cache_get",
    "nomp_b200_device_reductions", "nomp_b200_graph_begin", "nomp_b200_graph_end", "nomp_b200_graph_launch", "nomp_b200_graph_free",
]


class NompkPeers(C.Structure):
    """nompk_peers_t of include/nompk.h."""
    _fields_ = [("peer_xchg", C.c_void_p), ("rank", C.c_int), ("world", C.c_int), ("seq", C.c_ulonglong),
                ("seq_dev", C.c_void_p), ("error_host_mapped", C.c_void_p)]


class NativeLibraryMissing(RuntimeError):
    pass


def _load(name: str) -> C.CDLL:
    path = Path(LIB_DIR) / name
    if not path.exists():
        raise NativeLibraryMissing(
            f"{path} is not built; run `python -m libnomp_b200.build` (there is no CPU fallback)")
    return C.CDLL(str(path), mode=C.RTLD_GLOBAL)


_nompk = None
_nomp = None


def nompk() -> C.CDLL:
    """libnompk.so with argtypes set."""
    global _nompk
    if _nompk is None:
        lib = _load("libnompk.so")
        lib.nompk_version.restype = C.c_int
        lib.nompk_last_error.restype = C.c_char_p
        lib.nompk_dtype_size.restype = C.c_size_t
        lib.nompk_dtype_size.argtypes = [C.c_int]
        lib.nompk_map.restype = C.c_int
        lib.nompk_map.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
        lib.nompk_reduce_workspace_bytes.restype = C.c_size_t
        lib.nompk_reduce_workspace_layout.restype = None
        lib.nompk_reduce_workspace_layout.argtypes = [C.POINTER(C.c_size_t)]
        lib.nompk_reduce.restype = C.c_int
        lib.nompk_reduce.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_ulonglong, C.c_void_p, C.c_void_p]
        lib.nompk_allreduce_xchg_bytes.restype = C.c_size_t
        lib.nompk_allreduce_xchg_bytes.argtypes = [C.c_int]
        lib.nompk_allreduce_scalar.restype = C.c_int
        lib.nompk_allreduce_scalar.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_int,
                                               C.c_int, C.c_ulonglong, C.c_void_p]
        lib.nompk_allreduce_scalar_peers.restype = C.c_int
        lib.nompk_allreduce_scalar_peers.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_ulonglong, C.POINTER(NompkPeers),
                                                     C.c_void_p]
        lib.nompk_ax_f64.restype = C.c_int
        lib.nompk_ax_f64.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint,
                                     C.c_void_p]
        lib.nompk_ax_dot_f64.restype = C.c_int
        lib.nompk_ax_dot_f64.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_uint, C.c_void_p]
        lib.nompk_reduce_peers.restype = C.c_int
        lib.nompk_reduce_peers.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_ulonglong, C.c_void_p, C.POINTER(NompkPeers), C.c_void_p]
        lib.nompk_ax_dot_peers_f64.restype = C.c_int
        lib.nompk_ax_dot_peers_f64.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_ulonglong, C.c_void_p, C.POINTER(NompkPeers), C.c_uint,
                                               C.c_void_p]
        lib.nompk_ax_xpay_dot_peers_f64.restype = C.c_int
        lib.nompk_ax_xpay_dot_peers_f64.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_void_p,
                                                    C.POINTER(NompkPeers), C.c_uint, C.c_void_p]
        lib.nompk_ax_supported.restype = C.c_int
        lib.nompk_ax_supported.argtypes = [C.c_int]
        lib.nompk_ax_set_variant.restype = C.c_int
        lib.nompk_ax_set_variant.argtypes = [C.c_int]
        lib.nompk_launch_count.restype = C.c_ulonglong
        vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
        for name, args in (("nompk_gs_create", [vp, sz, C.POINTER(vp), vp]),
                           ("nompk_gs_unique", [vp, C.POINTER(vp), C.POINTER(sz)]),
                           ("nompk_gs_match_peer", [vp, i, i, vp, sz, C.POINTER(sz), vp]),
                           ("nompk_gs_finalize_setup", [vp, i, i, C.POINTER(sz), vp]),
                           ("nompk_gs_recv_offsets", [vp, C.POINTER(sz), C.POINTER(sz)]),
                           ("nompk_gs_connect", [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz), vp]),
                           ("nompk_gs_apply", [vp, i, i, vp, vp, vp]),
                           ("nompk_gs_stats", [vp, C.POINTER(sz * 8)])):
            getattr(lib, name).restype = i
            getattr(lib, name).argtypes = args
        lib.nompk_gs_destroy.restype = None
        lib.nompk_gs_destroy.argtypes = [vp]
        _nompk = lib
    return _nompk


def nompk_check(rc: int, what: str = "nompk call"):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {nompk().nompk_last_error().decode()}")


def nomp() -> C.CDLL:
    """libnomp.so.  nomp_jit / nomp_run are variadic: callers pass ctypes objects of the exact C types."""
    global _nomp
    if _nomp is None:
        nompk()  # dependency, loaded first so that $ORIGIN lookups are not needed from Python
        lib = _load("libnomp.so")
        lib.nomp_init.restype = C.c_int
        lib.nomp_init.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        lib.nomp_update.restype = C.c_int
        lib.nomp_update.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
        lib.nomp_jit.restype = C.c_int
        lib.nomp_run.restype = C.c_int
        lib.nomp_sync.restype = C.c_int
        lib.nomp_get_err_str.restype = C.c_void_p  # caller frees
        lib.nomp_get_err_str.argtypes = [C.c_uint]
        lib.nomp_get_err_no.restype = C.c_int
        lib.nomp_get_err_no.argtypes = [C.c_uint]
        lib.nomp_finalize.restype = C.c_int
        lib.nomp_finalize_excluding_interpreter.restype = C.c_int
        lib.nomp_b200_stream.restype = C.c_void_p
        lib.nomp_b200_update_async.restype = C.c_int
        lib.nomp_b200_update_async.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
        lib.nomp_b200_device_ptr.restype = C.c_void_p
        lib.nomp_b200_device_ptr.argtypes = [C.c_void_p]
        lib.nomp_b200_launch_count.restype = C.c_ulonglong
        lib.nomp_b200_comm_rank.restype = C.c_int
        lib.nomp_b200_comm_size.restype = C.c_int
        lib.nomp_b200_comm_uses_nvlink_kernel.restype = C.c_int
        lib.nomp_b200_prog_info.restype = C.c_char_p
        lib.nomp_b200_prog_info.argtypes = [C.c_int]
        lib.nomp_b200_exchange_blob.restype = C.c_int
        lib.nomp_b200_exchange_blob.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_size_t]
        lib.nomp_b200_gs_setup.restype = C.c_int
        lib.nomp_b200_gs_setup.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_size_t]
        lib.nomp_b200_gs.restype = C.c_int
        lib.nomp_b200_gs.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_char_p]
        lib.nomp_b200_gs_info.restype = C.c_int
        lib.nomp_b200_gs_info.argtypes = [C.c_int, C.POINTER(C.c_size_t * 8)]
        lib.nomp_b200_gs_free.restype = C.c_int
        lib.nomp_b200_gs_free.argtypes = [C.c_int]
        lib.nomp_b200_jit_cache_stats.restype = None
        lib.nomp_b200_jit_cache_stats.argtypes = [C.POINTER(C.c_ulonglong * 4)]
        lib.nomp_b200_jit_cache_dir.restype = C.c_char_p
        lib.nomp_b200_jit_cache_put.restype = C.c_int
        lib.nomp_b200_jit_cache_put.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        lib.nomp_b200_jit_cache_get.restype = C.c_int
        lib.nomp_b200_jit_cache_get.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.nomp_b200_sha256_hex.restype = None
        lib.nomp_b200_sha256_hex.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        _nomp = lib
    return _nomp


def gs_setup(ids) -> int:
    """nomp_b200_gs_setup for a contiguous int64 numpy array of global ids; returns the handle."""
    handle = C.c_int(-1)
    check(nomp().nomp_b200_gs_setup(C.byref(handle), ids.ctypes.data, ids.size))
    return handle.value


def gs(handle: int, vec, op: str = "+") -> None:
    """nomp_b200_gs on a mapped numpy array."""
    kind = NOMP_FLOAT if vec.dtype.kind == "f" else NOMP_UINT if vec.dtype.kind == "u" else NOMP_INT
    check(nomp().nomp_b200_gs(handle, vec.ctypes.data, vec.itemsize, kind, op.encode()))


def gs_info(handle: int) -> dict:
    out = (C.c_size_t * 8)()
    check(nomp().nomp_b200_gs_info(handle, C.byref(out)))
    return dict(zip(("n", "distinct", "groups", "copies", "shared_groups", "shared_pairs", "neighbours", "shared_ids"),
                    (int(v) for v in out)))


def jit_cache_stats() -> dict:
    """Counters of the on-disk JIT cache since the library was loaded (include/nomp-b200.h)."""
    out = (C.c_ulonglong * 4)()
    nomp().nomp_b200_jit_cache_stats(C.byref(out))
    return dict(zip(("knl_hits", "knl_misses", "cubin_hits", "cubin_misses"), (int(v) for v in out)))


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


class NompError(RuntimeError):
    def __init__(self, errno: int, text: str):
        super().__init__(f"libnomp error {errno}: {text}")
        self.errno = errno
        self.text = text


def err_info(err_id: int):
    """(error number, message) for a positive log id returned by a nomp_* call."""
    lib = nomp()
    no = lib.nomp_get_err_no(C.c_uint(err_id & 0xFFFFFFFF))
    p = lib.nomp_get_err_str(C.c_uint(err_id & 0xFFFFFFFF))
    text = ""
    if p:
        text = C.string_at(p).decode(errors="replace")
        _libc.free(p)
    return no, text


def check(err_id: int):
    if err_id != 0:
        raise NompError(*err_info(err_id))


def init(backend="cuda", device=0, verbose=0, extra=(), install_dir=None, annotations_script=None, scripts_dir=None):
    args = ["python", "--nomp-backend", backend, "--nomp-device", str(device), "--nomp-platform", "0",
            "--nomp-install-dir", str(install_dir or INSTALL_DIR), "--nomp-verbose", str(verbose)]
    if annotations_script:
        args += ["--nomp-annotations-script", annotations_script]
    if scripts_dir:
        args += ["--nomp-scripts-dir", str(scripts_dir)]
    args += list(extra)
    argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
    return nomp().nomp_init(len(args), argv)


def clauses(*triples):
    """NULL-terminated clause array, e.g. clauses(("transform", "mod", "fn"), ("reduce", "s", "+"))."""
    flat = []
    for t in triples:
        flat += [None if x is None else x.encode() for x in t]
    flat.append(None)
    return (C.c_char_p * len(flat))(*flat)


def jit(src: str, clause_arr, args):
    """nomp_jit wrapper. args: list of (name, size, type[, value_ctypes_obj for NOMP_JIT])."""
    kid = C.c_int(-1)
    va = []
    for a in args:
        va += [C.c_char_p(a[0].encode()), C.c_size_t(a[1]), C.c_int(a[2])]
        if a[2] & NOMP_JIT:
            va.append(C.cast(C.pointer(a[3]), C.c_void_p))
    err = nomp().nomp_jit(C.byref(kid), C.c_char_p(src.encode()), clause_arr, C.c_int(len(args)), *va)
    return err, kid.value


def run(kid: int, *ptrs):
    """nomp_run wrapper; each argument is an int address or a ctypes object passed by reference."""
    va = []
    for p in ptrs:
        if isinstance(p, int):
            va.append(C.c_void_p(p))
        else:
            va.append(C.cast(C.pointer(p), C.c_void_p))
    return nomp().nomp_run(C.c_int(kid), *va)


def update(ptr: int, i0: int, i1: int, usize: int, op: int):
    return nomp().nomp_update(C.c_void_p(ptr), i0, i1, usize, op)


def update_async(ptr: int, i0: int, i1: int, usize: int, op: int):
    return nomp().nomp_b200_update_async(C.c_void_p(ptr), i0, i1, usize, op)
