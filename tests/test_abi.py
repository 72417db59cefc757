"""The C-ABI boundary without a GPU: both shared libraries load, export every function their headers declare, and the
parts of the public API that need no device behave like the reference (error ids, messages, configuration parsing)."""
import ctypes as C
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from libnomp_b200 import INSTALL_DIR, capi  # noqa: E402

DECL = re.compile(r"^\s*(?:const\s+)?(?:unsigned\s+long\s+long|unsigned|int|void|char|size_t)\s*\*?\s*(nomp\w+|nompk\w+)\s*\(", re.M)


def declared(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(DECL.findall(text)))


def test_headers_and_libraries_agree():
    k, n = capi.nompk(), capi.nomp()
    want_k = declared("nompk.h")
    assert len(want_k) >= 10
    for sym in want_k:
        assert hasattr(k, sym), f"libnompk.so does not export {sym}"
    want_n = declared("nomp.h") + declared("nomp-aux.h") + declared("nomp-b200.h")
    assert {"nomp_init", "nomp_update", "nomp_jit", "nomp_run", "nomp_sync", "nomp_get_err_str", "nomp_get_err_no",
            "nomp_finalize", "nomp_finalize_excluding_interpreter", "nomp_copy_env"} <= set(want_n)
    for sym in want_n:
        assert hasattr(n, sym), f"libnomp.so does not export {sym}"
    assert sorted(capi.NOMPK_SYMBOLS) == want_k
    assert k.nompk_version() == 200 and k.nompk_reduce_workspace_bytes() >= 65536 * 8
    assert [k.nompk_dtype_size(d) for d in range(6)] == [4, 4, 8, 8, 4, 8]
    assert [k.nompk_ax_supported(x) for x in (6, 7, 8, 9, 10, 12)] == [1, 0, 1, 0, 1, 1]


def test_generated_reduce_kernels_use_the_library_workspace_layout():
    sys.path.insert(0, str(ROOT / "libnomp_b200" / "python"))
    from nomp_bridge import reduction as r
    off = (C.c_size_t * 4)()
    capi.nompk().nompk_reduce_workspace_layout(off)
    assert list(off) == [r.WS_TICKET, r.WS_GROUP_TICKET, r.WS_L2, r.WS_L1]
    assert capi.nompk().nompk_reduce_workspace_bytes() == r.WS_L1 + 8 * r.RED_MAX_CTAS


def test_public_enum_values_are_the_reference_abi():
    text = (ROOT / "include" / "nomp.h").read_text()
    for name, value in dict(NOMP_INT=2048, NOMP_UINT=4096, NOMP_FLOAT=8192, NOMP_PTR=16384, NOMP_ALLOC=1, NOMP_TO=2,
                            NOMP_FROM=4, NOMP_FREE=8, NOMP_JIT=1).items():
        assert re.search(rf"\b{name}\s*=\s*{value}\b", text), name
    for name, value in dict(NOMP_USER_INPUT_IS_INVALID=-128, NOMP_USER_MAP_PTR_IS_INVALID=-130, NOMP_USER_MAP_OP_IS_INVALID=-132,
                            NOMP_USER_LOG_ID_IS_INVALID=-134, NOMP_INITIALIZE_FAILURE=-256, NOMP_FINALIZE_FAILURE=-258,
                            NOMP_PY_CALL_FAILURE=-384, NOMP_LOOPY_CONVERSION_FAILURE=-386, NOMP_LOOPY_KNL_NAME_NOT_FOUND=-388,
                            NOMP_LOOPY_CODEGEN_FAILURE=-390, NOMP_LOOPY_GRIDSIZE_FAILURE=-392, NOMP_CUDA_FAILURE=-512).items():
        assert re.search(rf"#define {name} \({value}\)", text), name


def test_kernel_library_has_only_sm_100a_code():
    """No multi-architecture dispatch: the fat binary holds exactly one target."""
    out = subprocess.run(["cuobjdump", "--list-elf", str(ROOT / "libnomp_b200" / "lib" / "libnompk.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


# The following run in a child process: nomp_init() touches process-wide state (embedded interpreter, CUDA).
CHILD = r"""
import ctypes as C, sys, os
sys.path.insert(0, {root!r})
from libnomp_b200 import capi
lib = capi.nomp()
def argv(*a):
    return len(a), (C.c_char_p * len(a))(*[x.encode() for x in a])
out = []
# finalize before init: the raw code, and it is not a valid log id
e = lib.nomp_finalize(); out.append(("finalize_first", e, lib.nomp_get_err_no(C.c_uint(e & 0xffffffff))))
# missing value after the last flag
e = lib.nomp_init(*argv("--nomp-backend", "cuda", "--nomp-device", "0", "--nomp-platform")); out.append(("missing_value",) + capi.err_info(e))
# backend / install dir missing
os.environ.pop("NOMP_INSTALL_DIR", None)
e = lib.nomp_init(*argv("prog", "--nomp-backend", "cuda")); out.append(("no_install_dir",) + capi.err_info(e))
e = lib.nomp_init(*argv("prog", "--nomp-install-dir", {inst!r})); out.append(("no_backend",) + capi.err_info(e))
# invalid numeric values, from the environment (which overrides the command line)
os.environ["NOMP_DEVICE"] = "invalid"
e = lib.nomp_init(*argv("prog", "--nomp-backend", "cuda", "--nomp-install-dir", {inst!r}, "--nomp-device", "0")); out.append(("bad_device",) + capi.err_info(e))
del os.environ["NOMP_DEVICE"]
os.environ["NOMP_BACKEND"] = "invalid"
e = lib.nomp_init(*argv("prog", "--nomp-backend", "cuda", "--nomp-install-dir", {inst!r})); out.append(("bad_backend",) + capi.err_info(e))
del os.environ["NOMP_BACKEND"]
e = lib.nomp_init(*argv("prog", "--nomp-backend", "OpenCL", "--nomp-install-dir", {inst!r})); out.append(("opencl",) + capi.err_info(e))
# every failed init leaves the runtime uninitialised
out.append(("finalize_after_failures", lib.nomp_finalize()))
out.append(("run_invalid",) + capi.err_info(lib.nomp_run(C.c_int(-1))))
out.append(("bad_id", lib.nomp_get_err_no(C.c_uint(0)), lib.nomp_get_err_no(C.c_uint(100000))))
for o in out: print(repr(o))
"""


def test_configuration_and_error_registry_without_a_device():
    code = CHILD.format(root=str(ROOT), inst=str(INSTALL_DIR))
    env = {k: v for k, v in os.environ.items() if not k.startswith("NOMP_")}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = {}
    for line in r.stdout.splitlines():
        t = eval(line)
        rows[t[0]] = t[1:]
    assert rows["finalize_first"] == (capi.NOMP_FINALIZE_FAILURE, capi.NOMP_USER_LOG_ID_IS_INVALID)
    no, text = rows["missing_value"]
    assert no == capi.NOMP_USER_INPUT_IS_INVALID and re.search(r"libnomp/src/nomp\.c:\d+ Missing argument value after: --nomp-platform\.", text)
    assert rows["no_install_dir"][0] == capi.NOMP_USER_INPUT_IS_INVALID and "NOMP_INSTALL_DIR is missing or invalid" in rows["no_install_dir"][1]
    assert rows["no_backend"][0] == capi.NOMP_USER_INPUT_IS_INVALID and "NOMP_BACKEND is missing or invalid" in rows["no_backend"][1]
    assert rows["bad_device"][0] == capi.NOMP_USER_INPUT_IS_INVALID and "NOMP_DEVICE is missing or invalid" in rows["bad_device"][1]
    assert rows["bad_backend"] == (capi.NOMP_USER_INPUT_IS_INVALID, rows["bad_backend"][1]) and "Invalid backend: invalid." in rows["bad_backend"][1]
    assert "Invalid backend: opencl." in rows["opencl"][1], "CUDA is the only backend; names are lower-cased"
    assert rows["finalize_after_failures"] == (capi.NOMP_FINALIZE_FAILURE,)
    no, text = rows["run_invalid"]
    assert no == capi.NOMP_USER_INPUT_IS_INVALID and re.search(r"\[Error\] .*/src/nomp\.c:\d+ Kernel id -1 passed to nomp_run is not valid\.", text)
    assert rows["bad_id"] == (capi.NOMP_USER_LOG_ID_IS_INVALID, capi.NOMP_USER_LOG_ID_IS_INVALID)


def test_product_does_not_reference_the_oracle():
    """No CPU fallback: neither library links the oracle, and no product source mentions it."""
    for lib in ("libnomp.so", "libnompk.so"):
        out = subprocess.run(["ldd", str(ROOT / "libnomp_b200" / "lib" / lib)], capture_output=True, text=True).stdout
        assert "oracle" not in out
    for path in list((ROOT / "libnomp_b200").rglob("*.c")) + list((ROOT / "libnomp_b200").rglob("*.cu")) + \
            list((ROOT / "libnomp_b200").rglob("*.py")):
        if "build" in path.parts:
            continue
        text = path.read_text()
        assert "libnomp_oracle" not in text and "from oracle" not in text and "import oracle" not in text or path.name == "build.py", path


def test_aux_helpers():
    lib = capi.nomp()
    lib.nomp_str_toui.argtypes = [C.c_char_p, C.c_size_t]
    assert [lib.nomp_str_toui(s, 128) for s in (b"0", b"17", b"invalid", b"-3", b"", b"12x")] == [0, 17, -1, -1, -1, -1]
    lib.nomp_copy_env.restype = C.c_void_p
    lib.nomp_copy_env.argtypes = [C.c_char_p, C.c_size_t]
    os.environ["NOMP_TEST_ENV_VALUE"] = "hello world"
    p = lib.nomp_copy_env(b"NOMP_TEST_ENV_VALUE", 5)
    assert C.string_at(p) == b"hello"
    assert lib.nomp_copy_env(b"NOMP_TEST_ENV_UNSET", 5) is None
    lib.nomp_max.restype = C.c_int
    assert lib.nomp_max(C.c_uint(3), C.c_int(-5), C.c_int(9), C.c_int(2)) == 9


def test_jit_cache_hash_is_sha256():
    """Cache keys (src/jitcache.c) are SHA-256: FIPS 180-4 test vectors and random lengths around the block edges."""
    import hashlib

    import numpy as np
    lib = capi.nomp()
    out = C.create_string_buffer(65)
    lib.nomp_b200_sha256_hex(b"abc", 3, out)
    assert out.value == b"ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    lib.nomp_b200_sha256_hex(b"", 0, out)
    assert out.value == b"e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
    rng = np.random.default_rng(5)
    for n in (1, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 1000, 70001):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        lib.nomp_b200_sha256_hex(data, n, out)
        assert out.value.decode() == hashlib.sha256(data).hexdigest(), n
    assert capi.jit_cache_stats() == {"knl_hits": 0, "knl_misses": 0, "cubin_hits": 0, "cubin_misses": 0}


def test_jit_cache_entry_store(tmp_path, monkeypatch):
    """src/jitcache.c without a device: directory selection, atomic put, get, and the checksum trailer that keeps a
    truncated or damaged entry away from the CUDA driver."""
    import hashlib
    lib = capi.nomp()
    monkeypatch.setenv("NOMP_JIT_CACHE", "1")
    monkeypatch.setenv("NOMP_JIT_CACHE_DIR", str(tmp_path / "a" / "b"))
    assert lib.nomp_b200_jit_cache_dir() == str(tmp_path / "a" / "b").encode() and (tmp_path / "a" / "b").is_dir()
    monkeypatch.delenv("NOMP_JIT_CACHE_DIR")
    monkeypatch.setenv("XDG_CACHE_HOME", str(tmp_path / "xdg"))
    assert lib.nomp_b200_jit_cache_dir() == str(tmp_path / "xdg" / "libnomp_b200").encode()
    monkeypatch.delenv("XDG_CACHE_HOME")
    monkeypatch.setenv("HOME", str(tmp_path / "home"))
    assert lib.nomp_b200_jit_cache_dir() == str(tmp_path / "home" / ".cache" / "libnomp_b200").encode()
    monkeypatch.setenv("NOMP_JIT_CACHE", "0")
    assert lib.nomp_b200_jit_cache_dir() is None
    assert lib.nomp_b200_jit_cache_put(b"00", b"knl", b"x", 1) != 0
    monkeypatch.setenv("NOMP_JIT_CACHE", "1")
    monkeypatch.setenv("NOMP_JIT_CACHE_DIR", str(tmp_path / "jit"))
    assert lib.nomp_b200_jit_cache_dir() == str(tmp_path / "jit").encode()

    payload = bytes(range(256)) * 33 + b"tail"
    key = hashlib.sha256(b"key").hexdigest().encode()
    n = C.c_size_t()
    buf = C.create_string_buffer(len(payload))
    assert lib.nomp_b200_jit_cache_get(key, b"cubin", buf, len(payload), C.byref(n)) == 1           # nothing there yet
    assert lib.nomp_b200_jit_cache_put(key, b"cubin", payload, len(payload)) == 0
    files = sorted(p.name for p in (tmp_path / "jit").iterdir())
    assert files == [key.decode() + ".cubin"]                                                      # no temporary left
    raw = (tmp_path / "jit" / files[0]).read_bytes()
    assert raw[:-64] == payload and raw[-64:] == hashlib.sha256(payload).hexdigest().encode()
    assert lib.nomp_b200_jit_cache_get(key, b"cubin", buf, len(payload), C.byref(n)) == 0
    assert n.value == len(payload) and buf.raw == payload
    assert lib.nomp_b200_jit_cache_get(key, b"knl", buf, len(payload), C.byref(n)) == 1             # other kind, other file
    entry = tmp_path / "jit" / files[0]
    for damaged in (raw[: len(raw) // 2], raw[:-1], raw[:100] + bytes([raw[100] ^ 1]) + raw[101:], b"", raw + b"x"):
        entry.write_bytes(damaged)
        assert lib.nomp_b200_jit_cache_get(key, b"cubin", buf, len(payload), C.byref(n)) == 1, len(damaged)
    assert lib.nomp_b200_jit_cache_put(key, b"cubin", payload, len(payload)) == 0                   # rewritten
    assert lib.nomp_b200_jit_cache_get(key, b"cubin", buf, len(payload), C.byref(n)) == 0 and buf.raw == payload
    assert lib.nomp_b200_jit_cache_put(key, b"knl", b"", 0) == 0                                    # empty payloads are fine
    assert lib.nomp_b200_jit_cache_get(key, b"knl", buf, len(payload), C.byref(n)) == 0 and n.value == 0


def test_public_headers_compile_as_c99_and_cxx17(tmp_path):
    """Every header under include/ is self-contained, pedantic C99 and valid C++ (extern "C" guards)."""
    headers = sorted(p.name for p in (ROOT / "include").glob("*.h"))
    assert {"nomp.h", "nomp-aux.h", "nomp-mem.h", "nomp-b200.h", "nompk.h"} <= set(headers)
    body = "".join(f'#include "{h}"\n' for h in headers) + "int main(void) { return 0; }\n"
    for name, cmd in (("t.c", ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror"]),
                      ("t.cpp", ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror"])):
        src = tmp_path / name
        src.write_text(body)
        r = subprocess.run(cmd + ["-I", str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / (name + ".o"))],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    # each header alone, too (no hidden dependency on include order)
    for h in headers:
        src = tmp_path / "one.c"
        src.write_text(f'#include "{h}"\nint main(void) {{ return 0; }}\n')
        r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o",
                            str(tmp_path / "one.o")], capture_output=True, text=True)
        assert r.returncode == 0, (h, r.stderr)
