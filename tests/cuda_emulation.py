"""Test helpers for generated CUDA source that work WITHOUT a GPU.

nvrtc_compile(): compile a kernel for sm_100a with NVRTC (the same call sequence the backend uses) -- proves the
                 emitted source is valid CUDA for the target.
emulate():       run a generated kernel on the host by compiling it with g++ against a tiny shim that turns
                 blockIdx / threadIdx into loop variables (threads of a block run one after the other, so only
                 kernels whose threads do not exchange data through __shared__ memory mid-kernel are meaningful).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import re
import subprocess
import tempfile
from pathlib import Path

_DIR = Path(tempfile.mkdtemp(prefix="nomp-cuda-emu-"))
_NVRTC = None


def _nvrtc():
    global _NVRTC
    if _NVRTC is None:
        for name in ("libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"):
            try:
                _NVRTC = C.CDLL(name)
                break
            except OSError:
                continue
        if _NVRTC is None:
            raise RuntimeError("libnvrtc not found")
    return _NVRTC


def nvrtc_compile(src: str, name: str = "kernel.cu", arch: str = "sm_100a"):
    """Returns (ok, log).  Options match backends/cuda.c."""
    lib = _nvrtc()
    prog = C.c_void_p()
    rc = lib.nvrtcCreateProgram(C.byref(prog), src.encode(), name.encode(), 0, None, None)
    assert rc == 0
    opts = [f"--gpu-architecture={arch}".encode(), b"--fmad=false", b"--std=c++17"]
    arr = (C.c_char_p * len(opts))(*opts)
    rc = lib.nvrtcCompileProgram(prog, len(opts), arr)
    size = C.c_size_t()
    lib.nvrtcGetProgramLogSize(prog, C.byref(size))
    buf = C.create_string_buffer(size.value + 1)
    lib.nvrtcGetProgramLog(prog, buf)
    lib.nvrtcDestroyProgram(C.byref(prog))
    return rc == 0, buf.value.decode(errors="replace")


_SHIM = r"""
#include <cstdint>
#include <cmath>
struct nomp_emu_dim3 { unsigned x, y, z; };
static nomp_emu_dim3 blockIdx, threadIdx, blockDim, gridDim;
#define __global__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
static inline void __syncthreads() {}
struct int4 { int x, y, z, w; };
"""


def emulate(src: str, kernel: str, grid, block, argtypes, args):
    """Compile `src` for the host and call `kernel` once per (block, thread).  argtypes: C type strings."""
    key = hashlib.sha256((src + kernel + repr(argtypes)).encode()).hexdigest()[:16]
    so = _DIR / f"e{key}.so"
    if not so.exists():
        params = ", ".join(f"{t} a{i}" for i, t in enumerate(argtypes))
        call = ", ".join(f"a{i}" for i in range(len(argtypes)))
        body = re.sub(r'extern "C"\s*', "", src)
        driver = f"""
extern "C" void nomp_emu_launch(unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, {params}) {{
  gridDim = {{gx, gy, gz}}; blockDim = {{bx, by, bz}};
  for (unsigned Z = 0; Z < gz; Z++) for (unsigned Y = 0; Y < gy; Y++) for (unsigned X = 0; X < gx; X++)
    for (unsigned z = 0; z < bz; z++) for (unsigned y = 0; y < by; y++) for (unsigned x = 0; x < bx; x++) {{
      blockIdx = {{X, Y, Z}}; threadIdx = {{x, y, z}};
      {kernel}({call});
    }}
}}
"""
        cpp = _DIR / f"e{key}.cpp"
        cpp.write_text(_SHIM + body + driver)
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", str(so), str(cpp)],
                       check=True)
    lib = C.CDLL(str(so))
    fn = lib.nomp_emu_launch
    fn.restype = None
    fn(*[C.c_uint(v) for v in (*grid, *block)], *args)
