"""Test helpers for generated CUDA source that work WITHOUT a GPU.

nvrtc_compile(): compile a kernel for sm_100a with NVRTC (the same call sequence the backend uses) -- proves the
                 emitted source is valid CUDA for the target.
emulate():       run a generated kernel on the host by compiling it with g++ against a tiny shim that turns
                 blockIdx / threadIdx into loop variables (threads of a block run one after the other, so only
                 kernels whose threads do not exchange data through __shared__ memory mid-kernel are meaningful).
emulate_cooperative(): the same with one coroutine per thread and real __syncthreads / warp-shuffle / atomic-ticket
                 semantics, for the kernels whose threads do cooperate: the single-pass reduction skeleton (block tree,
                 one- and two-level ticket finish, publication) and the one-block-per-element kernels with temporaries
                 in shared memory.
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


# ----------------------------------------------------------------------------------------------------------------------
# Cooperative emulation: kernels whose threads DO exchange data (shared memory + __syncthreads, warp shuffles, atomic
# tickets across blocks) run on the host with one coroutine (ucontext) per thread of a block.  A thread runs until it
# reaches a barrier or a shuffle, the scheduler releases a barrier when every live thread of the block (or of the
# warp) has arrived, blocks run one after the other.  Faithful for race-free kernels; a deadlock aborts.
# ----------------------------------------------------------------------------------------------------------------------
_COOP_SHIM = r"""
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ucontext.h>
struct nomp_emu_dim3 { unsigned x, y, z; };
struct nomp_emu_thread { ucontext_t ctx; nomp_emu_dim3 tid; int state, warp, lane, bar_id, bar_count; unsigned or_calls; char *stack; };
enum { EMU_READY = 0, EMU_BLOCK_BARRIER = 1, EMU_WARP_BARRIER = 2, EMU_DONE = 3, EMU_NAMED_BARRIER = 4 };
static nomp_emu_dim3 blockIdx, blockDim, gridDim;
static nomp_emu_thread *nomp_emu_cur;
static ucontext_t nomp_emu_sched;
static unsigned long long nomp_emu_xchg[64][32];
#define threadIdx (nomp_emu_cur->tid)
#define __global__
#define __device__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
struct alignas(16) int4 { int x, y, z, w; };
static void nomp_emu_yield(int state) {
  nomp_emu_thread *t = nomp_emu_cur;
  t->state = state;
  swapcontext(&t->ctx, &nomp_emu_sched);
}
static inline void __syncthreads() { nomp_emu_yield(EMU_BLOCK_BARRIER); }
static inline void __syncwarp(unsigned = 0xffffffffu) { nomp_emu_yield(EMU_WARP_BARRIER); }
// PTX "bar.sync id, count": the first `count` threads that arrive at barrier `id` release each other
static inline void nomp_emu_named_barrier(int id, int count) {
  nomp_emu_cur->bar_id = id, nomp_emu_cur->bar_count = count;
  nomp_emu_yield(EMU_NAMED_BARRIER);
}
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r = {x, y}; return r; }
static int nomp_emu_or[2];
static inline int __syncthreads_or(int pred) {
  const int k = (int)(nomp_emu_cur->or_calls++ & 1u);
  nomp_emu_or[k] |= pred != 0;
  nomp_emu_yield(EMU_BLOCK_BARRIER);      // everybody has contributed
  const int r = nomp_emu_or[k];
  nomp_emu_or[k ^ 1] = 0;                 // the other accumulator is idle between these two barriers
  nomp_emu_yield(EMU_BLOCK_BARRIER);
  return r;
}
#include <atomic>
#include <time.h>
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() {}
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned long long nomp_emu_now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T, typename U> static inline T atomicAdd(T *p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }
template <typename T, typename U> static inline T atomicExch(T *p, U v) { T old = *p; *p = (T)v; return old; }
template <typename T> static inline T nomp_emu_exchange(T v, int from_lane) {
  nomp_emu_thread *t = nomp_emu_cur;
  unsigned long long w = 0;
  memcpy(&w, &v, sizeof(T));
  nomp_emu_xchg[t->warp][t->lane] = w;
  nomp_emu_yield(EMU_WARP_BARRIER);                 // everybody has written
  const unsigned long long o = nomp_emu_xchg[t->warp][from_lane & 31];
  nomp_emu_yield(EMU_WARP_BARRIER);                 // everybody has read: the slots may be reused
  T r;
  memcpy(&r, &o, sizeof(T));
  return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int mask) { return nomp_emu_exchange(v, nomp_emu_cur->lane ^ mask); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane) { return nomp_emu_exchange(v, lane); }
static inline int __any_sync(unsigned, int pred) {
  int any = 0;
  for (int l = 0; l < 32; l++) any |= nomp_emu_exchange<int>(pred != 0, l);
  return any;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)nomp_emu_exchange<int>(pred != 0, l) << l;
  return m;
}
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
"""

_COOP_DRIVER = r"""
static void nomp_emu_entry() {
  NOMP_EMU_CALL;
  nomp_emu_cur->state = EMU_DONE;
  swapcontext(&nomp_emu_cur->ctx, &nomp_emu_sched);
}
extern "C" __attribute__((visibility("default"))) int nomp_emu_launch(unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz NOMP_EMU_PARAMS) {
  NOMP_EMU_STORE
  gridDim = {gx, gy, gz}; blockDim = {bx, by, bz};
  const unsigned nt = bx * by * bz;
  const size_t stack_bytes = 512 * 1024;
  nomp_emu_thread *th = (nomp_emu_thread *)calloc(nt, sizeof(nomp_emu_thread));
  for (unsigned i = 0; i < nt; i++) th[i].stack = (char *)malloc(stack_bytes);
  int rc = 0;
  for (unsigned Z = 0; Z < gz && !rc; Z++) for (unsigned Y = 0; Y < gy && !rc; Y++) for (unsigned X = 0; X < gx && !rc; X++) {
    blockIdx = {X, Y, Z};
    nomp_emu_or[0] = nomp_emu_or[1] = 0;
    for (unsigned i = 0; i < nt; i++) {
      nomp_emu_thread *t = &th[i];
      t->tid = {i % bx, (i / bx) % by, i / (bx * by)};
      t->state = EMU_READY, t->warp = (int)(i / 32), t->lane = (int)(i % 32), t->or_calls = 0;
      getcontext(&t->ctx);
      t->ctx.uc_stack.ss_sp = t->stack, t->ctx.uc_stack.ss_size = stack_bytes, t->ctx.uc_link = &nomp_emu_sched;
      makecontext(&t->ctx, nomp_emu_entry, 0);
    }
    for (;;) {
      bool progressed = false;
      unsigned live = 0, at_block = 0;
      for (unsigned i = 0; i < nt; i++) {
        if (th[i].state == EMU_READY) {
          nomp_emu_cur = &th[i];
          swapcontext(&nomp_emu_sched, &th[i].ctx);
          progressed = true;
        }
      }
      for (unsigned w = 0; w * 32 < nt; w++) {          // warp barriers
        unsigned alive = 0, waiting = 0;
        for (unsigned i = w * 32; i < nt && i < (w + 1) * 32; i++) {
          alive += th[i].state != EMU_DONE;
          waiting += th[i].state == EMU_WARP_BARRIER;
        }
        if (waiting && waiting == alive) {
          for (unsigned i = w * 32; i < nt && i < (w + 1) * 32; i++)
            if (th[i].state == EMU_WARP_BARRIER) th[i].state = EMU_READY;
          progressed = true;
        }
      }
      for (int id = 0; id < 16; id++) {                 // named barriers
        unsigned waiting = 0, count = 0;
        for (unsigned i = 0; i < nt; i++)
          if (th[i].state == EMU_NAMED_BARRIER && th[i].bar_id == id) waiting++, count = (unsigned)th[i].bar_count;
        if (waiting && waiting >= count) {
          for (unsigned i = 0; i < nt; i++)
            if (th[i].state == EMU_NAMED_BARRIER && th[i].bar_id == id) th[i].state = EMU_READY;
          progressed = true;
        }
      }
      for (unsigned i = 0; i < nt; i++) live += th[i].state != EMU_DONE, at_block += th[i].state == EMU_BLOCK_BARRIER;
      if (live == 0) break;
      if (at_block == live) {
        for (unsigned i = 0; i < nt; i++) if (th[i].state == EMU_BLOCK_BARRIER) th[i].state = EMU_READY;
        progressed = true;
      }
      if (!progressed) { fprintf(stderr, "nomp emulation: deadlock in block (%u, %u, %u)\n", X, Y, Z); rc = 1; break; }
    }
  }
  for (unsigned i = 0; i < nt; i++) free(th[i].stack);
  free(th);
  return rc;
}
"""


def emulate_cooperative(src: str, kernel: str, grid, block, argtypes, args, instance: int = 0):
    """Run a generated kernel on the host with real barrier / shuffle / ticket semantics (see above).  Different
    `instance` numbers give separate copies of the library (own globals), so several "ranks" can run at the same time
    in different Python threads and talk through host memory the way GPUs talk through peer memory."""
    so = cooperative_library(src, kernel, argtypes, instance)
    lib = C.CDLL(str(so))
    fn = lib.nomp_emu_launch
    fn.restype = C.c_int
    rc = fn(*[C.c_uint(v) for v in (*grid, *block)], *args)
    assert rc == 0, "the emulated kernel deadlocked"


def cooperative_library(src: str, kernel: str, argtypes, instance: int = 0, out: Path = None) -> Path:
    """The shared object behind emulate_cooperative(): exports
    int nomp_emu_launch(gx, gy, gz, bx, by, bz, <one argument per entry of argtypes>)."""
    key = hashlib.sha256((f"coop{instance}" + src + kernel + repr(argtypes)).encode()).hexdigest()[:16]
    so = Path(out) if out is not None else _DIR / f"c{key}.so"
    if out is not None or not so.exists():
        body = re.sub(r'extern "C"\s*', "", src)
        # the only inline PTX the bridge emits reads the global timer (time-out of the peer exchange)
        body = re.sub(r'asm volatile\("mov\.u64 %0, %globaltimer;" : "=l"\((\w+)\)\);', r"\1 = nomp_emu_now_ns();", body)
        params = "".join(f", {t} a{i}" for i, t in enumerate(argtypes))
        decls = "\n".join(f"static {t.replace('const ', '')} nomp_emu_a{i};" for i, t in enumerate(argtypes))
        store = " ".join(f"nomp_emu_a{i} = ({t.replace('const ', '')})a{i};" for i, t in enumerate(argtypes))
        call = f"{kernel}(" + ", ".join(f"nomp_emu_a{i}" for i in range(len(argtypes))) + ")"
        driver = (_COOP_DRIVER.replace("NOMP_EMU_CALL", call).replace("NOMP_EMU_PARAMS", params)
                  .replace("NOMP_EMU_STORE", store))
        cpp = _DIR / f"c{key}.cpp"
        cpp.write_text(_COOP_SHIM + body + "\n" + decls + "\n" + driver)
        # -fno-gnu-unique: static locals of templates (__shared__ arrays of templated device code) must stay private to
        # each copy of the library, or two emulated ranks in one process would share their "shared memory"
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-fno-gnu-unique",
                        "-fvisibility=hidden", "-o", str(so), str(cpp)], check=True)
    return so


# ----------------------------------------------------------------------------------------------------------------------
# The same emulation behind a driver-style launch (tests/hostdev): the kernel's arguments arrive as `void **params`, one
# pointer per parameter in declaration order, as cuLaunchKernel passes them.
# ----------------------------------------------------------------------------------------------------------------------
_KERNEL_HEAD = re.compile(r'extern\s+"C"\s+__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(([^)]*)\)')


def kernel_signature(src: str):
    """(name, [parameter types]) of the single `extern "C" __global__` function in generated CUDA source."""
    heads = _KERNEL_HEAD.findall(src)
    assert len(heads) == 1, f"expected one kernel, found {[h[0] for h in heads]}"
    name, params = heads[0]
    types = []
    for prm in (x.strip() for x in params.split(",") if x.strip()):
        m = re.match(r"(.*?)(\w+)\s*$", prm, re.S)
        types.append(m.group(1).replace("__restrict__", "").strip())
    return name, types


def build_param_launcher(src: str, out: Path):
    """Compile generated CUDA source for the host into `out` (a shared object) that exports
    nomp_emu_launch(gx, gy, gz, bx, by, bz, void **params) and nomp_hostdev_kernel_name."""
    name, types = kernel_signature(src)
    body = re.sub(r'extern "C"\s*', "", src)
    body = re.sub(r'asm volatile\("mov\.u64 %0, %globaltimer;" : "=l"\((\w+)\)\);', r"\1 = nomp_emu_now_ns();", body)
    plain = [t.replace("const ", "") for t in types]
    decls = "\n".join(f"static {t} nomp_emu_a{i};" for i, t in enumerate(plain))
    store = " ".join(f"nomp_emu_a{i} = *({t} *)nomp_emu_p[{i}];" for i, t in enumerate(plain))
    call = f"{name}(" + ", ".join(f"nomp_emu_a{i}" for i in range(len(types))) + ")"
    driver = (_COOP_DRIVER.replace("NOMP_EMU_CALL", call).replace("NOMP_EMU_PARAMS", ", void **nomp_emu_p")
              .replace("NOMP_EMU_STORE", store))
    sizes = ", ".join(f"sizeof({t})" for t in plain) or "0"
    tail = (f'\nextern "C" __attribute__((visibility("default"))) const char *nomp_hostdev_kernel_name = "{name}";\n'
            f'extern "C" __attribute__((visibility("default"))) const int nomp_hostdev_param_count = {len(plain)};\n'
            f'extern "C" __attribute__((visibility("default"))) const size_t nomp_hostdev_param_sizes[] = {{{sizes}}};\n')
    cpp = Path(str(out) + ".cpp")
    cpp.write_text(_COOP_SHIM + body + "\n" + decls + "\n" + driver + tail)
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-fno-gnu-unique",
                    "-fvisibility=hidden", "-o", str(out), str(cpp)], check=True)
