"""The grid-wide finish of the hand-written reductions (libnomp_b200/csrc/kernels/nompk_gridreduce.cuh: grid_finish,
finish_result -- shared by reduce.cu and the fused Ax + dot kernel) compiled for the HOST and executed with the
cooperative emulator of tests/cuda_emulation.py: one coroutine per thread, real barriers, shuffles and atomic tickets.
The header is taken as it is (only its one line of inline PTX, the global timer, is swapped for the host clock), so this
checks the product's device code -- ticket logic of the three grid regimes, reset of the workspace, publication order,
and the all-reduce between ranks fused into the finish -- without a GPU."""
import ctypes as C
import hashlib
import re
import subprocess
import threading
from pathlib import Path

import numpy as np
import pytest

from tests import cuda_emulation as emu

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "libnomp_b200" / "csrc" / "kernels" / "nompk_gridreduce.cuh"

KERNEL = r"""
struct Sum {
  static double identity() { return 0.0; }
  static double combine(double a, double b) { return a + b; }
};
struct MaxL {
  static long long identity() { return -(1ll << 62); }
  static long long combine(long long a, long long b) { return b > a ? b : a; }
};
// every thread sums a strided share, thread 0 folds the block through shared memory, then the grid-wide finish
template <typename Op, typename T> static void finish_kernel(const T *x, unsigned long long n, void *ws, T *result, T *result_host,
                                                              unsigned long long seq, void *const *peers, int rank, int world,
                                                              unsigned long long cseq, unsigned long long *cseq_dev) {
  __shared__ T part[256];
  T v = Op::identity();
  for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256)
    v = Op::combine(v, x[i]);
  part[threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 1; i < 256; i++) v = Op::combine(v, part[i]);
  nompk::PeerExchange px;
  px.peer_xchg = peers, px.rank = rank, px.world = world, px.seq = cseq, px.seq_dev = cseq_dev;
  nompk::grid_finish<Op, T, 256>(v, ws, result, result_host, seq, px);
}
static void sum_f64(const double *x, unsigned long long n, void *ws, double *r, double *rh, unsigned long long seq, void **p, int rank,
                    int world, unsigned long long cseq, unsigned long long *cd) { finish_kernel<Sum, double>(x, n, ws, r, rh, seq, p, rank, world, cseq, cd); }
static void max_i64(const long long *x, unsigned long long n, void *ws, long long *r, long long *rh, unsigned long long seq, void **p,
                    int rank, int world, unsigned long long cseq, unsigned long long *cd) { finish_kernel<MaxL, long long>(x, n, ws, r, rh, seq, p, rank, world, cseq, cd); }
"""


def device_source():
    text = HEADER.read_text()
    text = text.replace('#include "nompk_common.cuh"', "#include <cstddef>")
    text = text.replace("#pragma once", "")
    text, n = re.subn(r'asm volatile\("mov\.u64 %0, %globaltimer;" : "=l"\((\w+)\)\);', r"\1 = nomp_emu_now_ns();", text)
    assert n == 1, "the header's inline PTX changed: teach this test about it"
    return text + KERNEL


def run(kernel, T, x, blocks, seq, peers=None, rank=0, world=1, cseq=0, state=None, instance=0, counter=None):
    dt = {"double": np.float64, "long long": np.int64}[T]
    state = state or dict(ws=np.zeros(548928 // 8 + 8, dtype=np.uint64), res=np.zeros(1, dtype=dt), pub=np.zeros(3, dtype=np.uint64))
    ptr = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    emu.emulate_cooperative(device_source(), kernel, (blocks, 1, 1), (256, 1, 1),
                            [f"const {T} *", "unsigned long long", "void *", f"{T} *", f"{T} *", "unsigned long long", "void **", "int",
                             "int", "unsigned long long", "unsigned long long *"],
                            [ptr(x), C.c_ulonglong(x.size), ptr(state["ws"]), ptr(state["res"]), ptr(state["pub"]), C.c_ulonglong(seq),
                             ptr(peers) if peers is not None else C.c_void_p(0), C.c_int(rank), C.c_int(world), C.c_ulonglong(cseq),
                             ptr(counter) if counter is not None else C.c_void_p(0)],
                            instance=instance)
    return state


@pytest.mark.parametrize("blocks", [1, 2, 31, 32, 33, 700, 2048, 2049, 2100, 4097])
def test_grid_finish_in_every_regime(blocks):
    """One CTA (direct), up to 2048 CTAs (one ticket level), more (groups of 32 + global ticket; a last group that is
    not full): the value is the exact sum, it is published after the value with the caller's sequence number, and every
    ticket is back at zero -- twice in a row on the same workspace."""
    n = 50021
    x = (np.arange(n) * 7 % 13).astype(np.float64)
    st = None
    for seq in (5, 6):
        st = run("sum_f64", "double", x, blocks, seq, state=st)
        assert st["res"][0] == x.sum() == st["pub"].view(np.float64)[0] and st["pub"][1] == seq and st["pub"][2] == 0
        tickets = st["ws"][: (64 + 4 * 2048) // 8]
        assert not tickets.any(), "a ticket counter was left non-zero"
    y = ((np.arange(n) * 2654435761) % 100003).astype(np.int64)
    st = run("max_i64", "long long", y, blocks, 9)
    assert st["res"][0] == y.max() == st["pub"].view(np.int64)[0]


@pytest.mark.parametrize("world,blocks", [(2, 1), (2, 40), (3, 2100), (4, 7)])
def test_finish_fused_with_the_all_reduce_between_host_ranks(world, blocks):
    """finish_result with peers: `world` copies of the kernel run at the same time in `world` threads (each its own
    library instance, workspace and exchange buffer; "peer memory" is plain host memory).  Every rank ends with the
    fold of all partial sums in rank order and publishes it; three calls in a row alternate the two slots.  Odd ranks
    take the number of the call from a counter in their own "device" memory (nompk_peers_t.seq_dev: what the runtime
    uses, so that a captured launch can be replayed), even ranks get it as a launch parameter: one protocol."""
    xchg = [np.zeros(2 * world * 2, dtype=np.uint64) for _ in range(world)]
    table = np.array([b.ctypes.data for b in xchg], dtype=np.uint64)
    data = [(np.arange(3000 + 17 * r) * (3 + r) % 11).astype(np.float64) for r in range(world)]
    states = [None] * world
    counters = [np.zeros(1, dtype=np.uint64) for _ in range(world)]
    for call in (1, 2, 3):
        errors = []

        def rank_main(r):
            try:
                states[r] = run("sum_f64", "double", data[r], blocks, 100 + call, peers=table, rank=r, world=world,
                                cseq=0 if r % 2 else call, counter=counters[r] if r % 2 else None, state=states[r], instance=10 + r)
            except BaseException as exc:   # pragma: no cover
                errors.append(exc)

        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(120)
        assert not errors and not any(t.is_alive() for t in threads)
        want = data[0].sum()
        for r in range(1, world):
            want = want + data[r].sum()
        for r in range(world):
            assert states[r]["res"][0] == want == states[r]["pub"].view(np.float64)[0]
            assert states[r]["pub"][1] == 100 + call and states[r]["pub"][2] == 0
            assert counters[r][0] == (call if r % 2 else 0)


def allreduce_source():
    """The stand-alone one-shot all-reduce kernel of reduce.cu (used when a reduction could not fuse the exchange)."""
    text = (ROOT / "libnomp_b200" / "csrc" / "kernels" / "reduce.cu").read_text()
    a = text.index("template <int OP, typename T> __device__ __forceinline__ T red_identity()")
    b = text.index("// acc <- acc (op) f(x, y)")
    c = text.index("constexpr int kMaxRanks = 64;")
    d = text.index("template <int OP, typename T>\nint launch_allreduce(")
    body = ("#include <cfloat>\n#include <climits>\nenum { NOMPK_RED_SUM, NOMPK_RED_PROD, NOMPK_RED_MIN, NOMPK_RED_MAX };\n"
            + device_source().replace(KERNEL, "") + "namespace nompk {\ntemplate <typename T> struct Limits;\n"
            "template <> struct Limits<double> { static double lo() { return -INFINITY; } static double hi() { return INFINITY; } };\n"
            + text[a:b] + text[c:d] + "}\n")
    body += ("static void allreduce_sum_f64(double *value, double *host, unsigned long long host_seq, void **peers, int rank, int world,"
             " unsigned long long seq, unsigned long long *seq_dev, unsigned long long *err) {"
             " nompk::allreduce_scalar_kernel<NOMPK_RED_SUM, double>(value, host, host_seq, peers, rank, world, seq, seq_dev, err); }\n")
    return body


@pytest.mark.parametrize("world", [2, 5])
def test_stand_alone_all_reduce_kernel_between_host_ranks(world):
    xchg = [np.zeros(2 * world * 2, dtype=np.uint64) for _ in range(world)]
    table = np.array([b.ctypes.data for b in xchg], dtype=np.uint64)
    values = [np.array([1000.5 + 3 * r]) for r in range(world)]
    pubs = [np.zeros(3, dtype=np.uint64) for _ in range(world)]
    counters = [np.zeros(1, dtype=np.uint64) for _ in range(world)]
    src = allreduce_source()
    for call in (1, 2):
        contributions = [float(v[0]) for v in values]
        errors = []

        def rank_main(r):
            try:
                ptr = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
                emu.emulate_cooperative(src, "allreduce_sum_f64", (1, 1, 1), (64, 1, 1),
                                        ["double *", "double *", "unsigned long long", "void **", "int", "int", "unsigned long long",
                                         "unsigned long long *", "unsigned long long *"],
                                        [ptr(values[r]), ptr(pubs[r]), C.c_ulonglong(50 + call), ptr(table), C.c_int(r), C.c_int(world),
                                         C.c_ulonglong(0 if r % 2 else call), ptr(counters[r]) if r % 2 else C.c_void_p(0),
                                         ptr(pubs[r][2:])], instance=30 + r)
            except BaseException as exc:   # pragma: no cover
                errors.append(exc)

        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(60)
        assert not errors and not any(t.is_alive() for t in threads), errors
        want = contributions[0]
        for c_ in contributions[1:]:
            want = want + c_
        for r in range(world):
            assert values[r][0] == want == pubs[r].view(np.float64)[0] and pubs[r][1] == 50 + call and pubs[r][2] == 0
