"""TEST DOUBLE -- device code of the PRODUCT compiled for the host, for tests/hostdev/fake_nompk.c.

The exchange between ranks that libnompk's reduction kernels run in their last CTA (finish_result in
libnomp_b200/csrc/kernels/nompk_gridreduce.cuh: stores into the peers' exchange buffers, waits for theirs, folds in
rank order, publishes {value, sequence number}) is taken from the header as it is and wrapped in one kernel that the
cooperative emulator runs as a single warp.  fake_nompk.c calls it for nompk_reduce_peers / nompk_ax_dot_peers_f64 (after
the oracle produced this rank's value) and for nompk_allreduce_scalar (same protocol on the same buffers: reduce.cu), so
two emulated ranks in two PROCESSES exchange their values through shared memory with the product's own protocol code.

usage: build_devicecode.py <out.so>
"""
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from tests import cuda_emulation as emu  # noqa: E402

HEADER = ROOT / "libnomp_b200" / "csrc" / "kernels" / "nompk_gridreduce.cuh"

WRAPPER = r"""
template <int OP, typename T> struct HostdevOp {          // operation codes of include/nompk.h: nompk_red_op_t
  static T identity() { return OP == 1 ? T(1) : T(0); }  // only folded for lanes >= world, which finish_result ignores
  static T combine(T a, T b) { return OP == 0 ? T(a + b) : OP == 1 ? T(a * b) : OP == 2 ? (b < a ? b : a) : (b > a ? b : a); }
};
template <int OP, typename T> static void finish_as(const void *value, void *result, void *result_host, unsigned long long host_seq,
                                                    void *const *peers, int rank, int world, unsigned long long cseq) {
  nompk::PeerExchange px;
  px.peer_xchg = peers, px.rank = rank, px.world = world, px.seq = cseq;
  nompk::finish_result<HostdevOp<OP, T>, T>(*static_cast<const T *>(value), static_cast<T *>(result), static_cast<T *>(result_host),
                                            host_seq, px);
}
template <int OP> static void finish_op(int dt, const void *v, void *r, void *rh, unsigned long long hs, void *const *p, int rank, int world,
                                        unsigned long long cs) {
  switch (dt) {                                           // nompk_dtype_t
  case 0: finish_as<OP, int>(v, r, rh, hs, p, rank, world, cs); break;
  case 1: finish_as<OP, unsigned>(v, r, rh, hs, p, rank, world, cs); break;
  case 2: finish_as<OP, long long>(v, r, rh, hs, p, rank, world, cs); break;
  case 3: finish_as<OP, unsigned long long>(v, r, rh, hs, p, rank, world, cs); break;
  case 4: finish_as<OP, float>(v, r, rh, hs, p, rank, world, cs); break;
  default: finish_as<OP, double>(v, r, rh, hs, p, rank, world, cs); break;
  }
}
static void hostdev_finish(int op, int dt, const void *v, void *r, void *rh, unsigned long long hs, void **p, int rank, int world,
                           unsigned long long cs) {
  switch (op) {
  case 0: finish_op<0>(dt, v, r, rh, hs, p, rank, world, cs); break;
  case 1: finish_op<1>(dt, v, r, rh, hs, p, rank, world, cs); break;
  case 2: finish_op<2>(dt, v, r, rh, hs, p, rank, world, cs); break;
  default: finish_op<3>(dt, v, r, rh, hs, p, rank, world, cs); break;
  }
}
"""

ARGTYPES = ["int", "int", "const void *", "void *", "void *", "unsigned long long", "void **", "int", "int", "unsigned long long"]


def source():
    text = HEADER.read_text().replace('#include "nompk_common.cuh"', "#include <cstddef>").replace("#pragma once", "")
    text, n = re.subn(r'asm volatile\("mov\.u64 %0, %globaltimer;" : "=l"\((\w+)\)\);', r"\1 = nomp_emu_now_ns();", text)
    assert n == 1, "the header's inline PTX changed"
    return text + WRAPPER


if __name__ == "__main__":
    emu.cooperative_library(source(), "hostdev_finish", ARGTYPES, instance=77, out=Path(sys.argv[1]))
