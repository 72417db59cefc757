// Gather-scatter ("direct stiffness summation", the gs_op of Nekbone / gslib): every degree of freedom that carries
// the same global id -- on this GPU or on another rank -- ends up holding the combination (sum, product, min, max) of
// all its copies.  It is the step either side of the local Ax in a real spectral-element solve; the
// reference has no such operator.
//
// Data layout in HBM (built once per numbering by the setup entry points, all on the device):
//   groups     one per global id that needs work here: at least two local copies, or shared with another rank.
//              The groups shared with another rank come first (their partial results travel while the rest is
//              folded, and "shared" is a comparison with Q), each part ordered by the position of its first local
//              copy, so neighbouring threads touch neighbouring memory.
//   offsets    unsigned[G+1], indices unsigned[nnz]: CSR of the local copies of each group, ascending.
//   remote     for the Q groups shared with other ranks: the group, and a CSR of (peer rank, position) pairs.  The
//              position indexes the list S(me, peer) of ids the two ranks share, in ascending id order -- both
//              sides derive the same list, so a position means the same id on both.
//   exchange   one buffer per rank, mapped into every peer (CUDA IPC over NVLink): arrival flags, then two slots
//              (alternating by call parity) of 8-byte values, one segment per peer.
//
//   warp layout  (built when no group has more than 32 local copies, i.e. for every mesh numbering): the copies of the
//              groups once more as `pidx`, padded so that no group straddles a multiple of 32, with one bit mask per
//              32 slots that marks the first copy of each group.  gs_local_warp_kernel works on it with one COPY per
//              lane -- see there.
//
// apply = two launches.  gs_local_kernel folds the local copies of each group in ascending index order; local-only
// groups are written back at once; for shared groups the partial is kept and stored straight into the peers'
// exchange buffers (P2P stores over NVLink, no staging copy, no NCCL), and the last of the CTAs that hold shared groups
// raises this rank's flag on every neighbour.  gs_remote_kernel waits for the neighbours' flags, folds own and received partials in
// ascending RANK order (same order on every rank: bit-identical results everywhere) and writes the copies back.
// With one rank the second launch does not happen.
#include <cub/cub.cuh>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "nompk_common.cuh"

struct nompk_gs {
  size_t n = 0, n_unique = 0;
  long long *unique_ids = nullptr;  // [U] ascending
  unsigned *run_count = nullptr;    // [U] local copies of each unique id
  unsigned *run_start = nullptr;    // [U+1]
  unsigned *sorted_idx = nullptr;   // [n] local indices ordered by (id, index)
  int world = 1, rank = 0;
  std::vector<unsigned *> peer_pos;  // per rank: device [U], position + 1 in S(me, r), 0 = not shared with r
  std::vector<size_t> shared;        // per rank: |S(me, r)|
  std::vector<size_t> recv_off, send_off;
  bool finalized = false;
  size_t G = 0, nnz = 0, Q = 0, R = 0, total_shared = 0;
  unsigned remote_ctas = 0;  // CTAs of gs_local_kernel (kGsThreads consecutive groups each) that hold a shared group
  unsigned *offsets = nullptr, *indices = nullptr;
  // warp layout (nwarps == 0: not built, gs_local_kernel is used)
  size_t nwarps = 0;
  unsigned *pidx = nullptr, *heads = nullptr, *wgroup = nullptr;
  unsigned short *rowinfo = nullptr;   // per row of 32 slots: longest group | occupied slots << 8
  unsigned remote_ctas_warp = 0;
  int rows = 4;                      // rows of 32 slots per warp of gs_local_warp_kernel
  bool force_group_kernel = false;   // NOMPK_GS_KERNEL=group: profiling / tests of the one-group-per-thread kernel
  int *remote_slot = nullptr;
  unsigned *rgroup = nullptr, *roffsets = nullptr, *rpos = nullptr;
  int *rpeer = nullptr;
  unsigned long long *partial = nullptr;
  unsigned *ticket = nullptr;
  size_t *d_recv_off = nullptr, *d_send_off = nullptr;
  void **d_peer_xchg = nullptr;
  int *d_neighbours = nullptr;
  int n_neighbours = 0;
  bool connected = false;
  unsigned long long seq = 0;
};

namespace nompk {
namespace {

constexpr int kGsThreads = 256;
constexpr unsigned long long kGsTimeoutNs = 20ull * 1000 * 1000 * 1000;  // 20 s

__device__ __forceinline__ unsigned long long timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int OP, typename T> __device__ __forceinline__ T combine(T a, T b) {
  if constexpr (OP == NOMPK_RED_SUM) return op_add(a, b);
  if constexpr (OP == NOMPK_RED_PROD) return op_mul(a, b);
  if constexpr (OP == NOMPK_RED_MIN) return b < a ? b : a;
  if constexpr (OP == NOMPK_RED_MAX) return b > a ? b : a;
  return a;
}

size_t flags_bytes(int world) { return ((size_t)2 * world * sizeof(unsigned long long) + 255) / 256 * 256; }

// ---- setup kernels ----------------------------------------------------------------------------------------------
__global__ void iota_kernel(unsigned *out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (unsigned)i;
}

// found[u] = 1 if unique_ids[u] (> 0) also occurs in the peer's ascending list
__global__ void match_kernel(const long long *__restrict__ ids, size_t n_ids, const long long *__restrict__ peer,
                             size_t n_peer, unsigned *__restrict__ found) {
  const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_ids) return;
  const long long id = ids[u];
  unsigned hit = 0;
  if (id > 0 && n_peer > 0) {
    size_t lo = 0, hi = n_peer;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (peer[mid] < id) lo = mid + 1;
      else hi = mid;
    }
    hit = lo < n_peer && peer[lo] == id;
  }
  found[u] = hit;
}

// pos[u] <- found ? exclusive position + 1 : 0   (pos holds the exclusive scan of found on entry)
__global__ void position_kernel(const unsigned *__restrict__ found, unsigned *__restrict__ pos, size_t n) {
  const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n) pos[u] = found[u] ? pos[u] + 1 : 0;
}

// active[u] = needs work; rcount[u] = number of peers sharing it
__global__ void classify_kernel(const long long *__restrict__ ids, const unsigned *__restrict__ count,
                                unsigned *const *__restrict__ peer_pos, int world, size_t n,
                                unsigned char *__restrict__ active, unsigned *__restrict__ rcount) {
  const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  unsigned rc = 0;
  for (int r = 0; r < world; r++)
    if (peer_pos[r] && peer_pos[r][u]) rc++;
  const bool on = ids[u] > 0 && (count[u] >= 2 || rc > 0);
  active[u] = on;
  rcount[u] = on ? rc : 0;
}

// Sort key of a group: groups shared with a peer first, then by first local copy.  The shared groups then sit in the
// first CTAs of gs_local_kernel, which the hardware schedules first: this rank's partial results are in the neighbours'
// buffers and its flag is up a few microseconds into the kernel instead of at its end, and a neighbour's gs_remote_kernel
// no longer waits for the tail of this rank's local kernel (0.14 ms against 0.07 ms ideal at 8 ranks in round 1).
__global__ void first_index_kernel(const unsigned *__restrict__ sel, const unsigned *__restrict__ run_start,
                                   const unsigned *__restrict__ sorted_idx, const unsigned *__restrict__ rcount,
                                   unsigned long long *__restrict__ first, size_t G) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) first[g] = ((rcount[sel[g]] > 0 ? 0ull : 1ull) << 32) | sorted_idx[run_start[sel[g]]];
}

__global__ void group_sizes_kernel(const unsigned *__restrict__ order, const unsigned *__restrict__ count,
                                   const unsigned *__restrict__ rcount, unsigned *__restrict__ cnt,
                                   unsigned *__restrict__ rcnt, unsigned *__restrict__ rflag, size_t G) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const unsigned u = order[g];
  cnt[g] = count[u], rcnt[g] = rcount[u], rflag[g] = rcount[u] > 0;
}

// offsets / rstart / rslot hold exclusive scans on entry
__global__ void fill_kernel(const unsigned *__restrict__ order, const unsigned *__restrict__ run_start,
                            const unsigned *__restrict__ count, const unsigned *__restrict__ sorted_idx,
                            const unsigned *__restrict__ offsets, const unsigned *__restrict__ rcnt,
                            const unsigned *__restrict__ rstart, const unsigned *__restrict__ rslot,
                            unsigned *const *__restrict__ peer_pos, int world, unsigned *__restrict__ indices,
                            int *__restrict__ remote_slot, unsigned *__restrict__ rgroup,
                            unsigned *__restrict__ roffsets, int *__restrict__ rpeer, unsigned *__restrict__ rpos,
                            size_t G) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const unsigned u = order[g], b = run_start[u], c = count[u], o = offsets[g];
  for (unsigned j = 0; j < c; j++) indices[o + j] = sorted_idx[b + j];
  if (rcnt[g] == 0) {
    remote_slot[g] = -1;
    return;
  }
  const unsigned q = rslot[g];
  remote_slot[g] = (int)q, rgroup[q] = (unsigned)g, roffsets[q] = rstart[g];
  unsigned k = rstart[g];
  for (int r = 0; r < world; r++) {
    const unsigned p = peer_pos[r] ? peer_pos[r][u] : 0;
    if (p) rpeer[k] = r, rpos[k] = p - 1, k++;
  }
}

// blk[b] = 1 if CTA b of gs_local_kernel (groups b * threads .. ) holds a group shared with a peer
__global__ void mark_remote_ctas_kernel(const int *__restrict__ remote_slot, size_t G, int threads, unsigned *__restrict__ blk) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G && remote_slot[g] >= 0) blk[g / threads] = 1u;
}

// ---- warp layout ------------------------------------------------------------------------------------------------
// The copies of the groups, in group order, placed so that no group straddles a multiple of 32 slots (a group that
// would is moved to the next multiple; the slots in between stay empty = 0xffffffff).  One thread lays out a tile of
// kLayoutTile consecutive groups (a sequential recurrence, 256 short steps); tiles start at multiples of 32.
constexpr int kLayoutTile = 256;
constexpr int kGsWarps = 256 / 32;   // warps per CTA of gs_local_warp_kernel (kGsThreads / 32)
constexpr int kGsRowsDefault = 4;    // rows of 32 slots per warp (NOMPK_GS_ROWS=8 at setup time selects eight)

__global__ void layout_size_kernel(const unsigned *__restrict__ offsets, size_t G, unsigned *__restrict__ tile_warps) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t g0 = t * kLayoutTile;
  if (g0 >= G) return;
  const size_t g1 = g0 + kLayoutTile < G ? g0 + kLayoutTile : G;
  unsigned pos = 0;
  for (size_t g = g0; g < g1; g++) {
    const unsigned m = offsets[g + 1] - offsets[g];
    if ((pos & 31u) + m > 32u) pos = (pos + 31u) & ~31u;
    pos += m;
  }
  tile_warps[t] = (pos + 31u) / 32u;
}

// tile_start holds the exclusive scan of tile_warps on entry; pidx is 0xff-filled, heads zero-filled
__global__ void layout_fill_kernel(const unsigned *__restrict__ offsets, const unsigned *__restrict__ indices, size_t G,
                                   const unsigned *__restrict__ tile_start, unsigned *__restrict__ pidx,
                                   unsigned *__restrict__ heads, unsigned *__restrict__ wgroup,
                                   unsigned short *__restrict__ rowinfo) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t g0 = t * kLayoutTile;
  if (g0 >= G) return;
  const size_t g1 = g0 + kLayoutTile < G ? g0 + kLayoutTile : G;
  const size_t base = (size_t)tile_start[t] * 32;
  unsigned pos = 0;
  for (size_t g = g0; g < g1; g++) {
    const unsigned b = offsets[g], m = offsets[g + 1] - b;
    if ((pos & 31u) + m > 32u) pos = (pos + 31u) & ~31u;
    const size_t at = base + pos;
    for (unsigned c = 0; c < m; c++) pidx[at + c] = indices[b + c];
    if ((at & 31) == 0) wgroup[at / 32] = (unsigned)g;   // every 32 slots start with the first copy of a group
    heads[at / 32] |= 1u << (at & 31);                    // (the tile's slots belong to this thread alone)
    const unsigned longest = rowinfo[at / 32] & 0xffu;    // row summary: longest group, occupied slots
    rowinfo[at / 32] = (unsigned short)((longest > m ? longest : m) | (((unsigned)(at & 31) + m) << 8));
    pos += m;
  }
}

// blk[c] = 1 if CTA c of gs_local_warp_kernel holds a group shared with a peer (the shared groups are the first Q)
__global__ void mark_remote_warp_ctas_kernel(const unsigned *__restrict__ wgroup, size_t nwarps, size_t Q, int rows_per_cta,
                                             unsigned *__restrict__ blk) {
  const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < nwarps && wgroup[w] < Q) blk[w / rows_per_cta] = 1u;
}

// ---- apply ----------------------------------------------------------------------------------------------------
struct GsView {
  const unsigned *offsets, *indices;
  const unsigned *pidx, *heads, *wgroup;   // warp layout
  const unsigned short *rowinfo;
  size_t nwarps;
  const int *remote_slot;
  const unsigned *rgroup, *roffsets, *rpos;
  const int *rpeer;
  unsigned long long *partial;
  unsigned *ticket;
  const size_t *recv_off, *send_off;
  void *const *peer_xchg;
  const int *neighbours;
  size_t G, Q, values_base;  // values_base: byte offset of this call's slot in THIS rank's exchange buffer
  unsigned remote_ctas;      // how many CTAs of gs_local_kernel hold shared groups (they take the ticket)
  size_t flags_bytes;        // size of the flag block in front of the values of every exchange buffer
  int n_neighbours, rank, world, slot;
  unsigned long long seq;
  unsigned long long *error_host;
};

// One group per thread.  Two other schedules were measured on a box mesh of 1.3e8 points (0.58 ms as written) and
// dropped: plain local pairs moved to the front and handled two per thread from one 16-byte index load (0.66 ms: the
// pairs and the edge / vertex groups of a cache line are then swept in two passes, DRAM traffic 3.5 GB instead of
// 2.5 GB), and four groups per thread with all loads of a level issued together (0.61 ms).  The traffic is what it
// has to be -- faces normal to the fastest index touch both 32-byte sectors of every 64-byte line, so the whole
// vector is read and written once (ncu: 1.51 GB read, 1.01 GB written for a 1.07 GB vector) -- at 4.35 TB/s.
template <int OP, typename T> __global__ void __launch_bounds__(kGsThreads) gs_local_kernel(T *__restrict__ v, GsView s) {
  const size_t g = (size_t)blockIdx.x * kGsThreads + threadIdx.x;
  bool remote_here = false;
  if (g < s.G) {
    const unsigned b = s.offsets[g], e = s.offsets[g + 1];
    T acc = v[s.indices[b]];
    for (unsigned j = b + 1; j < e; j++) acc = combine<OP, T>(acc, v[s.indices[j]]);
    const int q = s.Q ? s.remote_slot[g] : -1;
    if (q < 0) {
      for (unsigned j = b; j < e; j++) v[s.indices[j]] = acc;
    } else {
      unsigned long long word = 0;
      memcpy(&word, &acc, sizeof(T));
      s.partial[q] = word;
      for (unsigned k = s.roffsets[q]; k < s.roffsets[q + 1]; k++) {
        const int r = s.rpeer[k];
        // the slot of this call in the PEER's buffer: its stride is the peer's number of shared ids, not ours
        char *dst = static_cast<char *>(s.peer_xchg[r]) + s.flags_bytes + (s.send_off[s.slot * s.world + r] + s.rpos[k]) * 8;
        *reinterpret_cast<volatile unsigned long long *>(dst) = word;
      }
      __threadfence_system();
      remote_here = true;
    }
  }
  if (s.Q == 0) return;
  // This rank's values are complete on every neighbour once all CTAs that hold shared groups have passed here; the
  // last of THEM raises the flags.  (One ticket per CTA of the whole grid would serialise 10^4 - 10^5 atomics on one
  // address: tens of microseconds for a kernel of a few hundred.)
  if (!__syncthreads_or(remote_here)) return;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(s.ticket, 1u);
    if (t == s.remote_ctas - 1) {
      *s.ticket = 0;
      __threadfence_system();
      for (int i = 0; i < s.n_neighbours; i++) {
        unsigned long long *flags = static_cast<unsigned long long *>(s.peer_xchg[s.neighbours[i]]);
        *reinterpret_cast<volatile unsigned long long *>(flags + (size_t)s.slot * s.world + s.rank) = s.seq;
      }
    }
  }
}

// One COPY per lane instead of one group per thread.  gs_local_kernel walks offsets -> indices -> values: three loads that
// each wait for the one before, 40 warp-cycles of long-scoreboard stall per issued instruction, DRAM half idle (ncu,
// profiles/r02_ncu_summary.md).  Here a warp reads kGsRows rows of 32 consecutive slots of `pidx` (coalesced loads that
// depend on nothing), gathers the values, and folds every group inside its row with shuffles: two dependent loads, no
// offsets, and the index traffic is one 128-byte line per row.  0.504 ms against 0.580 ms on 1.3e8 points.  The fold is SERIAL in ascending copy order, led by the
// lane of a group's first copy (head), so the bits are those of gs_local_kernel and of the oracle: step s combines
// what the lane s places further on holds, for the heads whose group is longer than s.  Groups shared with a peer are
// the first Q groups (first_index_kernel), so "shared" is a comparison and the slot of the partial is the group number.
template <int OP, typename T, int kGsRows> __global__ void __launch_bounds__(kGsThreads) gs_local_warp_kernel(T *__restrict__ v, GsView s) {
  const unsigned lane = threadIdx.x & 31u;
  // a warp takes kGsRows consecutive rows of 32 slots: all index loads, then all value loads are in flight together
  const size_t w0 = ((size_t)blockIdx.x * kGsWarps + (threadIdx.x >> 5)) * kGsRows;
  bool remote_here = false;
  unsigned idx[kGsRows], hm[kGsRows], info[kGsRows];
  T val[kGsRows];
#pragma unroll
  for (int r = 0; r < kGsRows; r++) {
    const bool on = w0 + r < s.nwarps;   // warp-uniform
    idx[r] = on ? s.pidx[(w0 + r) * 32 + lane] : 0xffffffffu;
    hm[r] = on ? s.heads[w0 + r] : 1u;
    info[r] = on ? s.rowinfo[w0 + r] : 0x0101u;
  }
#pragma unroll
  for (int r = 0; r < kGsRows; r++) {
    val[r] = T(0);
    if (idx[r] != 0xffffffffu) val[r] = v[idx[r]];   // empty slots are the last slots of a row
  }
  const unsigned le = 0xffffffffu >> (31u - lane);                        // lanes up to and including this one
#pragma unroll
  for (int r = 0; r < kGsRows; r++) {
    if (w0 + r >= s.nwarps) break;                                        // warp-uniform
    const bool valid = idx[r] != 0xffffffffu;
    const int maxlen = (int)(info[r] & 0xffu), occupied = (int)(info[r] >> 8);   // the same for the whole row
    const unsigned vm = occupied >= 32 ? 0xffffffffu : (1u << occupied) - 1u;
    const int head = 31 - __clz((int)(hm[r] & le));                       // first copy of this lane's group (slot 0 is a head)
    const bool is_head = valid && ((hm[r] >> lane) & 1u);
    const unsigned after = (hm[r] & ~le) | ~vm;                           // the next head, or the first empty slot
    const int len = (after ? __ffs((int)after) - 1 : 32) - (int)lane;     // copies of the group, for its head
    T acc = val[r];
    for (int step = 1; step < maxlen; step++) {
      const T x = __shfl_sync(0xffffffffu, val[r], (int)((lane + step) & 31u));
      if (is_head && step < len) acc = combine<OP, T>(acc, x);
    }
    const T res = __shfl_sync(0xffffffffu, acc, head);
    const size_t g = (size_t)s.wgroup[w0 + r] + (unsigned)__popc(hm[r] & ((1u << head) - 1u));
    const bool shared = g < s.Q;
    if (valid && !shared) v[idx[r]] = res;
    if (is_head && shared) {
      unsigned long long word = 0;
      memcpy(&word, &acc, sizeof(T));
      s.partial[g] = word;
      for (unsigned k = s.roffsets[g]; k < s.roffsets[g + 1]; k++) {
        const int rk = s.rpeer[k];
        char *dst = static_cast<char *>(s.peer_xchg[rk]) + s.flags_bytes + (s.send_off[s.slot * s.world + rk] + s.rpos[k]) * 8;
        *reinterpret_cast<volatile unsigned long long *>(dst) = word;
      }
      __threadfence_system();
      remote_here = true;
    }
  }
  if (s.Q == 0) return;
  if (!__syncthreads_or(remote_here)) return;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(s.ticket, 1u);
    if (t == s.remote_ctas - 1) {
      *s.ticket = 0;
      __threadfence_system();
      for (int i = 0; i < s.n_neighbours; i++) {
        unsigned long long *flags = static_cast<unsigned long long *>(s.peer_xchg[s.neighbours[i]]);
        *reinterpret_cast<volatile unsigned long long *>(flags + (size_t)s.slot * s.world + s.rank) = s.seq;
      }
    }
  }
}

template <int OP, typename T> __global__ void __launch_bounds__(kGsThreads) gs_remote_kernel(T *__restrict__ v, GsView s) {
  // wait until every neighbour has delivered call `seq` (a peer that never arrives must not hang the GPU)
  if (threadIdx.x < s.n_neighbours) {
    const volatile unsigned long long *flag = static_cast<const unsigned long long *>(s.peer_xchg[s.rank]) +
                                              (size_t)s.slot * s.world + s.neighbours[threadIdx.x];
    const unsigned long long t0 = timer_ns();
    while (*flag != s.seq) {
      if (timer_ns() - t0 > kGsTimeoutNs) {
        if (s.error_host) *reinterpret_cast<volatile unsigned long long *>(s.error_host) = s.seq;
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  const size_t q = (size_t)blockIdx.x * kGsThreads + threadIdx.x;
  if (q >= s.Q) return;
  const char *mine = static_cast<const char *>(s.peer_xchg[s.rank]) + s.values_base;
  T own;
  {
    const unsigned long long word = s.partial[q];
    memcpy(&own, &word, sizeof(T));
  }
  T acc = own;
  bool first = true, own_done = false;
  for (unsigned k = s.roffsets[q]; k < s.roffsets[q + 1]; k++) {
    const int r = s.rpeer[k];
    if (!own_done && r > s.rank) {
      acc = first ? own : combine<OP, T>(acc, own);
      first = false, own_done = true;
    }
    const unsigned long long word =
        *reinterpret_cast<const volatile unsigned long long *>(mine + (s.recv_off[r] + s.rpos[k]) * 8);
    T x;
    memcpy(&x, &word, sizeof(T));
    acc = first ? x : combine<OP, T>(acc, x);
    first = false;
  }
  if (!own_done) acc = first ? own : combine<OP, T>(acc, own);
  const unsigned g = s.rgroup[q];
  for (unsigned j = s.offsets[g]; j < s.offsets[g + 1]; j++) v[s.indices[j]] = acc;
}

template <int OP, typename T> int launch_gs(nompk_gs *gs, void *v, unsigned long long *error_host, cudaStream_t stream) {
  if (gs->G == 0) return NOMPK_OK;
  gs->seq++;
  GsView s;
  s.offsets = gs->offsets, s.indices = gs->indices, s.remote_slot = gs->remote_slot;
  s.rgroup = gs->rgroup, s.roffsets = gs->roffsets, s.rpos = gs->rpos, s.rpeer = gs->rpeer;
  s.partial = gs->partial, s.ticket = gs->ticket, s.recv_off = gs->d_recv_off, s.send_off = gs->d_send_off;
  s.peer_xchg = gs->d_peer_xchg, s.neighbours = gs->d_neighbours;
  const bool warp_layout = gs->nwarps > 0 && !gs->force_group_kernel;
  s.pidx = gs->pidx, s.heads = gs->heads, s.wgroup = gs->wgroup, s.rowinfo = gs->rowinfo, s.nwarps = gs->nwarps;
  s.G = gs->G, s.Q = gs->Q, s.remote_ctas = warp_layout ? gs->remote_ctas_warp : gs->remote_ctas;
  const unsigned blocks = warp_layout ? (unsigned)((gs->nwarps + kGsWarps * gs->rows - 1) / (kGsWarps * gs->rows))
                                      : (unsigned)((gs->G + kGsThreads - 1) / kGsThreads);
  s.slot = (int)(gs->seq & 1ull);
  s.flags_bytes = flags_bytes(gs->world);
  s.values_base = s.flags_bytes + (size_t)s.slot * gs->total_shared * 8;
  s.n_neighbours = gs->n_neighbours, s.rank = gs->rank, s.world = gs->world;
  s.seq = gs->seq, s.error_host = error_host;
  if (warp_layout && gs->rows == 8) gs_local_warp_kernel<OP, T, 8><<<blocks, kGsThreads, 0, stream>>>(static_cast<T *>(v), s);
  else if (warp_layout) gs_local_warp_kernel<OP, T, kGsRowsDefault><<<blocks, kGsThreads, 0, stream>>>(static_cast<T *>(v), s);
  else gs_local_kernel<OP, T><<<blocks, kGsThreads, 0, stream>>>(static_cast<T *>(v), s);
  NOMPK_LAUNCH_CHECK("gs_local_kernel");
  if (gs->Q > 0) {
    gs_remote_kernel<OP, T><<<(unsigned)((gs->Q + kGsThreads - 1) / kGsThreads), kGsThreads, 0, stream>>>(static_cast<T *>(v), s);
    NOMPK_LAUNCH_CHECK("gs_remote_kernel");
  }
  return NOMPK_OK;
}

typedef int (*gs_fn)(nompk_gs *, void *, unsigned long long *, cudaStream_t);

template <int OP> gs_fn pick_gs(nompk_dtype_t dt) {
  switch (dt) {
  case NOMPK_I32: return launch_gs<OP, int>;
  case NOMPK_U32: return launch_gs<OP, unsigned>;
  case NOMPK_I64: return launch_gs<OP, long long>;
  case NOMPK_U64: return launch_gs<OP, unsigned long long>;
  case NOMPK_F32: return launch_gs<OP, float>;
  case NOMPK_F64: return launch_gs<OP, double>;
  }
  return nullptr;
}

template <typename T> int dev_alloc(T **p, size_t count) {
  *p = nullptr;
  NOMPK_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(p), (count ? count : 1) * sizeof(T)));
  return NOMPK_OK;
}

unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }

// exclusive prefix sum of unsigned values, out[n] = total; returns the total on the host
int exclusive_sum(const unsigned *in, unsigned *out, size_t n, size_t *total, cudaStream_t stream) {
  *total = 0;
  if (n == 0) {
    NOMPK_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(unsigned), stream));
    return NOMPK_OK;
  }
  void *tmp = nullptr;
  size_t bytes = 0;
  NOMPK_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, stream));
  NOMPK_CUDA_TRY(cudaMalloc(&tmp, bytes ? bytes : 1));
  unsigned last_in = 0, last_out = 0;
  NOMPK_CUDA_TRY(cudaMemcpyAsync(&last_in, in + n - 1, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
  NOMPK_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, stream));
  NOMPK_CUDA_TRY(cudaMemcpyAsync(&last_out, out + n - 1, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
  NOMPK_CUDA_TRY(cudaStreamSynchronize(stream));
  NOMPK_CUDA_TRY(cudaFree(tmp));
  *total = (size_t)last_in + last_out;
  const unsigned t = (unsigned)*total;
  NOMPK_CUDA_TRY(cudaMemcpyAsync(out + n, &t, sizeof(unsigned), cudaMemcpyHostToDevice, stream));
  NOMPK_CUDA_TRY(cudaStreamSynchronize(stream));
  return NOMPK_OK;
}

}  // namespace
}  // namespace nompk

using namespace nompk;

extern "C" void nompk_gs_destroy(nompk_gs_t *gs) {
  if (!gs) return;
  for (unsigned *p : gs->peer_pos) cudaFree(p);
  void *ptrs[] = {gs->unique_ids, gs->run_count, gs->run_start, gs->sorted_idx, gs->offsets, gs->indices, gs->pidx, gs->heads, gs->wgroup, gs->rowinfo,
                  gs->remote_slot, gs->rgroup, gs->roffsets, gs->rpos, gs->rpeer, gs->partial, gs->ticket,
                  gs->d_recv_off, gs->d_send_off, gs->d_peer_xchg, gs->d_neighbours};
  for (void *p : ptrs) cudaFree(p);
  delete gs;
}

extern "C" int nompk_gs_create(const long long *ids, size_t n, nompk_gs_t **out, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!out || (n && !ids) || n >= 0xffffffffull) {
    set_error("nompk_gs_create: bad arguments (n = %zu; at most 2^32 - 2 local degrees of freedom)", n);
    return NOMPK_EINVAL;
  }
  nompk_gs *gs = new nompk_gs;
  gs->n = n;
  *out = gs;
  long long *keys = nullptr, *uniq = nullptr;
  unsigned *vals = nullptr, *counts = nullptr, *nruns = nullptr;
  void *tmp = nullptr;
  int err = NOMPK_OK;
  auto body = [&]() -> int {
    if (int e = dev_alloc(&keys, n)) return e;
    if (int e = dev_alloc(&vals, n)) return e;
    if (int e = dev_alloc(&gs->sorted_idx, n)) return e;
    if (int e = dev_alloc(&uniq, n)) return e;
    if (int e = dev_alloc(&counts, n)) return e;
    if (int e = dev_alloc(&nruns, 1)) return e;
    size_t U = 0;
    if (n > 0) {
      iota_kernel<<<blocks_for(n), 256, 0, stream>>>(vals, n);
      NOMPK_LAUNCH_CHECK("iota_kernel");
      size_t b1 = 0, b2 = 0;
      NOMPK_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, b1, ids, keys, vals, gs->sorted_idx, n, 0, 64, stream));
      NOMPK_CUDA_TRY(cub::DeviceRunLengthEncode::Encode(nullptr, b2, keys, uniq, counts, nruns, n, stream));
      const size_t bytes = b1 > b2 ? b1 : b2;
      NOMPK_CUDA_TRY(cudaMalloc(&tmp, bytes ? bytes : 1));
      b1 = b2 = bytes;
      // LSD radix sort is stable: equal ids keep ascending local index
      NOMPK_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, b1, ids, keys, vals, gs->sorted_idx, n, 0, 64, stream));
      NOMPK_CUDA_TRY(cub::DeviceRunLengthEncode::Encode(tmp, b2, keys, uniq, counts, nruns, n, stream));
      unsigned h = 0;
      NOMPK_CUDA_TRY(cudaMemcpyAsync(&h, nruns, sizeof(h), cudaMemcpyDeviceToHost, stream));
      NOMPK_CUDA_TRY(cudaStreamSynchronize(stream));
      U = h;
    }
    gs->n_unique = U;
    if (int e = dev_alloc(&gs->unique_ids, U)) return e;
    if (int e = dev_alloc(&gs->run_count, U)) return e;
    if (int e = dev_alloc(&gs->run_start, U + 1)) return e;
    NOMPK_CUDA_TRY(cudaMemcpyAsync(gs->unique_ids, uniq, U * sizeof(long long), cudaMemcpyDeviceToDevice, stream));
    NOMPK_CUDA_TRY(cudaMemcpyAsync(gs->run_count, counts, U * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream));
    size_t total = 0;
    if (int e = exclusive_sum(gs->run_count, gs->run_start, U, &total, stream)) return e;
    return NOMPK_OK;
  };
  err = body();
  cudaStreamSynchronize(stream);
  cudaFree(keys), cudaFree(vals), cudaFree(uniq), cudaFree(counts), cudaFree(nruns), cudaFree(tmp);
  if (err) {
    nompk_gs_destroy(gs);
    *out = nullptr;
  }
  return err;
}

extern "C" int nompk_gs_unique(const nompk_gs_t *gs, const long long **ids, size_t *count) {
  if (!gs || !ids || !count) return NOMPK_EINVAL;
  *ids = gs->unique_ids, *count = gs->n_unique;
  return NOMPK_OK;
}

extern "C" int nompk_gs_match_peer(nompk_gs_t *gs, int peer, int world, const long long *peer_ids, size_t peer_count,
                                   size_t *n_shared, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!gs || gs->finalized || peer < 0 || peer >= world || world < 1 || world > 64 || (peer_count && !peer_ids)) {
    set_error("nompk_gs_match_peer: bad arguments (peer %d of %d)", peer, world);
    return NOMPK_EINVAL;
  }
  if (gs->peer_pos.empty()) gs->peer_pos.assign(world, nullptr), gs->shared.assign(world, 0);
  if ((int)gs->peer_pos.size() != world) {
    set_error("nompk_gs_match_peer: world size changed during setup");
    return NOMPK_EINVAL;
  }
  const size_t U = gs->n_unique;
  unsigned *found = nullptr, *pos = nullptr;
  if (int e = dev_alloc(&found, U)) return e;
  if (int e = dev_alloc(&pos, U + 1)) return e;
  size_t total = 0;
  int err = NOMPK_OK;
  if (U > 0) {
    match_kernel<<<blocks_for(U), 256, 0, stream>>>(gs->unique_ids, U, peer_ids, peer_count, found);
    err = exclusive_sum(found, pos, U, &total, stream);
    if (!err) position_kernel<<<blocks_for(U), 256, 0, stream>>>(found, pos, U);
    if (!err && cudaGetLastError() != cudaSuccess) err = NOMPK_ECUDA;
  }
  cudaStreamSynchronize(stream);
  cudaFree(found);
  if (err) {
    cudaFree(pos);
    set_error("nompk_gs_match_peer: device failure");
    return err;
  }
  cudaFree(gs->peer_pos[peer]);
  if (total == 0) cudaFree(pos), pos = nullptr;
  gs->peer_pos[peer] = pos, gs->shared[peer] = total;
  if (n_shared) *n_shared = total;
  return NOMPK_OK;
}

extern "C" int nompk_gs_finalize_setup(nompk_gs_t *gs, int rank, int world, size_t *xchg_bytes, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!gs || gs->finalized || world < 1 || world > 64 || rank < 0 || rank >= world) {
    set_error("nompk_gs_finalize_setup: bad arguments (rank %d of %d)", rank, world);
    return NOMPK_EINVAL;
  }
  if (gs->peer_pos.empty()) gs->peer_pos.assign(world, nullptr), gs->shared.assign(world, 0);
  gs->rank = rank, gs->world = world;
  const size_t U = gs->n_unique;
  unsigned **d_pos = nullptr;
  unsigned char *active = nullptr;
  unsigned *iota = nullptr, *rcount = nullptr, *sel = nullptr, *nsel = nullptr, *order = nullptr;
  unsigned long long *first = nullptr, *first_sorted = nullptr;
  unsigned *cnt = nullptr, *rcnt = nullptr, *rflag = nullptr, *rstart = nullptr, *rslot = nullptr;
  void *tmp = nullptr;
  auto body = [&]() -> int {
    if (int e = dev_alloc(&d_pos, (size_t)world)) return e;
    NOMPK_CUDA_TRY(cudaMemcpyAsync(d_pos, gs->peer_pos.data(), world * sizeof(unsigned *), cudaMemcpyHostToDevice, stream));
    if (int e = dev_alloc(&active, U)) return e;
    if (int e = dev_alloc(&rcount, U)) return e;
    if (int e = dev_alloc(&sel, U)) return e;
    if (int e = dev_alloc(&nsel, 1)) return e;
    size_t G = 0;
    if (U > 0) {
      classify_kernel<<<blocks_for(U), 256, 0, stream>>>(gs->unique_ids, gs->run_count, d_pos, world, U, active, rcount);
      NOMPK_LAUNCH_CHECK("classify_kernel");
      if (int e = dev_alloc(&iota, U)) return e;
      iota_kernel<<<blocks_for(U), 256, 0, stream>>>(iota, U);
      NOMPK_LAUNCH_CHECK("iota_kernel");
      size_t bytes = 0;
      NOMPK_CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, bytes, iota, active, sel, nsel, U, stream));
      NOMPK_CUDA_TRY(cudaMalloc(&tmp, bytes ? bytes : 1));
      NOMPK_CUDA_TRY(cub::DeviceSelect::Flagged(tmp, bytes, iota, active, sel, nsel, U, stream));
      unsigned h = 0;
      NOMPK_CUDA_TRY(cudaMemcpyAsync(&h, nsel, sizeof(h), cudaMemcpyDeviceToHost, stream));
      NOMPK_CUDA_TRY(cudaStreamSynchronize(stream));
      NOMPK_CUDA_TRY(cudaFree(tmp));
      tmp = nullptr;
      G = h;
    }
    gs->G = G;
    // order the groups: shared with a peer first, then by their first local copy (first_index_kernel)
    if (int e = dev_alloc(&first, G)) return e;
    if (int e = dev_alloc(&first_sorted, G)) return e;
    if (int e = dev_alloc(&order, G)) return e;
    if (int e = dev_alloc(&cnt, G)) return e;
    if (int e = dev_alloc(&rcnt, G)) return e;
    if (int e = dev_alloc(&rflag, G)) return e;
    if (int e = dev_alloc(&rstart, G + 1)) return e;
    if (int e = dev_alloc(&rslot, G + 1)) return e;
    if (int e = dev_alloc(&gs->offsets, G + 1)) return e;
    if (int e = dev_alloc(&gs->remote_slot, G)) return e;
    size_t nnz = 0, R = 0, Q = 0;
    if (G > 0) {
      first_index_kernel<<<blocks_for(G), 256, 0, stream>>>(sel, gs->run_start, gs->sorted_idx, rcount, first, G);
      NOMPK_LAUNCH_CHECK("first_index_kernel");
      size_t bytes = 0;
      NOMPK_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, first, first_sorted, sel, order, G, 0, 33, stream));
      NOMPK_CUDA_TRY(cudaMalloc(&tmp, bytes ? bytes : 1));
      NOMPK_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, bytes, first, first_sorted, sel, order, G, 0, 33, stream));
      group_sizes_kernel<<<blocks_for(G), 256, 0, stream>>>(order, gs->run_count, rcount, cnt, rcnt, rflag, G);
      NOMPK_LAUNCH_CHECK("group_sizes_kernel");
      if (int e = exclusive_sum(cnt, gs->offsets, G, &nnz, stream)) return e;
      if (int e = exclusive_sum(rcnt, rstart, G, &R, stream)) return e;
      if (int e = exclusive_sum(rflag, rslot, G, &Q, stream)) return e;
    } else {
      NOMPK_CUDA_TRY(cudaMemsetAsync(gs->offsets, 0, sizeof(unsigned), stream));
    }
    if (nnz >= 0xffffffffull) {
      set_error("nompk_gs_finalize_setup: more than 2^32 shared copies");
      return NOMPK_EUNSUPPORTED;
    }
    gs->nnz = nnz, gs->R = R, gs->Q = Q;
    if (int e = dev_alloc(&gs->indices, nnz)) return e;
    if (int e = dev_alloc(&gs->rgroup, Q)) return e;
    if (int e = dev_alloc(&gs->roffsets, Q + 1)) return e;
    if (int e = dev_alloc(&gs->rpeer, R)) return e;
    if (int e = dev_alloc(&gs->rpos, R)) return e;
    if (int e = dev_alloc(&gs->partial, Q)) return e;
    if (int e = dev_alloc(&gs->ticket, 1)) return e;
    NOMPK_CUDA_TRY(cudaMemsetAsync(gs->ticket, 0, sizeof(unsigned), stream));
    if (G > 0) {
      fill_kernel<<<blocks_for(G), 256, 0, stream>>>(order, gs->run_start, gs->run_count, gs->sorted_idx, gs->offsets,
                                                      rcnt, rstart, rslot, d_pos, world, gs->indices, gs->remote_slot,
                                                      gs->rgroup, gs->roffsets, gs->rpeer, gs->rpos, G);
      NOMPK_LAUNCH_CHECK("fill_kernel");
    }
    if (Q > 0) {
      const size_t nblk = (G + kGsThreads - 1) / kGsThreads;
      unsigned *blk = nullptr, *blk_scan = nullptr;
      if (int e = dev_alloc(&blk, nblk)) return e;
      if (int e = dev_alloc(&blk_scan, nblk + 1)) {
        cudaFree(blk);
        return e;
      }
      size_t count = 0;
      cudaMemsetAsync(blk, 0, nblk * sizeof(unsigned), stream);
      mark_remote_ctas_kernel<<<blocks_for(G), 256, 0, stream>>>(gs->remote_slot, G, kGsThreads, blk);
      const int e = exclusive_sum(blk, blk_scan, nblk, &count, stream);
      cudaFree(blk), cudaFree(blk_scan);
      if (e) return e;
      gs->remote_ctas = (unsigned)count;
    }
    // warp layout for gs_local_warp_kernel: only if every group fits into one warp (true for any mesh numbering: a point is
    // shared by at most 8 hexahedra; an artificial numbering with longer groups keeps the one-group-per-thread kernel)
    if (G > 0) {
      unsigned *maxcnt = nullptr, *tile_warps = nullptr, *tile_start = nullptr;
      void *tmp2 = nullptr;
      auto layout = [&]() -> int {
        if (int e = dev_alloc(&maxcnt, 1)) return e;
        size_t bytes = 0;
        NOMPK_CUDA_TRY(cub::DeviceReduce::Max(nullptr, bytes, cnt, maxcnt, (int)G, stream));
        NOMPK_CUDA_TRY(cudaMalloc(&tmp2, bytes ? bytes : 1));
        NOMPK_CUDA_TRY(cub::DeviceReduce::Max(tmp2, bytes, cnt, maxcnt, (int)G, stream));
        unsigned h = 0;
        NOMPK_CUDA_TRY(cudaMemcpyAsync(&h, maxcnt, sizeof(h), cudaMemcpyDeviceToHost, stream));
        NOMPK_CUDA_TRY(cudaStreamSynchronize(stream));
        if (h > 32 || G > 0x7fffffffull) return NOMPK_OK;   // no layout: gs->nwarps stays 0
        const char *rows_env = getenv("NOMPK_GS_ROWS");
        gs->rows = rows_env && atoi(rows_env) == 8 ? 8 : kGsRowsDefault;
        const size_t ntiles = (G + kLayoutTile - 1) / kLayoutTile;
        if (int e = dev_alloc(&tile_warps, ntiles)) return e;
        if (int e = dev_alloc(&tile_start, ntiles + 1)) return e;
        layout_size_kernel<<<blocks_for(ntiles), 256, 0, stream>>>(gs->offsets, G, tile_warps);
        NOMPK_LAUNCH_CHECK("layout_size_kernel");
        size_t nwarps = 0;
        if (int e = exclusive_sum(tile_warps, tile_start, ntiles, &nwarps, stream)) return e;
        if (nwarps * 32 >= 0xffffffffull) return NOMPK_OK;
        if (int e = dev_alloc(&gs->pidx, nwarps * 32)) return e;
        if (int e = dev_alloc(&gs->heads, nwarps)) return e;
        if (int e = dev_alloc(&gs->wgroup, nwarps)) return e;
        if (int e = dev_alloc(&gs->rowinfo, nwarps)) return e;
        NOMPK_CUDA_TRY(cudaMemsetAsync(gs->rowinfo, 0, nwarps * sizeof(unsigned short), stream));
        NOMPK_CUDA_TRY(cudaMemsetAsync(gs->pidx, 0xff, nwarps * 32 * sizeof(unsigned), stream));
        NOMPK_CUDA_TRY(cudaMemsetAsync(gs->heads, 0, nwarps * sizeof(unsigned), stream));
        NOMPK_CUDA_TRY(cudaMemsetAsync(gs->wgroup, 0xff, nwarps * sizeof(unsigned), stream));
        layout_fill_kernel<<<blocks_for(ntiles), 256, 0, stream>>>(gs->offsets, gs->indices, G, tile_start, gs->pidx, gs->heads, gs->wgroup, gs->rowinfo);
        NOMPK_LAUNCH_CHECK("layout_fill_kernel");
        if (Q > 0) {
          const size_t nblk = (nwarps + kGsWarps * gs->rows - 1) / (kGsWarps * gs->rows);
          unsigned *blk = nullptr, *blk_scan = nullptr;
          if (int e = dev_alloc(&blk, nblk)) return e;
          if (int e = dev_alloc(&blk_scan, nblk + 1)) {
            cudaFree(blk);
            return e;
          }
          size_t count = 0;
          cudaMemsetAsync(blk, 0, nblk * sizeof(unsigned), stream);
          mark_remote_warp_ctas_kernel<<<blocks_for(nwarps), 256, 0, stream>>>(gs->wgroup, nwarps, Q, kGsWarps * gs->rows, blk);
          const int e = exclusive_sum(blk, blk_scan, nblk, &count, stream);
          cudaFree(blk), cudaFree(blk_scan);
          if (e) return e;
          gs->remote_ctas_warp = (unsigned)count;
        }
        NOMPK_CUDA_TRY(cudaStreamSynchronize(stream));
        gs->nwarps = nwarps;
        return NOMPK_OK;
      };
      const int e = layout();
      cudaStreamSynchronize(stream);
      cudaFree(maxcnt), cudaFree(tile_warps), cudaFree(tile_start), cudaFree(tmp2);
      if (e) return e;
      const char *force = getenv("NOMPK_GS_KERNEL");
      gs->force_group_kernel = force && !strcmp(force, "group");
    }
    const unsigned r32 = (unsigned)R;
    NOMPK_CUDA_TRY(cudaMemcpyAsync(gs->roffsets + Q, &r32, sizeof(unsigned), cudaMemcpyHostToDevice, stream));
    // exchange-buffer layout: flags, then two slots with one segment per peer (in rank order)
    gs->recv_off.assign(world, 0), gs->send_off.assign(world, 0);
    std::vector<int> neighbours;
    size_t off = 0;
    for (int r = 0; r < world; r++) {
      gs->recv_off[r] = off, off += gs->shared[r];
      if (gs->shared[r] > 0) neighbours.push_back(r);
    }
    gs->total_shared = off, gs->n_neighbours = (int)neighbours.size();
    if (int e = dev_alloc(&gs->d_recv_off, (size_t)world)) return e;
    if (int e = dev_alloc(&gs->d_send_off, (size_t)2 * world)) return e;
    if (int e = dev_alloc(&gs->d_peer_xchg, (size_t)world)) return e;
    if (int e = dev_alloc(&gs->d_neighbours, (size_t)world)) return e;
    NOMPK_CUDA_TRY(cudaMemcpyAsync(gs->d_recv_off, gs->recv_off.data(), world * sizeof(size_t), cudaMemcpyHostToDevice, stream));
    NOMPK_CUDA_TRY(cudaMemsetAsync(gs->d_peer_xchg, 0, world * sizeof(void *), stream));
    if (!neighbours.empty())
      NOMPK_CUDA_TRY(cudaMemcpyAsync(gs->d_neighbours, neighbours.data(), neighbours.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    NOMPK_CUDA_TRY(cudaStreamSynchronize(stream));
    return NOMPK_OK;
  };
  const int err = body();
  cudaStreamSynchronize(stream);
  void *scratch[] = {d_pos, iota, active, rcount, sel, nsel, first, first_sorted, order, cnt, rcnt, rflag, rstart, rslot, tmp};
  for (void *p : scratch) cudaFree(p);
  if (err) return err;
  // setup-only arrays are no longer needed
  for (unsigned *&p : gs->peer_pos) cudaFree(p), p = nullptr;
  cudaFree(gs->sorted_idx), gs->sorted_idx = nullptr;
  cudaFree(gs->run_start), gs->run_start = nullptr;
  cudaFree(gs->run_count), gs->run_count = nullptr;
  gs->finalized = true;
  gs->connected = gs->Q == 0;
  if (xchg_bytes) *xchg_bytes = gs->Q ? flags_bytes(world) + 2 * gs->total_shared * 8 : 0;
  return NOMPK_OK;
}

extern "C" int nompk_gs_recv_offsets(const nompk_gs_t *gs, size_t *offsets, size_t *counts) {
  if (!gs || !gs->finalized || !offsets) return NOMPK_EINVAL;
  for (int r = 0; r < gs->world; r++) {
    offsets[r] = gs->recv_off[r];
    if (counts) counts[r] = gs->shared[r];
  }
  return NOMPK_OK;
}

extern "C" int nompk_gs_connect(nompk_gs_t *gs, void *const *peer_xchg, const size_t *send_offsets,
                                const size_t *peer_totals, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!gs || !gs->finalized || !peer_xchg || !send_offsets || !peer_totals) {
    set_error("nompk_gs_connect: bad arguments");
    return NOMPK_EINVAL;
  }
  // where this rank's segment starts in rank r's buffer, per slot: slot * (ids rank r shares in total) + offset
  std::vector<size_t> table((size_t)2 * gs->world, 0);
  for (int r = 0; r < gs->world; r++) {
    if ((gs->shared[r] > 0 || r == gs->rank) && gs->Q > 0 && peer_xchg[r] == nullptr) {
      set_error("nompk_gs_connect: no exchange buffer for rank %d", r);
      return NOMPK_EINVAL;
    }
    if (gs->shared[r] > 0 && send_offsets[r] + gs->shared[r] > peer_totals[r]) {
      set_error("nompk_gs_connect: segment [%zu, %zu) does not fit rank %d's %zu shared ids", send_offsets[r],
                send_offsets[r] + gs->shared[r], r, peer_totals[r]);
      return NOMPK_EINVAL;
    }
    gs->send_off[r] = send_offsets[r];
    table[r] = send_offsets[r];
    table[(size_t)gs->world + r] = peer_totals[r] + send_offsets[r];
  }
  NOMPK_CUDA_TRY(cudaMemcpyAsync(gs->d_peer_xchg, peer_xchg, gs->world * sizeof(void *), cudaMemcpyHostToDevice, stream));
  NOMPK_CUDA_TRY(cudaMemcpyAsync(gs->d_send_off, table.data(), table.size() * sizeof(size_t), cudaMemcpyHostToDevice, stream));
  NOMPK_CUDA_TRY(cudaStreamSynchronize(stream));
  gs->connected = true;
  return NOMPK_OK;
}

extern "C" int nompk_gs_stats(const nompk_gs_t *gs, size_t out[8]) {
  if (!gs || !out) return NOMPK_EINVAL;
  out[0] = gs->n, out[1] = gs->n_unique, out[2] = gs->G, out[3] = gs->nnz, out[4] = gs->Q, out[5] = gs->R;
  out[6] = (size_t)gs->n_neighbours, out[7] = gs->total_shared;
  return NOMPK_OK;
}

extern "C" int nompk_gs_apply(nompk_gs_t *gs, nompk_red_op_t op, nompk_dtype_t dt, void *v,
                              unsigned long long *error_host_mapped, void *stream) {
  if (!gs || !gs->finalized || !gs->connected || (!v && gs->n)) {
    set_error("nompk_gs_apply: handle is not set up (create -> match_peer* -> finalize_setup -> connect)");
    return NOMPK_EINVAL;
  }
  gs_fn fn = nullptr;
  switch (op) {
  case NOMPK_RED_SUM: fn = pick_gs<NOMPK_RED_SUM>(dt); break;
  case NOMPK_RED_PROD: fn = pick_gs<NOMPK_RED_PROD>(dt); break;
  case NOMPK_RED_MIN: fn = pick_gs<NOMPK_RED_MIN>(dt); break;
  case NOMPK_RED_MAX: fn = pick_gs<NOMPK_RED_MAX>(dt); break;
  default: break;
  }
  if (!fn) {
    set_error("nompk_gs_apply: unsupported op %d / dtype %d", (int)op, (int)dt);
    return NOMPK_EINVAL;
  }
  return fn(gs, v, error_host_mapped, static_cast<cudaStream_t>(stream));
}
