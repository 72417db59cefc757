// Deterministic single-pass grid-wide finish of a reduction, shared by reduce.cu and the fused Ax + dot kernel (and
// mirrored in text by the generated reduce skeleton, python/nomp_bridge/reduction.py).
//
// The streaming front ends run fastest with MANY small CTAs (one tile each; the hardware CTA scheduler beats a
// persistent grid, tools/exp/exp_map.cu), so a launch can have up to kMaxCtas = 65536 partials -- too many for one
// CTA to fold at the end.  Two levels of atomic tickets keep the fold short and its association fixed:
//   level 1: CTA b stores its partial, then takes ticket[b / 32]; the CTA that takes the last ticket of a group of
//            32 folds the group's partials with one warp (lane i <- partial i, shuffle tree) into a level-2 partial;
//   level 2: that CTA then takes the global ticket; the last one folds the <= 2048 level-2 partials (256 threads, 8
//            each, block tree) and publishes the result.
// Grids of at most 2048 CTAs skip level 1 (the last CTA folds all partials); a single CTA publishes directly.
// Every fold has a fixed shape, so the result does not depend on the order in which CTAs finish.  All tickets are
// reset by the CTA that consumes them: the workspace is ready for the next launch on the same stream.
#pragma once

#include "nompk_common.cuh"

namespace nompk {

constexpr int kRedGroup = 32;
constexpr int kRedMaxGroups = 2048;
constexpr int kRedMaxCtas = kRedGroup * kRedMaxGroups;              // 65536
constexpr int kRedSingleLevel = 2048;                               // grids up to this size use one ticket level
constexpr size_t kWsTicket = 0;                                     // unsigned int
constexpr size_t kWsGroupTicket = 64;                               // unsigned int[2048]
constexpr size_t kWsL2 = kWsGroupTicket + 4 * kRedMaxGroups;        // 8-byte slots[2048]
constexpr size_t kWsL1 = kWsL2 + 8 * kRedMaxGroups;                 // 8-byte slots[65536]
constexpr size_t kWsBytes = kWsL1 + 8 * (size_t)kRedMaxCtas;        // 548928

// Host-visible result: mapped pinned memory, [0,8) the value, [8,16) a sequence number written AFTER the value
// (system-scope fence in between).  The host spins on the sequence number instead of synchronising the stream.
template <typename T> __device__ __forceinline__ void publish_to_host(T *result_host, T value, unsigned long long seq) {
  *reinterpret_cast<volatile T *>(result_host) = value;
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long *>(reinterpret_cast<char *>(result_host) + 8) = seq;
}

// All-reduce over ranks fused into the finish.  Every rank owns an exchange buffer xchg[2 slots][world][16 B] =
// {value, sequence} mapped by its peers (CUDA IPC over NVLink; reduce.cu describes the protocol and why two slots
// suffice).  world <= 1 means "no exchange".
struct PeerExchange {
  void *const *peer_xchg = nullptr;
  int rank = 0, world = 1;
  unsigned long long seq = 0;
  unsigned long long *seq_dev = nullptr;     // call number in device memory (then `seq` is unused): nompk.h, nompk_peers_t
  unsigned long long *error_host = nullptr;  // mapped host word that receives the call number on a timeout

  // host side: fill from the C-ABI description; false if it is malformed
  template <typename P> bool set(const P *p, int max_ranks) {
    if (p->world > max_ranks || p->rank < 0 || p->rank >= p->world || !p->peer_xchg || (p->seq == 0 && !p->seq_dev)) return false;
    peer_xchg = p->peer_xchg, rank = p->rank, world = p->world, seq = p->seq, seq_dev = p->seq_dev, error_host = p->error_host_mapped;
    return true;
  }
};

constexpr int kMaxFusedRanks = 32;  // one lane of the finishing warp per rank
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// The grid's result is in lane 0 of the calling warp (all 32 lanes call).  With peers: lane r stores it into rank r's
// buffer, waits for rank r's value in this rank's buffer, and the values are folded in rank order (the same on
// every rank: bit-identical results).  Then the value goes to device memory and, if asked, to the host block
// {value, sequence, error}.  The collective costs no launch of its own: the kernel that produced the value delivers
// it to the peers the moment its last CTA has it.
template <typename Op, typename T>
__device__ __forceinline__ void finish_result(T v, T *result, T *result_host, unsigned long long host_seq,
                                              const PeerExchange &px) {
  const int lane = threadIdx.x & 31;
  bool timed_out = false;
  unsigned long long cseq = px.seq;
  if (px.world > 1) {
    if (px.seq_dev) {  // the call number lives in device memory: this warp is its only user on the stream
      if (lane == 0) {
        cseq = *reinterpret_cast<volatile unsigned long long *>(px.seq_dev) + 1;
        *reinterpret_cast<volatile unsigned long long *>(px.seq_dev) = cseq;
      }
      cseq = __shfl_sync(0xffffffffu, cseq, 0);
    }
    v = __shfl_sync(0xffffffffu, v, 0);
    const size_t slot = (size_t)(cseq & 1ull) * (size_t)px.world;
    T got = Op::identity();
    bool ok = true;
    if (lane < px.world) {
      char *dst = static_cast<char *>(px.peer_xchg[lane]) + (slot + (size_t)px.rank) * 16;
      *reinterpret_cast<volatile T *>(dst) = v;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long *>(dst + 8) = cseq;
      const char *src = static_cast<const char *>(px.peer_xchg[px.rank]) + (slot + (size_t)lane) * 16;
      const unsigned long long t0 = global_timer_ns();
      while (*reinterpret_cast<const volatile unsigned long long *>(src + 8) != cseq) {
        if (global_timer_ns() - t0 > kPeerTimeoutNs) {  // a peer that never arrives must not hang the GPU
          ok = false;
          break;
        }
      }
      __threadfence_system();
      got = *reinterpret_cast<const volatile T *>(src);
    }
    T acc = __shfl_sync(0xffffffffu, got, 0);
    for (int r = 1; r < px.world; r++) acc = Op::combine(acc, __shfl_sync(0xffffffffu, got, r));
    timed_out = __any_sync(0xffffffffu, !ok);
    v = acc;
  }
  if (lane == 0) {
    *result = v;
    if (timed_out) {
      if (px.error_host) *reinterpret_cast<volatile unsigned long long *>(px.error_host) = cseq;
      else if (result_host) *reinterpret_cast<volatile unsigned long long *>(reinterpret_cast<char *>(result_host) + 16) = cseq;
    }
    if (result_host) publish_to_host(result_host, v, host_seq);
  }
}

// Combine functor interface: Op::identity(), Op::combine(a, b).
// `v` must hold the CTA's partial in thread 0.  Every thread of the CTA must call; kThreads >= 64, multiple of 32.
template <typename Op, typename T, int kThreads>
__device__ __forceinline__ void grid_finish(T v, void *ws_, T *result, T *result_host, unsigned long long host_seq,
                                            const PeerExchange &px = PeerExchange()) {
  char *ws = static_cast<char *>(ws_);
  unsigned int *ticket = reinterpret_cast<unsigned int *>(ws + kWsTicket);
  unsigned int *group_ticket = reinterpret_cast<unsigned int *>(ws + kWsGroupTicket);
  T *l2 = reinterpret_cast<T *>(ws + kWsL2);   // 8-byte slots: T occupies the first sizeof(T) bytes
  T *l1 = reinterpret_cast<T *>(ws + kWsL1);
  constexpr int kSlot = 8 / sizeof(T);          // stride between slots in units of T
  __shared__ int role;                          // 0: done, 1: last of its group, 2: last of all
  __shared__ T warp_part[kThreads / 32];

  const unsigned int b = blockIdx.x, nb = gridDim.x;
  if (nb == 1) {  // a single CTA: nothing to combine, no atomics
    if (threadIdx.x < 32) finish_result<Op, T>(v, result, result_host, host_seq, px);
    return;
  }
  if (nb <= kRedSingleLevel) {
    // Few CTAs: one ticket level is shorter (one atomic round trip less).  The last CTA folds the level-1 partials.
    if (threadIdx.x == 0) {
      l1[(size_t)b * kSlot] = v;
      __threadfence();
      role = (atomicAdd(ticket, 1u) == nb - 1) ? 2 : 0;
    }
    __syncthreads();
    if (role != 2) return;
    __threadfence();
    T w1 = Op::identity();
    for (unsigned int i = threadIdx.x; i < nb; i += kThreads) w1 = Op::combine(w1, __ldcg(l1 + (size_t)i * kSlot));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) w1 = Op::combine(w1, __shfl_xor_sync(0xffffffffu, w1, off));
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = w1;
    __syncthreads();
    if (threadIdx.x < 32) {
      w1 = threadIdx.x < kThreads / 32 ? warp_part[threadIdx.x] : Op::identity();
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) w1 = Op::combine(w1, __shfl_xor_sync(0xffffffffu, w1, off));
      if (threadIdx.x == 0) *ticket = 0u;
      finish_result<Op, T>(w1, result, result_host, host_seq, px);
    }
    return;
  }
  const unsigned int group = b / kRedGroup, ngroups = (nb + kRedGroup - 1) / kRedGroup;
  const unsigned int gsize = (group == ngroups - 1) ? nb - group * kRedGroup : kRedGroup;
  if (threadIdx.x == 0) {
    l1[(size_t)b * kSlot] = v;
    __threadfence();
    role = (atomicAdd(&group_ticket[group], 1u) == gsize - 1) ? 1 : 0;
  }
  __syncthreads();
  if (role == 0) return;

  if (threadIdx.x < 32) {
    __threadfence();
    T g = threadIdx.x < gsize ? __ldcg(l1 + ((size_t)group * kRedGroup + threadIdx.x) * kSlot) : Op::identity();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) g = Op::combine(g, __shfl_xor_sync(0xffffffffu, g, off));
    if (threadIdx.x == 0) {
      l2[(size_t)group * kSlot] = g;
      group_ticket[group] = 0u;
      __threadfence();
      role = (atomicAdd(ticket, 1u) == ngroups - 1) ? 2 : 0;
    }
  }
  __syncthreads();
  if (role != 2) return;

  __threadfence();
  T w = Op::identity();
  for (unsigned int i = threadIdx.x; i < ngroups; i += kThreads) w = Op::combine(w, __ldcg(l2 + (size_t)i * kSlot));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) w = Op::combine(w, __shfl_xor_sync(0xffffffffu, w, off));
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x < 32) {
    w = threadIdx.x < kThreads / 32 ? warp_part[threadIdx.x] : Op::identity();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) w = Op::combine(w, __shfl_xor_sync(0xffffffffu, w, off));
    if (threadIdx.x == 0) *ticket = 0u;
    finish_result<Op, T>(w, result, result_host, host_seq, px);
  }
}

}  // namespace nompk
