// Reduction family: result <- reduce_i f(x[i], y[i]) for sum / prod / min / max, f = x or x*y (dot).
//
// What it replaces: the reference's reduce clause turns `var[0] += expr` into one partial per 512-thread work
// group (reference python/reduction.py:30-134), then synchronises, copies ALL partials to the host and folds
// them in a serial loop (reference src/reduction.c:33-88; one partial per 512 elements, so n = 2^28 would need
// 4 MiB of a 256 KiB scratch buffer).  Here the whole reduction is one launch:
//   1. one CTA per tile of 256 threads x kUnroll 128-bit loads (up to 65536 CTAs, beyond that the CTAs stride): all
//      loads of a thread are issued before the first use, one accumulator per vector slot (independent chains);
//   2. warp-shuffle tree, then one shared-memory slot per warp, then one partial per CTA to the workspace;
//   3. two levels of atomic tickets (nompk_gridreduce.cuh) fold the partials with a fixed association -> the result
//      does not depend on CTA scheduling (deterministic);
//   4. the result is stored to device memory and, optionally, straight into mapped pinned host memory.
// Integer sums/products wrap mod 2^32 / 2^64 exactly like the reference's C loop, in any order (bit-exact).
// fp64 sums differ from the serial loop only by association; the error bound is ~log2(n) ulp per partial tree
// against ~n ulp for the serial loop (tests compare with a compensated oracle at 1e-12 relative).
// Algorithmic bytes per element: 8 (sum of 8-byte type), 16 (dot).
#include <cfloat>
#include <climits>
#include <type_traits>

#include "nompk_common.cuh"
#include "nompk_gridreduce.cuh"

namespace nompk {
namespace {

constexpr int kBlock = 256;
constexpr int kUnroll = 4;

template <typename T> struct Limits;
template <> struct Limits<int> { static __device__ int lo() { return INT_MIN; } static __device__ int hi() { return INT_MAX; } };
template <> struct Limits<unsigned int> { static __device__ unsigned lo() { return 0u; } static __device__ unsigned hi() { return UINT_MAX; } };
template <> struct Limits<long long> { static __device__ long long lo() { return LLONG_MIN; } static __device__ long long hi() { return LLONG_MAX; } };
template <> struct Limits<unsigned long long> { static __device__ unsigned long long lo() { return 0ull; } static __device__ unsigned long long hi() { return ULLONG_MAX; } };
template <> struct Limits<float> { static __device__ float lo() { return -INFINITY; } static __device__ float hi() { return INFINITY; } };
template <> struct Limits<double> { static __device__ double lo() { return -(double)INFINITY; } static __device__ double hi() { return (double)INFINITY; } };

template <int OP, typename T> __device__ __forceinline__ T red_identity() {
  if constexpr (OP == NOMPK_RED_SUM) return T(0);
  if constexpr (OP == NOMPK_RED_PROD) return T(1);
  if constexpr (OP == NOMPK_RED_MIN) return Limits<T>::hi();
  if constexpr (OP == NOMPK_RED_MAX) return Limits<T>::lo();
  return T(0);
}

template <int OP, typename T> __device__ __forceinline__ T red_combine(T a, T b) {
  if constexpr (OP == NOMPK_RED_SUM) return a + b;
  if constexpr (OP == NOMPK_RED_PROD) return a * b;
  if constexpr (OP == NOMPK_RED_MIN) return b < a ? b : a;
  if constexpr (OP == NOMPK_RED_MAX) return b > a ? b : a;
  return a;
}

// acc <- acc (op) f(x, y)
template <int OP, bool DOT, typename T> __device__ __forceinline__ T red_accumulate(T acc, T x, T y) {
  if constexpr (DOT) {
    // fp dot products accumulate with one rounding per term (FMA); integers multiply-add mod 2^k.
    if constexpr (OP == NOMPK_RED_SUM && std::is_same<T, double>::value) return fma(x, y, acc);
    else if constexpr (OP == NOMPK_RED_SUM && std::is_same<T, float>::value) return fmaf(x, y, acc);
    else return red_combine<OP, T>(acc, x * y);
  } else {
    return red_combine<OP, T>(acc, x);
  }
}

template <int OP, typename T> __device__ __forceinline__ T warp_reduce(T v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = red_combine<OP, T>(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// Block-wide reduction; the result is valid in thread 0.
template <int OP, typename T> __device__ __forceinline__ T block_reduce(T v) {
  __shared__ T warp_part[kBlock / 32];
  v = warp_reduce<OP, T>(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // protects warp_part across the two uses in one kernel
  if (lane == 0) warp_part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < kBlock / 32 ? warp_part[lane] : red_identity<OP, T>();
    v = warp_reduce<OP, T>(v);
  }
  return v;
}

template <int OP, typename T> struct RedOp {
  static __device__ __forceinline__ T identity() { return red_identity<OP, T>(); }
  static __device__ __forceinline__ T combine(T a, T b) { return red_combine<OP, T>(a, b); }
};

template <int OP, bool DOT, bool VEC, typename T>
__global__ void __launch_bounds__(kBlock)
reduce_kernel(const T *__restrict__ x, const T *__restrict__ y, size_t n, void *__restrict__ workspace,
              T *__restrict__ result, T *__restrict__ result_host, unsigned long long host_seq, PeerExchange px) {
  T acc[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; u++) acc[u] = red_identity<OP, T>();

  if constexpr (VEC) {
    constexpr int L = Vec16<T>::kLanes;
    const size_t nvec = n / L;
    constexpr size_t kTile = (size_t)kBlock * kUnroll;
    const size_t stride = (size_t)gridDim.x * kTile;
    for (size_t base = (size_t)blockIdx.x * kTile; base < nvec; base += stride) {
      Vec16<T> vx[kUnroll] = {}, vy[kUnroll] = {};
      if (base + kTile <= nvec) {
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
          const size_t e = (base + (size_t)u * kBlock + threadIdx.x) * L;
          vx[u] = ld_vec_ro(x + e);
          if constexpr (DOT) vy[u] = ld_vec_ro(y + e);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
#pragma unroll
          for (int l = 0; l < L; l++) acc[u] = red_accumulate<OP, DOT, T>(acc[u], vx[u].v[l], vy[u].v[l]);
      } else {
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
          const size_t iv = base + (size_t)u * kBlock + threadIdx.x;
          if (iv < nvec) {
            const size_t e = iv * L;
            vx[u] = ld_vec_ro(x + e);
            if constexpr (DOT) vy[u] = ld_vec_ro(y + e);
#pragma unroll
            for (int l = 0; l < L; l++) acc[u] = red_accumulate<OP, DOT, T>(acc[u], vx[u].v[l], vy[u].v[l]);
          }
        }
      }
    }
    if (blockIdx.x == 0) {  // n % L scalar tail
      const size_t e = nvec * L + threadIdx.x;
      if (e < n) acc[0] = red_accumulate<OP, DOT, T>(acc[0], x[e], DOT ? y[e] : T(0));
    }
  } else {
    const size_t stride = (size_t)gridDim.x * kBlock;
    for (size_t e = (size_t)blockIdx.x * kBlock + threadIdx.x; e < n; e += stride)
      acc[0] = red_accumulate<OP, DOT, T>(acc[0], x[e], DOT ? y[e] : T(0));
  }

  T v = acc[0];
#pragma unroll
  for (int u = 1; u < kUnroll; u++) v = red_combine<OP, T>(v, acc[u]);
  v = block_reduce<OP, T>(v);

  grid_finish<RedOp<OP, T>, T, kBlock>(v, workspace, result, result_host, host_seq, px);
}

template <int OP, typename T>
int launch_reduce(size_t n, const void *x_, const void *y_, void *result, void *result_host,
                  unsigned long long host_seq, void *workspace, const PeerExchange &px, cudaStream_t stream) {
  const T *x = static_cast<const T *>(x_);
  const T *y = static_cast<const T *>(y_);
  if (!result || !workspace || (n > 0 && !x)) {
    set_error("nompk_reduce: NULL result/workspace/operand");
    return NOMPK_EINVAL;
  }
  T *res = static_cast<T *>(result);
  T *res_h = static_cast<T *>(result_host);

  constexpr int L = Vec16<T>::kLanes;
  const bool vec = is_aligned16(x) && (!y || is_aligned16(y));
  const size_t per_block = vec ? (size_t)kBlock * kUnroll * L : (size_t)kBlock;
  // one tile per CTA (up to 65536 CTAs, then the CTAs stride): many small CTAs stream faster than a persistent grid
  size_t blocks = (n + per_block - 1) / per_block;
  // up to 2^26 elements a grid of <= 2048 striding CTAs with the one-level finish is faster end to end (the second
  // ticket level costs ~2 us); beyond that one tile per CTA wins (+3 % bandwidth at n = 2^28)
  if (n < ((size_t)1 << 26) && blocks > (size_t)kRedSingleLevel) blocks = kRedSingleLevel;
  if (blocks > (size_t)kRedMaxCtas) blocks = kRedMaxCtas;
  if (blocks == 0) blocks = 1;
  const unsigned g = (unsigned)blocks;

  if (y) {
    if (vec) reduce_kernel<OP, true, true, T><<<g, kBlock, 0, stream>>>(x, y, n, workspace, res, res_h, host_seq, px);
    else reduce_kernel<OP, true, false, T><<<g, kBlock, 0, stream>>>(x, y, n, workspace, res, res_h, host_seq, px);
  } else {
    if (vec) reduce_kernel<OP, false, true, T><<<g, kBlock, 0, stream>>>(x, y, n, workspace, res, res_h, host_seq, px);
    else reduce_kernel<OP, false, false, T><<<g, kBlock, 0, stream>>>(x, y, n, workspace, res, res_h, host_seq, px);
  }
  NOMPK_LAUNCH_CHECK("reduce_kernel");
  return NOMPK_OK;
}

typedef int (*reduce_fn)(size_t, const void *, const void *, void *, void *, unsigned long long, void *,
                         const PeerExchange &, cudaStream_t);

// SUM/PROD of integers wrap, so signed == unsigned bitwise; MIN/MAX need the real type.
template <int OP> reduce_fn pick_dtype(nompk_dtype_t dt) {
  constexpr bool ring = (OP == NOMPK_RED_SUM || OP == NOMPK_RED_PROD);
  switch (dt) {
  case NOMPK_I32:
    if constexpr (ring) return launch_reduce<OP, unsigned int>;
    else return launch_reduce<OP, int>;
  case NOMPK_U32: return launch_reduce<OP, unsigned int>;
  case NOMPK_I64:
    if constexpr (ring) return launch_reduce<OP, unsigned long long>;
    else return launch_reduce<OP, long long>;
  case NOMPK_U64: return launch_reduce<OP, unsigned long long>;
  case NOMPK_F32: return launch_reduce<OP, float>;
  case NOMPK_F64: return launch_reduce<OP, double>;
  }
  return nullptr;
}

// ---------------------------------------------------------------------------------------------------
// One-shot all-reduce of ONE scalar over NVLink peer memory (one process per GPU, buffers opened with CUDA IPC).
//
// Every rank owns an exchange buffer  xchg[2 slots][world][16 B] = {value, sequence}.  For reduction number `seq`
// thread r of the single block stores this rank's value and then `seq` into slot (seq & 1), entry `rank`, of peer
// r's buffer (P2P stores, system-scope fence between value and flag), then spins until entry r of its OWN buffer
// carries `seq`.  Thread 0 folds the `world` values in rank order -- the same order on every rank, so all ranks
// obtain bit-identical results -- and publishes the result to device memory and to the host.
// Two slots suffice: a peer can only write sequence s+2 into a slot after this rank contributed to s+1, which it
// does only after it finished reading s (stream order).
// This replaces an ncclAllReduce of 8 bytes (launch + protocol latency of tens of microseconds) by one tiny kernel
// whose cost is one NVLink round trip.
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 64;
constexpr unsigned long long kAllreduceTimeoutNs = 20ull * 1000 * 1000 * 1000;  // 20 s


template <int OP, typename T>
__global__ void __launch_bounds__(kMaxRanks)
allreduce_scalar_kernel(T *__restrict__ value, T *__restrict__ result_host, unsigned long long host_seq,
                        void *const *__restrict__ peer_xchg, int rank, int world, unsigned long long seq,
                        unsigned long long *seq_dev, unsigned long long *error_host) {
  __shared__ T vals[kMaxRanks];
  __shared__ int timed_out;
  __shared__ unsigned long long seq_shared;
  const int r = threadIdx.x;
  if (r == 0) {
    timed_out = 0;
    if (seq_dev) {  // call number in device memory (nompk_peers_t): take the next one
      seq = *reinterpret_cast<volatile unsigned long long *>(seq_dev) + 1;
      *reinterpret_cast<volatile unsigned long long *>(seq_dev) = seq;
    }
    seq_shared = seq;
  }
  __syncthreads();
  seq = seq_shared;
  const size_t slot = (size_t)(seq & 1ull) * (size_t)world;
  if (r < world) {
    const T mine = *value;
    char *dst = static_cast<char *>(peer_xchg[r]) + (slot + (size_t)rank) * 16;
    *reinterpret_cast<volatile T *>(dst) = mine;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(dst + 8) = seq;
    char *src = static_cast<char *>(peer_xchg[rank]) + (slot + (size_t)r) * 16;
    // A peer that never arrives (crashed rank, mismatched call order) must not hang the GPU: give up after
    // kAllreduceTimeoutNs and report it through the error word of the host block.
    const unsigned long long t0 = global_timer_ns();
    bool ok = true;
    while (*reinterpret_cast<volatile unsigned long long *>(src + 8) != seq) {
      if (global_timer_ns() - t0 > kAllreduceTimeoutNs) {
        ok = false;
        break;
      }
    }
    __threadfence_system();
    vals[r] = *reinterpret_cast<volatile T *>(src);
    if (!ok) atomicExch(&timed_out, 1);
  }
  __syncthreads();
  if (r == 0) {
    T acc = vals[0];
    for (int i = 1; i < world; i++) acc = red_combine<OP, T>(acc, vals[i]);
    *value = acc;
    if (timed_out) {
      if (error_host) *reinterpret_cast<volatile unsigned long long *>(error_host) = seq;
      else if (result_host) *reinterpret_cast<volatile unsigned long long *>(reinterpret_cast<char *>(result_host) + 16) = seq;
    }
    if (result_host) publish_to_host(result_host, acc, host_seq);
  }
}

template <int OP, typename T>
int launch_allreduce(void *value, void *result_host, unsigned long long host_seq, void *const *peer_xchg, int rank,
                     int world, unsigned long long seq, unsigned long long *seq_dev, unsigned long long *error_host,
                     cudaStream_t stream) {
  allreduce_scalar_kernel<OP, T><<<1, kMaxRanks, 0, stream>>>(static_cast<T *>(value), static_cast<T *>(result_host),
                                                             host_seq, peer_xchg, rank, world, seq, seq_dev, error_host);
  NOMPK_LAUNCH_CHECK("allreduce_scalar_kernel");
  return NOMPK_OK;
}

typedef int (*allreduce_fn)(void *, void *, unsigned long long, void *const *, int, int, unsigned long long,
                            unsigned long long *, unsigned long long *, cudaStream_t);

template <int OP> allreduce_fn pick_allreduce(nompk_dtype_t dt) {
  constexpr bool ring = (OP == NOMPK_RED_SUM || OP == NOMPK_RED_PROD);
  switch (dt) {
  case NOMPK_I32:
    if constexpr (ring) return launch_allreduce<OP, unsigned int>;
    else return launch_allreduce<OP, int>;
  case NOMPK_U32: return launch_allreduce<OP, unsigned int>;
  case NOMPK_I64:
    if constexpr (ring) return launch_allreduce<OP, unsigned long long>;
    else return launch_allreduce<OP, long long>;
  case NOMPK_U64: return launch_allreduce<OP, unsigned long long>;
  case NOMPK_F32: return launch_allreduce<OP, float>;
  case NOMPK_F64: return launch_allreduce<OP, double>;
  }
  return nullptr;
}

}  // namespace
}  // namespace nompk

extern "C" size_t nompk_allreduce_xchg_bytes(int world) { return (size_t)2 * (size_t)world * 16; }

extern "C" int nompk_allreduce_scalar(nompk_red_op_t op, nompk_dtype_t dt, void *value, void *result_host_mapped,
                                      unsigned long long host_seq, void *const *peer_xchg, int rank, int world,
                                      unsigned long long seq, void *stream) {
  const nompk_peers_t peers = {peer_xchg, rank, world, seq, nullptr, nullptr};
  return nompk_allreduce_scalar_peers(op, dt, value, result_host_mapped, host_seq, &peers, stream);
}

extern "C" int nompk_allreduce_scalar_peers(nompk_red_op_t op, nompk_dtype_t dt, void *value, void *result_host_mapped,
                                            unsigned long long host_seq, const nompk_peers_t *peers, void *stream) {
  using namespace nompk;
  PeerExchange px;
  if (!value || !peers || peers->world < 1 || !px.set(peers, kMaxRanks)) {
    set_error("nompk_allreduce_scalar: bad arguments (world %d, rank %d)", peers ? peers->world : 0, peers ? peers->rank : 0);
    return NOMPK_EINVAL;
  }
  allreduce_fn fn = nullptr;
  switch (op) {
  case NOMPK_RED_SUM: fn = pick_allreduce<NOMPK_RED_SUM>(dt); break;
  case NOMPK_RED_PROD: fn = pick_allreduce<NOMPK_RED_PROD>(dt); break;
  case NOMPK_RED_MIN: fn = pick_allreduce<NOMPK_RED_MIN>(dt); break;
  case NOMPK_RED_MAX: fn = pick_allreduce<NOMPK_RED_MAX>(dt); break;
  default: break;
  }
  if (!fn) {
    set_error("nompk_allreduce_scalar: unsupported op %d / dtype %d", (int)op, (int)dt);
    return NOMPK_EINVAL;
  }
  return fn(value, result_host_mapped, host_seq, px.peer_xchg, px.rank, px.world, px.seq, px.seq_dev, px.error_host,
            static_cast<cudaStream_t>(stream));
}

extern "C" size_t nompk_reduce_workspace_bytes(void) { return nompk::kWsBytes; }

extern "C" void nompk_reduce_workspace_layout(size_t offsets[4]) {
  offsets[0] = nompk::kWsTicket, offsets[1] = nompk::kWsGroupTicket, offsets[2] = nompk::kWsL2, offsets[3] = nompk::kWsL1;
}

extern "C" int nompk_reduce(nompk_red_op_t op, nompk_dtype_t dt, size_t n, const void *x, const void *y,
                            void *result, void *result_host_mapped, unsigned long long host_seq, void *workspace,
                            void *stream) {
  return nompk_reduce_peers(op, dt, n, x, y, result, result_host_mapped, host_seq, workspace, nullptr, stream);
}

extern "C" int nompk_reduce_peers(nompk_red_op_t op, nompk_dtype_t dt, size_t n, const void *x, const void *y,
                                  void *result, void *result_host_mapped, unsigned long long host_seq, void *workspace,
                                  const nompk_peers_t *peers, void *stream) {
  using namespace nompk;
  PeerExchange px;
  if (peers && peers->world > 1) {
    if (!px.set(peers, kMaxFusedRanks)) {
      set_error("nompk_reduce_peers: bad peer description (rank %d of %d; at most %d ranks)", peers->rank, peers->world,
                kMaxFusedRanks);
      return NOMPK_EINVAL;
    }
  }
  reduce_fn fn = nullptr;
  switch (op) {
  case NOMPK_RED_SUM: fn = pick_dtype<NOMPK_RED_SUM>(dt); break;
  case NOMPK_RED_PROD: fn = pick_dtype<NOMPK_RED_PROD>(dt); break;
  case NOMPK_RED_MIN: fn = pick_dtype<NOMPK_RED_MIN>(dt); break;
  case NOMPK_RED_MAX: fn = pick_dtype<NOMPK_RED_MAX>(dt); break;
  default: break;
  }
  if (!fn) {
    set_error("nompk_reduce: unsupported op %d / dtype %d", (int)op, (int)dt);
    return NOMPK_EINVAL;
  }
  return fn(n, x, y, result, result_host_mapped, host_seq, workspace, px, static_cast<cudaStream_t>(stream));
}
