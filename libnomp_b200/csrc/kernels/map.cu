// Elementwise map family: y <- op(y, x, z, alpha, beta).
//
// What it replaces: the loopy-generated "one element per thread, scalar load, bounds predicate" kernel that
// the reference launches with grid = ceil(N/512), block = 512 (reference tests/nomp_api_100.py:11-20,
// backends/unified-cuda-hip-impl.h:143-158) for loops such as `a[i] += b[i]`
// (reference tests/nomp-api-200-impl.h:36-40, tests/nomp-api-600-impl.h:36-40).
//
// Design (HBM-bound, <= 0.25 flop/B): every operand moves as 128-bit LDG/STG, ONE vector per thread, and the grid
// has one CTA per tile (up to 2^31 CTAs): measured on B200 (tools/exp/exp_map.cu, a[i] += b[i], n = 2^28 fp64) the
// hardware CTA scheduler streams 7.07-7.17 TB/s this way against 6.28 TB/s for a persistent grid of 4 CTAs per SM
// striding over the array with 4 vectors per thread in flight, and 6.6 TB/s for cudaMemcpy D2D.  CTAs of 1024 threads
// for large arrays, 256 for small ones (so that every SM gets work).
// A scalar variant handles operands that are not 16-byte aligned and the n % lanes tail.
// Arithmetic keeps the C expression's roundings (no FMA contraction) so results are bit-identical to the
// serial loop.  Algorithmic bytes per element: 24 (ADD/SUB/MUL/AXPY/XPAY/AXPBY/ADD3), 16 (SCALE/COPY), 8 (FILL)
// for 8-byte types.
#include "nompk_common.cuh"

namespace nompk {
namespace {

constexpr int kBlock = 256;
constexpr int kBigBlock = 1024;

template <int OP> struct MapTraits;
#define NOMPK_MAP_TRAITS(OP, RY, UX, UZ)                                                           \
  template <> struct MapTraits<OP> {                                                               \
    static constexpr bool reads_y = RY, uses_x = UX, uses_z = UZ;                                  \
  };
NOMPK_MAP_TRAITS(NOMPK_MAP_ADD, true, true, false)
NOMPK_MAP_TRAITS(NOMPK_MAP_SUB, true, true, false)
NOMPK_MAP_TRAITS(NOMPK_MAP_MUL, true, true, false)
NOMPK_MAP_TRAITS(NOMPK_MAP_AXPY, true, true, false)
NOMPK_MAP_TRAITS(NOMPK_MAP_XPAY, true, true, false)
NOMPK_MAP_TRAITS(NOMPK_MAP_AXPBY, true, true, false)
NOMPK_MAP_TRAITS(NOMPK_MAP_SCALE, true, false, false)
NOMPK_MAP_TRAITS(NOMPK_MAP_COPY, false, true, false)
NOMPK_MAP_TRAITS(NOMPK_MAP_FILL, false, false, false)
NOMPK_MAP_TRAITS(NOMPK_MAP_ADD3, false, true, true)
#undef NOMPK_MAP_TRAITS

template <int OP, typename T>
__device__ __forceinline__ T map_apply(T y, T x, T z, T alpha, T beta) {
  if constexpr (OP == NOMPK_MAP_ADD) return op_add(y, x);
  if constexpr (OP == NOMPK_MAP_SUB) return op_sub(y, x);
  if constexpr (OP == NOMPK_MAP_MUL) return op_mul(y, x);
  if constexpr (OP == NOMPK_MAP_AXPY) return op_add(y, op_mul(alpha, x));
  if constexpr (OP == NOMPK_MAP_XPAY) return op_add(x, op_mul(alpha, y));
  if constexpr (OP == NOMPK_MAP_AXPBY) return op_add(op_mul(alpha, x), op_mul(beta, y));
  if constexpr (OP == NOMPK_MAP_SCALE) return op_mul(alpha, y);
  if constexpr (OP == NOMPK_MAP_COPY) return x;
  if constexpr (OP == NOMPK_MAP_FILL) return alpha;
  if constexpr (OP == NOMPK_MAP_ADD3) return op_add(x, z);
  return y;
}

// Vector kernel.  nvec = number of complete 16-byte vectors; elements [nvec*lanes, n) are the scalar tail,
// done by the first threads of block 0.
template <int OP, typename T, int kUnroll, int kThreads>
__global__ void __launch_bounds__(kThreads)
map_vec_kernel(T *__restrict__ y, const T *__restrict__ x, const T *__restrict__ z, T alpha, T beta,
               size_t nvec, size_t n) {
  using Tr = MapTraits<OP>;
  constexpr int L = Vec16<T>::kLanes;
  constexpr int kBlock = kThreads;  // shadows the namespace constant inside this kernel
  constexpr size_t kTile = (size_t)kBlock * kUnroll;
  const size_t stride = (size_t)gridDim.x * kTile;

  for (size_t base = (size_t)blockIdx.x * kTile; base < nvec; base += stride) {
    Vec16<T> vy[kUnroll] = {}, vx[kUnroll] = {}, vz[kUnroll] = {};
    if (base + kTile <= nvec) {
      // Full tile: issue every load of the tile before the first dependent instruction.
#pragma unroll
      for (int u = 0; u < kUnroll; u++) {
        const size_t e = (base + (size_t)u * kBlock + threadIdx.x) * L;
        if constexpr (Tr::uses_x) vx[u] = ld_vec_ro(x + e);
        if constexpr (Tr::uses_z) vz[u] = ld_vec_ro(z + e);
        if constexpr (Tr::reads_y) vy[u] = ld_vec(y + e);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; u++) {
        const size_t e = (base + (size_t)u * kBlock + threadIdx.x) * L;
        Vec16<T> r;
#pragma unroll
        for (int l = 0; l < L; l++) r.v[l] = map_apply<OP, T>(vy[u].v[l], vx[u].v[l], vz[u].v[l], alpha, beta);
        st_vec(y + e, r);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kUnroll; u++) {
        const size_t iv = base + (size_t)u * kBlock + threadIdx.x;
        if (iv < nvec) {
          const size_t e = iv * L;
          if constexpr (Tr::uses_x) vx[u] = ld_vec_ro(x + e);
          if constexpr (Tr::uses_z) vz[u] = ld_vec_ro(z + e);
          if constexpr (Tr::reads_y) vy[u] = ld_vec(y + e);
          Vec16<T> r;
#pragma unroll
          for (int l = 0; l < L; l++)
            r.v[l] = map_apply<OP, T>(vy[u].v[l], vx[u].v[l], vz[u].v[l], alpha, beta);
          st_vec(y + e, r);
        }
      }
    }
  }

  if (blockIdx.x == 0) {
    const size_t e = nvec * L + threadIdx.x;
    if (e < n) {
      T vy_ = Tr::reads_y ? y[e] : T(0), vx_ = Tr::uses_x ? x[e] : T(0), vz_ = Tr::uses_z ? z[e] : T(0);
      y[e] = map_apply<OP, T>(vy_, vx_, vz_, alpha, beta);
    }
  }
}

// Scalar kernel for misaligned operands (sub-range mappings whose byte offset is not a multiple of 16).
template <int OP, typename T>
__global__ void __launch_bounds__(kBlock)
map_scalar_kernel(T *__restrict__ y, const T *__restrict__ x, const T *__restrict__ z, T alpha, T beta,
                  size_t n) {
  using Tr = MapTraits<OP>;
  const size_t stride = (size_t)gridDim.x * kBlock;
  for (size_t e = (size_t)blockIdx.x * kBlock + threadIdx.x; e < n; e += stride) {
    T vy_ = Tr::reads_y ? y[e] : T(0), vx_ = Tr::uses_x ? x[e] : T(0), vz_ = Tr::uses_z ? z[e] : T(0);
    y[e] = map_apply<OP, T>(vy_, vx_, vz_, alpha, beta);
  }
}

template <int OP, typename T>
int launch_map(size_t n, void *y_, const void *x_, const void *z_, const void *alpha_, const void *beta_,
               cudaStream_t stream) {
  using Tr = MapTraits<OP>;
  if (n == 0) return NOMPK_OK;
  T *y = static_cast<T *>(y_);
  const T *x = static_cast<const T *>(x_);
  const T *z = static_cast<const T *>(z_);
  if (!y || (Tr::uses_x && !x) || (Tr::uses_z && !z)) {
    set_error("nompk_map: missing operand pointer for op %d", OP);
    return NOMPK_EINVAL;
  }
  const T alpha = alpha_ ? *static_cast<const T *>(alpha_) : T(1);
  const T beta = beta_ ? *static_cast<const T *>(beta_) : T(1);

  constexpr int L = Vec16<T>::kLanes;
  const int sms = sm_count();
  const bool aligned = is_aligned16(y) && (!Tr::uses_x || is_aligned16(x)) && (!Tr::uses_z || is_aligned16(z));
  if (!aligned) {
    size_t blocks = (n + kBlock - 1) / kBlock;
    const size_t cap = (size_t)sms * 8;
    if (blocks > cap) blocks = cap;
    map_scalar_kernel<OP, T><<<(unsigned)blocks, kBlock, 0, stream>>>(y, x, z, alpha, beta, n);
    NOMPK_LAUNCH_CHECK("map_scalar_kernel");
    return NOMPK_OK;
  }

  const size_t nvec = n / L;
  // one vector per thread, one tile per CTA (see the header comment); 1024-thread CTAs once every SM has >= 2 of them
  if (nvec >= (size_t)sms * 2 * kBigBlock) {
    const size_t blocks = (nvec + kBigBlock - 1) / kBigBlock;
    map_vec_kernel<OP, T, 1, kBigBlock><<<(unsigned)blocks, kBigBlock, 0, stream>>>(y, x, z, alpha, beta, nvec, n);
  } else {
    size_t blocks = (nvec + kBlock - 1) / kBlock;
    if (blocks == 0) blocks = 1;
    map_vec_kernel<OP, T, 1, kBlock><<<(unsigned)blocks, kBlock, 0, stream>>>(y, x, z, alpha, beta, nvec, n);
  }
  NOMPK_LAUNCH_CHECK("map_vec_kernel");
  return NOMPK_OK;
}

typedef int (*map_fn)(size_t, void *, const void *, const void *, const void *, const void *, cudaStream_t);

// Integer add/sub/mul are ring operations: the signed and unsigned variants have identical bit patterns,
// so i32 runs as u32 and i64 as u64 (this also sidesteps signed-overflow UB).
template <int OP> map_fn pick_dtype(nompk_dtype_t dt) {
  switch (dt) {
  case NOMPK_I32:
  case NOMPK_U32: return launch_map<OP, unsigned int>;
  case NOMPK_I64:
  case NOMPK_U64: return launch_map<OP, unsigned long long>;
  case NOMPK_F32: return launch_map<OP, float>;
  case NOMPK_F64: return launch_map<OP, double>;
  }
  return nullptr;
}

}  // namespace
}  // namespace nompk

extern "C" int nompk_map(nompk_map_op_t op, nompk_dtype_t dt, size_t n, void *y, const void *x, const void *z,
                         const void *alpha_host, const void *beta_host, void *stream) {
  using namespace nompk;
  map_fn fn = nullptr;
  switch (op) {
  case NOMPK_MAP_ADD: fn = pick_dtype<NOMPK_MAP_ADD>(dt); break;
  case NOMPK_MAP_SUB: fn = pick_dtype<NOMPK_MAP_SUB>(dt); break;
  case NOMPK_MAP_MUL: fn = pick_dtype<NOMPK_MAP_MUL>(dt); break;
  case NOMPK_MAP_AXPY: fn = pick_dtype<NOMPK_MAP_AXPY>(dt); break;
  case NOMPK_MAP_XPAY: fn = pick_dtype<NOMPK_MAP_XPAY>(dt); break;
  case NOMPK_MAP_AXPBY: fn = pick_dtype<NOMPK_MAP_AXPBY>(dt); break;
  case NOMPK_MAP_SCALE: fn = pick_dtype<NOMPK_MAP_SCALE>(dt); break;
  case NOMPK_MAP_COPY: fn = pick_dtype<NOMPK_MAP_COPY>(dt); break;
  case NOMPK_MAP_FILL: fn = pick_dtype<NOMPK_MAP_FILL>(dt); break;
  case NOMPK_MAP_ADD3: fn = pick_dtype<NOMPK_MAP_ADD3>(dt); break;
  default: break;
  }
  if (!fn) {
    set_error("nompk_map: unsupported op %d / dtype %d", (int)op, (int)dt);
    return NOMPK_EINVAL;
  }
  return fn(n, y, x, z, alpha_host, beta_host, static_cast<cudaStream_t>(stream));
}
