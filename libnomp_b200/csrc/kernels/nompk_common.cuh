// Shared device/host helpers for the libnompk kernel families (map.cu, reduce.cu, ax.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "nompk.h"

namespace nompk {

// B200: 148 SMs (2 dies x 74).  Queried once; the constant is only a fallback for the (impossible on the
// target) case where the attribute query fails.
constexpr int kDefaultSMs = 148;

int sm_count();
void set_error(const char *fmt, ...);
void count_launch();

#define NOMPK_CUDA_TRY(call)                                                                       \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      nompk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,         \
                       __LINE__);                                                                  \
      return NOMPK_ECUDA;                                                                          \
    }                                                                                              \
  } while (0)

#define NOMPK_LAUNCH_CHECK(what)                                                                   \
  do {                                                                                             \
    cudaError_t e_ = cudaGetLastError();                                                           \
    if (e_ != cudaSuccess) {                                                                       \
      nompk::set_error("launch of %s failed: %s", what, cudaGetErrorString(e_));                  \
      return NOMPK_ECUDA;                                                                          \
    }                                                                                              \
    nompk::count_launch();                                                                         \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// Arithmetic without FMA contraction.  The reference semantics of a nomp kernel is its C string run
// serially; nvcc would fuse a*b+c into one rounding.  These helpers keep every rounding of
// the C expression so that map results are bit-identical to the serial loop.
// ---------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T op_add(T a, T b) { return a + b; }
template <typename T> __device__ __forceinline__ T op_sub(T a, T b) { return a - b; }
template <typename T> __device__ __forceinline__ T op_mul(T a, T b) { return a * b; }
template <> __device__ __forceinline__ float op_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ float op_sub<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ float op_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double op_add<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ double op_sub<double>(double a, double b) { return __dsub_rn(a, b); }
template <> __device__ __forceinline__ double op_mul<double>(double a, double b) { return __dmul_rn(a, b); }

// ---------------------------------------------------------------------------------------------------
// 128-bit vectors of T.  All bulk traffic of the map and reduce families moves as LDG.E.128 / STG.E.128.
// ---------------------------------------------------------------------------------------------------
template <typename T> struct alignas(16) Vec16 {
  static constexpr int kLanes = 16 / sizeof(T);
  T v[kLanes];
};

template <typename T> __device__ __forceinline__ Vec16<T> ld_vec(const T *p) {
  static_assert(sizeof(Vec16<T>) == 16, "Vec16 must be 16 bytes");
  int4 raw = *reinterpret_cast<const int4 *>(p);
  return *reinterpret_cast<Vec16<T> *>(&raw);
}

// Read-only streaming load: ld.global.nc (LDG.E.128.CONSTANT); used for operands no thread writes.
template <typename T> __device__ __forceinline__ Vec16<T> ld_vec_ro(const T *p) {
  int4 raw = __ldg(reinterpret_cast<const int4 *>(p));
  return *reinterpret_cast<Vec16<T> *>(&raw);
}

template <typename T> __device__ __forceinline__ void st_vec(T *p, const Vec16<T> &v) {
  *reinterpret_cast<int4 *>(p) = *reinterpret_cast<const int4 *>(&v);
}

__host__ __device__ inline bool is_aligned16(const void *p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

}  // namespace nompk
