// Library-level state of libnompk: error text, SM count cache, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "nompk_common.cuh"

namespace nompk {

static thread_local char g_error[1024] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kDefaultSMs;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kDefaultSMs;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace nompk

extern "C" int nompk_version(void) { return NOMPK_VERSION; }
extern "C" const char *nompk_last_error(void) { return nompk::g_error; }
extern "C" unsigned long long nompk_launch_count(void) { return nompk::g_launches.load(); }
extern "C" size_t nompk_dtype_size(nompk_dtype_t dt) {
  switch (dt) {
  case NOMPK_I32:
  case NOMPK_U32:
  case NOMPK_F32: return 4;
  case NOMPK_I64:
  case NOMPK_U64:
  case NOMPK_F64: return 8;
  }
  return 0;
}
