/* TEST DOUBLE -- not part of the product (see fake_cuda.c).
 *
 * Interposes the compute entry points of libnompk (include/nompk.h) that libnomp's backend calls for the native loop
 * families, and answers them with the CPU oracle (oracle/libnomp_oracle.so, found through NOMP_HOSTDEV_ORACLE) on the
 * "device" memory of the CUDA test double.  What this exercises is the RUNTIME: family recognition in the bridge, the
 * kernel descriptor, the binding of nomp_jit arguments to operands by name, operation and type codes, the {value,
 * sequence number} publication protocol of the reduce finish.  It says nothing about the kernels themselves -- those are
 * executed from their own text by tests/test_device_mapreduce_cpu.py, test_device_ax_cpu.py and test_device_gs_cpu.py,
 * and on the GPU by the `-m gpu` tier.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nompk.h"

#define EXPORT __attribute__((visibility("default")))

static unsigned long long calls = 0;

static void *oracle_sym(const char *name) {
  static void *lib = NULL;
  if (!lib) {
    const char *path = getenv("NOMP_HOSTDEV_ORACLE");
    lib = path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : NULL;
    if (!lib) {
      fprintf(stderr, "hostdev: NOMP_HOSTDEV_ORACLE does not name the oracle library (%s)\n", path ? dlerror() : "unset");
      abort();
    }
  }
  void *f = dlsym(lib, name);
  if (!f) {
    fprintf(stderr, "hostdev: the oracle has no %s\n", name);
    abort();
  }
  return f;
}

EXPORT unsigned long long nompk_launch_count(void) { return calls; }

EXPORT int nompk_map(nompk_map_op_t op, nompk_dtype_t dt, size_t n, void *y, const void *x, const void *z,
                     const void *alpha_host, const void *beta_host, void *stream) {
  (void)stream;
  int (*f)(int, int, size_t, void *, const void *, const void *, const void *, const void *) = oracle_sym("oracle_map");
  calls++;
  return f((int)op, (int)dt, n, y, x, z, alpha_host, beta_host) ? NOMPK_EINVAL : NOMPK_OK;
}

/* The exchange between ranks: the product's own finish_result (nompk_gridreduce.cuh) on the emulator, one warp
 * (tests/hostdev/build_devicecode.py).  `value` is this rank's contribution; result / result_host receive the fold. */
static int finish_with_peers(int op, int dt, const void *value, void *result, void *result_host, unsigned long long host_seq,
                             void *const *peer_xchg, int rank, int world, unsigned long long cseq) {
  static int (*launch)(unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, int, int, const void *, void *, void *,
                       unsigned long long, void *const *, int, int, unsigned long long) = NULL;
  if (!launch) {
    const char *path = getenv("NOMP_HOSTDEV_DEVICECODE");
    void *lib = path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : NULL;
    if (lib) *(void **)&launch = dlsym(lib, "nomp_emu_launch");
    if (!launch) {
      fprintf(stderr, "hostdev: NOMP_HOSTDEV_DEVICECODE does not name the device-code library\n");
      abort();
    }
  }
  return launch(1, 1, 1, 32, 1, 1, op, dt, value, result, result_host, host_seq, peer_xchg, rank, world, cseq) ? NOMPK_ECUDA : NOMPK_OK;
}

EXPORT size_t nompk_allreduce_xchg_bytes(int world) { return (size_t)2 * (size_t)world * 16; } /* as reduce.cu */

/* Stand-alone all-reduce of a device scalar, in place: the same protocol on the same buffers (reduce.cu). */
EXPORT int nompk_allreduce_scalar(nompk_red_op_t op, nompk_dtype_t dt, void *value, void *result_host_mapped,
                                  unsigned long long host_seq, void *const *peer_xchg, int rank, int world,
                                  unsigned long long seq, void *stream) {
  (void)stream;
  unsigned long long mine = 0;
  memcpy(&mine, value, 8);
  calls++;
  return finish_with_peers((int)op, (int)dt, &mine, value, result_host_mapped, host_seq, peer_xchg, rank, world, seq);
}

static void publish(const void *value, void *result, void *result_host_mapped, unsigned long long host_seq) {
  memcpy(result, value, 8);
  if (result_host_mapped) {
    memcpy(result_host_mapped, value, 8);
    __atomic_thread_fence(__ATOMIC_RELEASE);
    memcpy((char *)result_host_mapped + 8, &host_seq, 8);
  }
}

EXPORT int nompk_reduce_peers(nompk_red_op_t op, nompk_dtype_t dt, size_t n, const void *x, const void *y, void *result,
                              void *result_host_mapped, unsigned long long host_seq, void *workspace,
                              const nompk_peers_t *peers, void *stream) {
  (void)workspace, (void)stream;
  int (*f)(int, int, size_t, const void *, const void *, void *) = oracle_sym("oracle_reduce");
  unsigned long long value = 0;
  if (f((int)op, (int)dt, n, x, y, &value)) return NOMPK_EINVAL;
  calls++;
  if (peers && peers->world > 1)
    return finish_with_peers((int)op, (int)dt, &value, result, result_host_mapped, host_seq, peers->peer_xchg, peers->rank,
                             peers->world, peers->seq);
  publish(&value, result, result_host_mapped, host_seq);
  return NOMPK_OK;
}

EXPORT int nompk_ax_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w, unsigned flags,
                        void *stream) {
  (void)flags, (void)stream;
  int (*f)(int, size_t, const double *, const double *, const double *, double *) = oracle_sym("oracle_ax_f64");
  calls++;
  return f(n, E, u, g, D, w) ? NOMPK_EINVAL : NOMPK_OK;
}

EXPORT int nompk_ax_dot_peers_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w,
                                  double *result, double *result_host_mapped, unsigned long long host_seq,
                                  void *workspace, const nompk_peers_t *peers, unsigned flags, void *stream) {
  (void)workspace;
  if (nompk_ax_f64(n, E, u, g, D, w, flags, stream)) return NOMPK_EINVAL;
  double s = 0.0;
  for (size_t i = 0; i < E * (size_t)n * n * n; i++) s += u[i] * w[i];
  if (peers && peers->world > 1)
    return finish_with_peers(NOMPK_RED_SUM, NOMPK_F64, &s, result, result_host_mapped, host_seq, peers->peer_xchg, peers->rank,
                             peers->world, peers->seq);
  publish(&s, result, result_host_mapped, host_seq);
  return NOMPK_OK;
}
