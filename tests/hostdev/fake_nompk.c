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

/* While the CUDA double captures a stream (fake_cuda.c), a call is recorded with COPIES of its by-value arguments -- host
 * scalars included, as a real launch copies its kernel parameters -- and runs when the graph is launched. */
int nomp_hostdev_record(int (*fn)(void *), const void *blob, size_t bytes);

typedef struct {
  int op, dt;
  size_t n;
  void *y;
  const void *x, *z;
  unsigned long long alpha, beta;
  int has_alpha, has_beta;
} map_call_t;

static int run_map(void *blob) {
  map_call_t *c = (map_call_t *)blob;
  int (*f)(int, int, size_t, void *, const void *, const void *, const void *, const void *) = oracle_sym("oracle_map");
  return f(c->op, c->dt, c->n, c->y, c->x, c->z, c->has_alpha ? &c->alpha : NULL, c->has_beta ? &c->beta : NULL);
}

EXPORT int nompk_map(nompk_map_op_t op, nompk_dtype_t dt, size_t n, void *y, const void *x, const void *z,
                     const void *alpha_host, const void *beta_host, void *stream) {
  (void)stream;
  map_call_t c = {(int)op, (int)dt, n, y, x, z, 0, 0, alpha_host != NULL, beta_host != NULL};
  const size_t w = nompk_dtype_size(dt);
  if (alpha_host) memcpy(&c.alpha, alpha_host, w);
  if (beta_host) memcpy(&c.beta, beta_host, w);
  calls++;
  if (nomp_hostdev_record(run_map, &c, sizeof(c))) return NOMPK_OK;
  return run_map(&c) ? NOMPK_EINVAL : NOMPK_OK;
}

EXPORT size_t nompk_dtype_size(nompk_dtype_t dt) { return dt == NOMPK_I64 || dt == NOMPK_U64 || dt == NOMPK_F64 ? 8 : 4; }

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

typedef struct {
  int op, dt;
  void *value, *result_host;
  unsigned long long host_seq;
  nompk_peers_t peers;
} allreduce_call_t;

static int run_allreduce(void *blob) {
  allreduce_call_t *c = (allreduce_call_t *)blob;
  unsigned long long mine = 0;
  memcpy(&mine, c->value, 8);
  return finish_with_peers(c->op, c->dt, &mine, c->value, c->result_host, c->host_seq, c->peers.peer_xchg, c->peers.rank,
                           c->peers.world, c->peers.seq_dev ? ++*c->peers.seq_dev : c->peers.seq);
}

EXPORT int nompk_allreduce_scalar_peers(nompk_red_op_t op, nompk_dtype_t dt, void *value, void *result_host_mapped,
                                        unsigned long long host_seq, const nompk_peers_t *peers, void *stream) {
  (void)stream;
  allreduce_call_t c = {(int)op, (int)dt, value, result_host_mapped, host_seq, *peers};
  calls++;
  if (nomp_hostdev_record(run_allreduce, &c, sizeof(c))) return NOMPK_OK; /* captured: runs at every replay */
  return run_allreduce(&c);
}

static void publish(const void *value, void *result, void *result_host_mapped, unsigned long long host_seq) {
  memcpy(result, value, 8);
  if (result_host_mapped) {
    memcpy(result_host_mapped, value, 8);
    __atomic_thread_fence(__ATOMIC_RELEASE);
    memcpy((char *)result_host_mapped + 8, &host_seq, 8);
  }
}

typedef struct {
  int op, dt, n_ax; /* n_ax != 0: the Ax + dot call */
  size_t n;
  const void *x, *y, *g, *D;
  void *w, *result, *result_host;
  unsigned long long host_seq;
  nompk_peers_t peers;
  const double *r, *beta_dev; /* r != NULL: x <- r + beta x first */
  double beta;
} reduce_call_t;

static int run_reduce(void *blob) {
  reduce_call_t *c = (reduce_call_t *)blob;
  unsigned long long value = 0;
  if (c->r) { /* the direction update through the oracle's map (NOMPK_MAP_XPAY: y <- x + alpha y) */
    int (*map)(int, int, size_t, void *, const void *, const void *, const void *, const void *) = oracle_sym("oracle_map");
    const double beta = c->beta_dev ? c->beta_dev[0] : c->beta;
    if (map(NOMPK_MAP_XPAY, NOMPK_F64, c->n * (size_t)c->n_ax * c->n_ax * c->n_ax, (void *)c->x, c->r, NULL, &beta, NULL)) return NOMPK_EINVAL;
  }
  if (c->n_ax) {
    int (*ax)(int, size_t, const double *, const double *, const double *, double *) = oracle_sym("oracle_ax_f64");
    if (ax(c->n_ax, c->n, c->x, c->g, c->D, c->w)) return NOMPK_EINVAL;
    double s = 0.0;
    const double *u = c->x, *w = c->w;
    for (size_t i = 0; i < c->n * (size_t)c->n_ax * c->n_ax * c->n_ax; i++) s += u[i] * w[i];
    memcpy(&value, &s, 8);
  } else {
    int (*f)(int, int, size_t, const void *, const void *, void *) = oracle_sym("oracle_reduce");
    if (f(c->op, c->dt, c->n, c->x, c->y, &value)) return NOMPK_EINVAL;
  }
  if (c->peers.world > 1) /* the call number is taken when the launch EXECUTES (a replayed graph takes a new one) */
    return finish_with_peers(c->op, c->dt, &value, c->result, c->result_host, c->host_seq, c->peers.peer_xchg, c->peers.rank,
                             c->peers.world, c->peers.seq_dev ? ++*c->peers.seq_dev : c->peers.seq);
  publish(&value, c->result, c->result_host, c->host_seq);
  return NOMPK_OK;
}

EXPORT int nompk_reduce_peers(nompk_red_op_t op, nompk_dtype_t dt, size_t n, const void *x, const void *y, void *result,
                              void *result_host_mapped, unsigned long long host_seq, void *workspace,
                              const nompk_peers_t *peers, void *stream) {
  (void)workspace, (void)stream;
  reduce_call_t c = {(int)op, (int)dt, 0, n, x, y, NULL, NULL, NULL, result, result_host_mapped, host_seq, {NULL, 0, 1, 0, NULL, NULL}, NULL, NULL, 0.0};
  if (peers && peers->world > 1) c.peers = *peers;
  calls++;
  if (nomp_hostdev_record(run_reduce, &c, sizeof(c))) return NOMPK_OK;
  return run_reduce(&c);
}

typedef struct {
  int n;
  size_t E;
  const double *u, *g, *D;
  double *w;
} ax_call_t;

static int run_ax(void *blob) {
  ax_call_t *c = (ax_call_t *)blob;
  int (*f)(int, size_t, const double *, const double *, const double *, double *) = oracle_sym("oracle_ax_f64");
  return f(c->n, c->E, c->u, c->g, c->D, c->w) ? NOMPK_EINVAL : NOMPK_OK;
}

/* the flags of the last Ax call of any kind (what the backend found out about D): tests/test_hostdev_cpu.py */
static unsigned last_ax_flags;
EXPORT unsigned nomp_hostdev_last_ax_flags(void) { return last_ax_flags; }

EXPORT int nompk_ax_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w, unsigned flags,
                        void *stream) {
  (void)stream;
  last_ax_flags = flags;
  ax_call_t c = {n, E, u, g, D, w};
  calls++;
  if (nomp_hostdev_record(run_ax, &c, sizeof(c))) return NOMPK_OK;
  return run_ax(&c);
}

EXPORT int nompk_ax_dot_peers_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w,
                                  double *result, double *result_host_mapped, unsigned long long host_seq,
                                  void *workspace, const nompk_peers_t *peers, unsigned flags, void *stream) {
  (void)workspace, (void)stream;
  last_ax_flags = flags;
  reduce_call_t c = {NOMPK_RED_SUM, NOMPK_F64, n, E, u, NULL, g, D, w, result, result_host_mapped, host_seq, {NULL, 0, 1, 0, NULL, NULL}, NULL, NULL, 0.0};
  if (peers && peers->world > 1) c.peers = *peers;
  calls++;
  if (nomp_hostdev_record(run_reduce, &c, sizeof(c))) return NOMPK_OK;
  return run_reduce(&c);
}

EXPORT int nompk_ax_xpay_dot_peers_f64(int n, size_t E, double *p, const double *r, double beta, const double *beta_dev,
                                       const double *g, const double *D, double *w, double *result, double *result_host_mapped,
                                       unsigned long long host_seq, void *workspace, const nompk_peers_t *peers, unsigned flags,
                                       void *stream) {
  (void)workspace, (void)stream;
  last_ax_flags = flags;
  reduce_call_t c = {NOMPK_RED_SUM, NOMPK_F64, n, E, p, NULL, g, D, w, result, result_host_mapped, host_seq, {NULL, 0, 1, 0, NULL, NULL}, r, beta_dev, beta};
  if (peers && peers->world > 1) c.peers = *peers;
  calls++;
  if (nomp_hostdev_record(run_reduce, &c, sizeof(c))) return NOMPK_OK;
  return run_reduce(&c);
}
