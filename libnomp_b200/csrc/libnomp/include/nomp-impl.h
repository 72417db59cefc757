/* Internal types of the libnomp runtime: configuration, kernel arguments, programs, mappings and the backend
 * vtable.  The vtable (struct nomp_backend) keeps the reference's member names and function signatures
 * (reference include/nomp-impl.h:208-262) -- it is the internal drop-in boundary: core code calls
 * update / knl_build / knl_run / knl_free / sync / finalize and nothing else of a backend. */
#ifndef LIBNOMP_B200_IMPL_H_
#define LIBNOMP_B200_IMPL_H_

#define _POSIX_C_SOURCE 200809L
#define _GNU_SOURCE

#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <limits.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nomp-defs.h"
#include "nomp-log.h"
#include "nomp-mem.h"
#include "nomp.h"

#define NOMP_EXPORT __attribute__((visibility("default")))

typedef struct {
  int platform, device, verbose, profile;
  char backend[NOMP_MAX_BUFFER_SIZE + 1];
  char install_dir[PATH_MAX + 1];
  char scripts_dir[PATH_MAX + 1];
  char annotations_script[NOMP_MAX_BUFFER_SIZE + 1];
} nomp_config_t;

/* One runtime argument of a program (reference include/nomp-impl.h:69-86). */
typedef struct {
  char name[NOMP_MAX_BUFFER_SIZE + 1];
  size_t size;          /* sizeof() of the C object */
  nomp_arg_type_t type; /* NOMP_INT | NOMP_UINT | NOMP_FLOAT | NOMP_PTR */
  void *ptr;            /* per run: device pointer (NOMP_PTR) or the caller's pointer to the scalar */
  int is_const;         /* pointer argument declared const in the kernel source (read-only) */
  void *mem;            /* per run: the nomp_mem_t behind a mapped pointer argument, else NULL */
} nomp_arg_t;

/* NOMP_SUM / NOMP_PROD are the reference's operators (reference include/nomp-impl.h:99-102); MIN / MAX are new. */
typedef enum { NOMP_SUM = 0, NOMP_PROD = 1, NOMP_MIN = 2, NOMP_MAX = 3 } nomp_reduction_op_t;

/* A jitted kernel (reference include/nomp-impl.h:109-168).  The SymEngine vectors of the reference are replaced
 * by expression strings evaluated by src/gridexpr.c. */
typedef struct {
  unsigned nargs;
  nomp_arg_t *args;
  unsigned ndim;
  char *sym_global[3], *sym_local[3]; /* launch-size expressions over the integer arguments */
  int eval_grid;
  long int_values[NOMP_MAX_KERNEL_ARGS_SIZE]; /* last seen value of each integer argument */
  int int_valid;
  size_t global[3], local[3], gws[3];
  void *bptr; /* backend payload (module + function, or a native dispatch record) */
  int reduction_index;
  nomp_reduction_op_t reduction_op;
  nomp_arg_type_t reduction_type;
  int reduction_size;
  void *reduction_ptr;
  void *reduction_dev; /* device copy of the reduction variable when its result stays on the device
                          (nomp_b200_device_reductions), else NULL */
  PyObject *py_dict; /* JIT-fixed arguments: name -> value */
  char *info;        /* descriptor line of the generated kernel (nomp_b200_prog_info) */
} nomp_prog_t;

/* A host range mirrored on the device (reference include/nomp-impl.h:175-200). */
typedef struct {
  size_t idx0, idx1, usize;
  void *hptr;
  void *bptr;
  size_t bsize;
  unsigned long version; /* changes whenever the device image may have changed; values come from one process-wide
                            counter, so a new mapping at a recycled device address never repeats an old one */
  unsigned long long d2h_ticket; /* number of the last asynchronous device-to-host copy out of this mapping */
  unsigned transfers;    /* host <-> device copies of this mapping so far */
  int host_registered;   /* the backend page-locked the host range (cudaHostRegister) and must release it */
} nomp_mem_t;

struct nomp_backend {
  int (*update)(struct nomp_backend *, nomp_mem_t *, const nomp_map_direction_t op, size_t start, size_t end,
                size_t usize);
  int (*knl_build)(struct nomp_backend *, nomp_prog_t *, const char *source, const char *name);
  int (*knl_run)(struct nomp_backend *, nomp_prog_t *);
  int (*knl_free)(nomp_prog_t *);
  int (*sync)(struct nomp_backend *);
  int (*finalize)(struct nomp_backend *);
  nomp_mem_t scratch;
  PyObject *py_annotate;
  PyObject *py_context;
  void *bptr;
};
typedef struct nomp_backend nomp_backend_t;

#define NOMP_MEM_OFFSET(start, usize) ((start) * (usize))
#define NOMP_MEM_BYTES(start, end, usize) (((end) - (start)) * (usize))

#ifdef __cplusplus
extern "C" {
#endif

/* The one backend of this implementation (reference include/nomp-impl.h:319-323 lists opencl/cuda/hip). */
int cuda_init(nomp_backend_t *backend, int platform, int device);

/* Reduce-clause finish: src/reduction.c (type bookkeeping) -> backend (all-reduce across ranks, wait, store the
 * result through prg->reduction_ptr).  Replaces nomp_host_side_reduction (reference src/reduction.c:33-88). */
int nomp_device_side_reduction(nomp_backend_t *backend, nomp_prog_t *prg);
int nomp_cuda_reduction_finish(nomp_backend_t *backend, nomp_prog_t *prg, int dtype, size_t size);

int nomp_cuda_update_async(nomp_backend_t *backend, nomp_mem_t *m, nomp_map_direction_t op, size_t start, size_t end,
                           size_t usize);

/* Orders the compute stream behind an asynchronous device-to-host copy that may still be reading `m`; called before
 * anything that writes the mapping (kernels with a non-const pointer to it, nomp_update(TO), the gather-scatter). */
int nomp_cuda_before_write(nomp_mem_t *m);

/* core helpers used by the backend */
nomp_mem_t *nomp_lookup_mem(const void *hptr);

/* launch-size expressions (src/gridexpr.c) */
int nomp_gridexpr_eval(const char *expr, const char *const *names, const long *values, unsigned n, long *result);

/* NCCL communicator (src/comm.c) */
int nomp_comm_init(int device);
int nomp_comm_finalize(void);
int nomp_comm_rank(void);
int nomp_comm_size(void);
/* in-place allreduce of one scalar on `stream`; dtype codes of include/nompk.h, op = nomp_reduction_op_t.
 * *published = 1 if {value, host_seq} was also written to result_host_mapped (NVLink one-shot path). */
int nomp_comm_allreduce(void *dev_scalar, int dtype, int op, void *result_host_mapped, unsigned long long host_seq,
                        void *error_host_mapped, void *stream, int *published);
int nomp_comm_peers(void *peers /* nompk_peers_t * */, void *error_host_mapped);
unsigned long nomp_next_version(void);
int nomp_b200_exchange_blob(const char *path, int rank, void *blob, size_t bytes);
/* all[r] <- rank r's `bytes`-byte record, through files "<id file>.<tag>.<r>" (small setup-time records only) */
int nomp_comm_allgather(const char *tag, const void *mine, void *all, size_t bytes);
int nomp_comm_barrier(void);

/* gather-scatter handles (src/gs.c) */
void nomp_gs_finalize(void);
int nomp_gs_check(void);

/* on-disk JIT cache (src/jitcache.c) */
typedef struct {
  uint32_t h[8];
  uint64_t len;
  unsigned fill;
  unsigned char buf[64];
} nomp_sha256_t;
void nomp_sha256_init(nomp_sha256_t *c);
void nomp_sha256_update(nomp_sha256_t *c, const void *data, size_t n);
void nomp_sha256_field(nomp_sha256_t *c, const char *s);
int nomp_sha256_file(nomp_sha256_t *c, const char *path);
int nomp_sha256_dir(nomp_sha256_t *c, const char *dir, const char *suffix);
void nomp_sha256_hex(nomp_sha256_t *c, char hex[65]);
const char *nomp_jit_cache_dir(void); /* NULL when the cache is off */
void nomp_jit_cache_reset(void);
enum { NOMP_CACHE_KNL_HIT = 0, NOMP_CACHE_KNL_MISS = 1, NOMP_CACHE_CUBIN_HIT = 2, NOMP_CACHE_CUBIN_MISS = 3 };
void nomp_jit_cache_count(int which);
int nomp_jit_cache_load(const char *hex, const char *ext, char **data, size_t *size);
int nomp_jit_cache_store(const char *hex, const char *ext, const void *data, size_t size);

extern const char *ERR_STR_USER_MAP_PTR_IS_INVALID;
extern const char *ERR_STR_USER_DEVICE_IS_INVALID;

#ifdef __cplusplus
}
#endif

#endif
