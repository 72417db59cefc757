/* CUDA-only backend of libnomp for B200 (sm_100a).
 *
 * Fills the same six-entry vtable as the reference's CUDA backend (reference backends/cuda.c +
 * backends/unified-cuda-hip-impl.h:70-246; vtable at reference include/nomp-impl.h:213-237) but
 *   - knl_build() understands a one-line descriptor in front of the source: `kind=native` programs are bound to
 *     the hand-written kernels of libnompk.so (include/nompk.h), `kind=nvrtc` programs are compiled by NVRTC
 *     straight to a CUBIN for sm_100a (the reference asks NVRTC for PTX of sm_100 and lets the driver JIT it,
 *     reference backends/cuda.c:21-22, unified-cuda-hip-impl.h:101-123) with FMA contraction off, so generated
 *     code keeps the roundings of the C loop;
 *   - all work is issued on one non-blocking stream owned by the backend (the reference uses the NULL stream and
 *     cudaDeviceSynchronize, unified-cuda-hip-impl.h:151-169);
 *   - a reduce clause is one kernel whose last block writes the result into mapped pinned host memory; across
 *     ranks the scalar is all-reduced with NCCL first (reference: D2H of every partial + serial host loop,
 *     src/reduction.c:33-88);
 *   - there is no HIP / OpenCL / CPU path: if CUDA is unavailable nomp_init() fails with NOMP_CUDA_FAILURE.
 * The driver API (module load, launch) is reached through cudaGetDriverEntryPoint, so the library has no
 * link-time dependency on libcuda.so and loads on machines without a driver.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>

#include "nomp-impl.h"
#include "nomp-loopy.h"
#include "nompk.h"

static const char *ERR_STR_CUDA_FAILURE = "CUDA %s failure: %s.";

#define check_runtime(call)                                                                                      \
  do {                                                                                                           \
    cudaError_t e_ = (call);                                                                                     \
    if (e_ != cudaSuccess)                                                                                       \
      return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "runtime", cudaGetErrorName(e_));     \
  } while (0)

#define check_driver(call)                                                                                       \
  do {                                                                                                           \
    CUresult r_ = (call);                                                                                        \
    if (r_ != CUDA_SUCCESS) {                                                                                    \
      const char *m_ = "unknown";                                                                                \
      if (drv.GetErrorName) drv.GetErrorName(r_, &m_);                                                           \
      return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "driver", m_);                        \
    }                                                                                                            \
  } while (0)

#define check_nvrtc(call)                                                                                        \
  do {                                                                                                           \
    nvrtcResult r_ = (call);                                                                                     \
    if (r_ != NVRTC_SUCCESS)                                                                                     \
      return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "nvrtc", nvrtcGetErrorString(r_));    \
  } while (0)

#define check_nompk(call)                                                                                        \
  do {                                                                                                           \
    int r_ = (call);                                                                                             \
    if (r_ != NOMPK_OK)                                                                                          \
      return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "kernel library", nompk_last_error()); \
  } while (0)

/* driver entry points, resolved once per process */
static struct {
  CUresult (*ModuleLoadData)(CUmodule *, const void *);
  CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *);
  CUresult (*ModuleUnload)(CUmodule);
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                           CUstream, void **, void **);
  CUresult (*GetErrorName)(CUresult, const char **);
  int resolved;
} drv;

static int resolve_driver(void) {
  if (drv.resolved) return 0;
  struct {
    const char *name;
    void **slot;
  } table[] = {{"cuModuleLoadData", (void **)&drv.ModuleLoadData},
               {"cuModuleGetFunction", (void **)&drv.ModuleGetFunction},
               {"cuModuleUnload", (void **)&drv.ModuleUnload},
               {"cuLaunchKernel", (void **)&drv.LaunchKernel},
               {"cuGetErrorName", (void **)&drv.GetErrorName}};
  for (unsigned i = 0; i < sizeof(table) / sizeof(table[0]); i++) {
    enum cudaDriverEntryPointQueryResult status;
    check_runtime(cudaGetDriverEntryPoint(table[i].name, table[i].slot, cudaEnableDefault, &status));
    if (status != cudaDriverEntryPointSuccess || *table[i].slot == NULL)
      return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "driver entry point", table[i].name);
  }
  drv.resolved = 1;
  return 0;
}

/* ---- backend state ------------------------------------------------------------------------------------------ */

#define NOMP_MAX_GRAPHS 64

typedef struct {
  int device;
  struct cudaDeviceProp prop;
  cudaStream_t stream;
  cudaStream_t stream_h2d, stream_d2h; /* copy streams of nomp_b200_update_async (created on first use) */
  cudaEvent_t ev_compute, ev_h2d, ev_d2h;
  unsigned long long d2h_issued, d2h_waited; /* asynchronous D2H copies issued / already ordered before the compute stream */
  int async_used;
  int fused_allreduce; /* the reduction kernel just launched all-reduces its result itself */
  void *nv_peers;      /* peer description handed to a generated reduction kernel (by address, cuLaunchKernel) */
  int nv_rank, nv_world;
  void *nv_cseq_dev, *nv_err_host;
  void *pinned_host;   /* 64 bytes of mapped pinned memory: [0,8) reduction result, [8,16) its sequence number */
  unsigned long long host_seq; /* sequence number of the last reduction issued */
  void *pinned_dev;    /* its device alias */
  void *red_ws;        /* reduction workspace of libnompk inside the core's scratch buffer */
  void *red_result;    /* device slot of the reduced scalar, right after the workspace */
  void *red_result_host;
  void *red_result_arg; /* where this launch stores its result: red_result, or the mapped reduction variable */
  /* D staging cache of the Ax family */
  const void *ax_D;
  unsigned long ax_D_version;
  int ax_n;
  /* ... and what was found out about a D: is it centro-antisymmetric (even-odd contractions, ax_D_antisymmetric) */
  const void *ax_sym_D;
  unsigned long ax_sym_version;
  int ax_sym_n, ax_sym;
  cudaStream_t stream_aux; /* created on first use */
  unsigned long capture_version0; /* mappings with a younger version were written inside the capture in progress */
  unsigned long long nvrtc_launches;
  /* CUDA graphs of nomp_run sequences (nomp_b200_graph_*) */
  int capturing;
  cudaGraphExec_t graphs[NOMP_MAX_GRAPHS];
} cuda_state_t;

static cuda_state_t *g_state = NULL; /* for the nomp_b200_* accessors */

typedef enum { FAM_NVRTC = 0, FAM_MAP, FAM_REDUCE, FAM_AX, FAM_AXDOT, FAM_AXXPAYDOT } family_t;

#define SLOT_NONE (-1)
#define SLOT_WS (-2)
#define SLOT_RESULT (-4)
#define SLOT_RESULT_HOST (-5)
#define SLOT_SEQ (-6)
#define SLOT_PEERS (-7)  /* fused all-reduce of a generated reduction: exchange-buffer table, rank, world, call number */
#define SLOT_RANK (-8)
#define SLOT_WORLD (-9)
#define SLOT_CSEQ (-10) /* the collective call counter in device memory */
#define SLOT_ERR (-11)  /* error word in mapped host memory */

typedef struct {
  family_t family;
  /* nvrtc */
  CUmodule module;
  CUfunction function;
  int nparams;
  int param_slot[NOMP_MAX_KERNEL_ARGS_SIZE + 10]; /* index into prg->args, or SLOT_* */
  int has_peers;                                  /* the kernel takes the peer description (SLOT_PEERS ...) */
  int is_reduce;
  /* native */
  int op, dtype, ax_n;
  int a_y, a_x, a_z, a_alpha, a_beta, a_n, a_out; /* argument indices (SLOT_NONE if unused) */
  int a_u, a_g, a_D, a_w, a_E;
  int a_r, beta_dev; /* FAM_AXXPAYDOT: the residual vector; beta is a pointer into device memory (else a host scalar) */
  long n_literal; /* trip count when it is a literal in the source (a_n == SLOT_NONE) */
} cuda_prog_t;

/* ---- memory --------------------------------------------------------------------------------------------------- */
/* Page-lock the host range of a mapping that is being copied for the SECOND time: pageable copies are staged by the
 * driver at a fraction of the PCIe rate, and a range that moves twice usually moves every time step.  One copy is not
 * worth the registration (about as slow as the pageable copy it would save).  Ranges below 1 MiB, ranges somebody
 * else already pinned, and NOMP_PIN_HOST=0 leave the host memory alone. */
#define NOMP_PIN_MIN_BYTES ((size_t)1 << 20)

static void pin_host_range(nomp_mem_t *m) {
  static int enabled = -1;
  if (enabled < 0) {
    const char *e = getenv("NOMP_PIN_HOST");
    enabled = !(e && e[0] == '0');
  }
  const size_t bytes = NOMP_MEM_BYTES(m->idx0, m->idx1, m->usize);
  if (!enabled || m->host_registered || bytes < NOMP_PIN_MIN_BYTES) return;
  m->host_registered = -1; /* tried: do not try again */
  if (cudaHostRegister((char *)m->hptr + NOMP_MEM_OFFSET(m->idx0, m->usize), bytes, cudaHostRegisterDefault) == cudaSuccess)
    m->host_registered = 1;
  else
    cudaGetLastError(); /* already pinned by its owner, or not registrable: copies stay pageable */
}

static void unpin_host_range(nomp_mem_t *m) {
  if (m->host_registered == 1) {
    if (cudaHostUnregister((char *)m->hptr + NOMP_MEM_OFFSET(m->idx0, m->usize)) != cudaSuccess) cudaGetLastError();
  }
  m->host_registered = 0;
}

/* While a graph is being captured (nomp_b200_graph_begin) the backend stream records work instead of executing it:
 * anything that would wait for the stream, or that is not stream-ordered, is refused. */
static int capture_forbids(const char *what) {
  return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "%s is not allowed while a graph is being captured.", what);
}

int nomp_cuda_before_write(nomp_mem_t *m) {
  cuda_state_t *st = g_state;
  if (st == NULL || m == NULL || m->d2h_ticket <= st->d2h_waited) return 0;
  /* ev_d2h was recorded behind the LAST asynchronous D2H copy; the copy stream is in order, so waiting for it covers
   * every earlier copy as well */
  check_runtime(cudaStreamWaitEvent(st->stream, st->ev_d2h, 0));
  st->d2h_waited = st->d2h_issued;
  return 0;
}

static int cuda_update(nomp_backend_t *bnd, nomp_mem_t *m, const nomp_map_direction_t op, size_t start, size_t end,
                       size_t usize) {
  cuda_state_t *st = (cuda_state_t *)bnd->bptr;
  if (st->capturing) return capture_forbids("nomp_update");
  if (((op & NOMP_TO) || op == NOMP_FROM) && m != &bnd->scratch && m->transfers++ == 1) pin_host_range(m);
  if (op & NOMP_ALLOC) {
    size_t bytes = NOMP_MEM_BYTES(start, end, usize);
    check_runtime(cudaMalloc(&m->bptr, bytes ? bytes : 1));
    m->bsize = bytes;
    if (m == &bnd->scratch) { /* reduction workspace: the ticket counter must start at zero */
      check_runtime(cudaMemsetAsync(m->bptr, 0, bytes, st->stream));
      if (bytes < nompk_reduce_workspace_bytes() + 64)
        return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "scratch", "buffer smaller than the reduction workspace");
      st->red_ws = m->bptr;
      st->red_result = (char *)m->bptr + nompk_reduce_workspace_bytes();
    }
  }
  if (op & NOMP_TO) {
    nomp_check(nomp_cuda_before_write(m));
    /* blocking on purpose: the caller may overwrite the host range as soon as nomp_update returns */
    check_runtime(cudaMemcpyAsync((char *)m->bptr + NOMP_MEM_OFFSET(start - m->idx0, usize),
                                  (const char *)m->hptr + NOMP_MEM_OFFSET(start, usize),
                                  NOMP_MEM_BYTES(start, end, usize), cudaMemcpyHostToDevice, st->stream));
    check_runtime(cudaStreamSynchronize(st->stream));
    m->version = nomp_next_version();
  }
  if (op == NOMP_FROM) {
    check_runtime(cudaMemcpyAsync((char *)m->hptr + NOMP_MEM_OFFSET(start, usize),
                                  (const char *)m->bptr + NOMP_MEM_OFFSET(start - m->idx0, usize),
                                  NOMP_MEM_BYTES(start, end, usize), cudaMemcpyDeviceToHost, st->stream));
    check_runtime(cudaStreamSynchronize(st->stream));
  } else if (op == NOMP_FREE) {
    check_runtime(cudaStreamSynchronize(st->stream));
    if (st->async_used) { /* asynchronous copies may still be using the device image and the page-locked host range */
      check_runtime(cudaStreamSynchronize(st->stream_h2d));
      check_runtime(cudaStreamSynchronize(st->stream_d2h));
    }
    if (st->ax_D && (const char *)st->ax_D >= (const char *)m->bptr && (const char *)st->ax_D < (const char *)m->bptr + (m->bsize ? m->bsize : 1))
      st->ax_D = NULL; /* the staged derivative matrix came from this mapping */
    if (st->ax_sym_D && (const char *)st->ax_sym_D >= (const char *)m->bptr &&
        (const char *)st->ax_sym_D < (const char *)m->bptr + (m->bsize ? m->bsize : 1))
      st->ax_sym_D = NULL;
    unpin_host_range(m);
    check_runtime(cudaFree(m->bptr));
    m->bptr = NULL; /* tells the core to drop the entry (reference src/nomp.c:360) */
  }
  return 0;
}

/* ---- descriptor parsing --------------------------------------------------------------------------------------- */
/* value of `key=` in the first line of src, copied into buf; 0 if absent */
static int desc_get(const char *src, const char *key, char *buf, size_t cap) {
  const char *eol = strchr(src, '\n');
  size_t linelen = eol ? (size_t)(eol - src) : strlen(src);
  size_t klen = strlen(key);
  for (const char *p = src; p + klen + 1 <= src + linelen; p++) {
    if ((p == src || p[-1] == ' ') && !strncmp(p, key, klen) && p[klen] == '=') {
      const char *v = p + klen + 1;
      size_t n = 0;
      while (v + n < src + linelen && v[n] != ' ' && n + 1 < cap) n++;
      memcpy(buf, v, n);
      buf[n] = '\0';
      return 1;
    }
  }
  return 0;
}

static int arg_index(const nomp_prog_t *prg, const char *name) {
  for (unsigned i = 0; i < prg->nargs; i++)
    if (!strncmp(prg->args[i].name, name, NOMP_MAX_BUFFER_SIZE)) return (int)i;
  return SLOT_NONE;
}

/* resolve descriptor key -> argument index; "-" or absent means unused */
static int desc_arg(const nomp_prog_t *prg, const char *src, const char *key, int required, int *out) {
  char name[NOMP_MAX_BUFFER_SIZE + 1];
  *out = SLOT_NONE;
  if (!desc_get(src, key, name, sizeof(name)) || !strcmp(name, "-")) {
    if (required)
      return nomp_log(NOMP_LOOPY_CODEGEN_FAILURE, NOMP_ERROR, "Kernel descriptor is missing \"%s\".", key);
    return 0;
  }
  *out = arg_index(prg, name);
  if (*out == SLOT_NONE)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR,
                    "Kernel argument \"%s\" was not declared in nomp_jit().", name);
  return 0;
}

static int desc_count(const nomp_prog_t *prg, const char *src, const char *key, int *idx, long *literal) {
  char v[NOMP_MAX_BUFFER_SIZE + 1];
  if (!desc_get(src, key, v, sizeof(v)))
    return nomp_log(NOMP_LOOPY_CODEGEN_FAILURE, NOMP_ERROR, "Kernel descriptor is missing \"%s\".", key);
  if (v[0] >= '0' && v[0] <= '9') {
    *idx = SLOT_NONE;
    *literal = strtol(v, NULL, 10);
    return 0;
  }
  *idx = arg_index(prg, v);
  if (*idx == SLOT_NONE)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR,
                    "Kernel argument \"%s\" was not declared in nomp_jit().", v);
  return 0;
}

static void mark_const_args(nomp_prog_t *prg, const char *src) {
  char list[BUFSIZ];
  if (!desc_get(src, "ro", list, sizeof(list))) return;
  for (char *tok = strtok(list, ","); tok; tok = strtok(NULL, ",")) {
    int i = arg_index(prg, tok);
    if (i >= 0) prg->args[i].is_const = 1;
  }
}

/* ---- build ---------------------------------------------------------------------------------------------------- */
static int build_nvrtc(cuda_state_t *st, cuda_prog_t *cp, nomp_prog_t *prg, const char *source, const char *name) {
  nomp_check(resolve_driver());

  char params[BUFSIZ];
  if (!desc_get(source, "params", params, sizeof(params))) params[0] = '\0';
  cp->nparams = 0;
  for (char *tok = strtok(params, ","); tok; tok = strtok(NULL, ",")) {
    int slot;
    if (!strcmp(tok, "nomp_ws")) slot = SLOT_WS;
    else if (!strcmp(tok, "nomp_result")) slot = SLOT_RESULT;
    else if (!strcmp(tok, "nomp_result_host")) slot = SLOT_RESULT_HOST;
    else if (!strcmp(tok, "nomp_seq")) slot = SLOT_SEQ;
    else if (!strcmp(tok, "nomp_peers")) slot = SLOT_PEERS, cp->has_peers = 1;
    else if (!strcmp(tok, "nomp_rank")) slot = SLOT_RANK;
    else if (!strcmp(tok, "nomp_world")) slot = SLOT_WORLD;
    else if (!strcmp(tok, "nomp_cseq_dev")) slot = SLOT_CSEQ;
    else if (!strcmp(tok, "nomp_err_host")) slot = SLOT_ERR;
    else {
      slot = arg_index(prg, tok);
      if (slot == SLOT_NONE)
        return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR,
                        "Kernel argument \"%s\" was not declared in nomp_jit().", tok);
    }
    if (cp->nparams >= NOMP_MAX_KERNEL_ARGS_SIZE + 10)
      return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Too many kernel arguments.");
    cp->param_slot[cp->nparams++] = slot;
  }
  char flag[8];
  cp->is_reduce = desc_get(source, "reduce", flag, sizeof(flag)) && flag[0] == '1';

  char arch[64];
  /* architecture-specific target ("a" suffix) for Hopper and newer: sm_100a on B200 */
  snprintf(arch, sizeof(arch), "--gpu-architecture=sm_%d%d%s", st->prop.major, st->prop.minor,
           st->prop.major >= 9 ? "a" : "");
  const char *options[] = {arch, "--fmad=false", "--std=c++17", "-lineinfo"};
  const int noptions = (int)(sizeof(options) / sizeof(options[0]));

  /* on-disk cache of the CUBIN (src/jitcache.c): key = source, entry name, options, NVRTC version */
  char key[65] = "", *cubin = NULL;
  size_t size = 0;
  int cacheable = nomp_jit_cache_dir() != NULL, major = 0, minor = 0;
  if (cacheable) {
    nomp_sha256_t c;
    char version[64];
    nvrtcVersion(&major, &minor);
    snprintf(version, sizeof(version), "libnomp_b200 cubin 1 nvrtc %d.%d", major, minor);
    nomp_sha256_init(&c);
    nomp_sha256_field(&c, version), nomp_sha256_field(&c, source), nomp_sha256_field(&c, name);
    for (int i = 0; i < noptions; i++) nomp_sha256_field(&c, options[i]);
    nomp_sha256_hex(&c, key);
    if (!nomp_jit_cache_load(key, "cubin", &cubin, &size)) {
      /* a damaged entry fails to load as a module: fall through and recompile */
      if (size > 0 && drv.ModuleLoadData(&cp->module, cubin) == CUDA_SUCCESS &&
          drv.ModuleGetFunction(&cp->function, cp->module, name) == CUDA_SUCCESS) {
        free(cubin);
        nomp_jit_cache_count(NOMP_CACHE_CUBIN_HIT);
        return 0;
      }
      if (cp->module) drv.ModuleUnload(cp->module), cp->module = NULL;
      free(cubin), cubin = NULL;
    }
  }

  nvrtcProgram prog;
  check_nvrtc(nvrtcCreateProgram(&prog, source, name, 0, NULL, NULL));
  nvrtcResult result = nvrtcCompileProgram(prog, noptions, options);
  if (result != NVRTC_SUCCESS) {
    size_t log_size = 0;
    nvrtcGetProgramLogSize(prog, &log_size);
    char *log = nomp_calloc(char, log_size + 1);
    nvrtcGetProgramLog(prog, log);
    int err = nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "CUDA build failure: %s: %s.", nvrtcGetErrorString(result), log);
    free(log);
    nvrtcDestroyProgram(&prog);
    return err;
  }
  check_nvrtc(nvrtcGetCUBINSize(prog, &size));
  cubin = nomp_calloc(char, size + 1);
  check_nvrtc(nvrtcGetCUBIN(prog, cubin));
  check_nvrtc(nvrtcDestroyProgram(&prog));
  if (cacheable && !nomp_jit_cache_store(key, "cubin", cubin, size)) nomp_jit_cache_count(NOMP_CACHE_CUBIN_MISS);
  CUresult r = drv.ModuleLoadData(&cp->module, cubin);
  free(cubin);
  check_driver(r);
  check_driver(drv.ModuleGetFunction(&cp->function, cp->module, name));
  return 0;
}

static int cuda_knl_build(nomp_backend_t *bnd, nomp_prog_t *prg, const char *source, const char *name) {
  cuda_state_t *st = (cuda_state_t *)bnd->bptr;
  char kind[32] = "", family[32] = "", num[32];
  if (strncmp(source, "//!nomp ", 8) || !desc_get(source, "kind", kind, sizeof(kind)) ||
      !desc_get(source, "family", family, sizeof(family)))
    return nomp_log(NOMP_LOOPY_CODEGEN_FAILURE, NOMP_ERROR, "Generated kernel \"%s\" has no descriptor line.", name);
  mark_const_args(prg, source);

  cuda_prog_t *cp = nomp_calloc(cuda_prog_t, 1);
  cp->a_y = cp->a_x = cp->a_z = cp->a_alpha = cp->a_beta = cp->a_n = cp->a_out = SLOT_NONE;
  cp->a_u = cp->a_g = cp->a_D = cp->a_w = cp->a_E = cp->a_r = SLOT_NONE;
  int err = 0;
  if (!strcmp(kind, "nvrtc")) {
    cp->family = FAM_NVRTC;
    err = build_nvrtc(st, cp, prg, source, name);
  } else if (!strcmp(kind, "native")) {
    cp->op = desc_get(source, "op", num, sizeof(num)) ? atoi(num) : 0;
    cp->dtype = desc_get(source, "dtype", num, sizeof(num)) ? atoi(num) : NOMPK_F64;
    if (!strcmp(family, "map")) {
      cp->family = FAM_MAP;
      (void)((err = desc_arg(prg, source, "y", 1, &cp->a_y)) || (err = desc_arg(prg, source, "x", 0, &cp->a_x)) ||
             (err = desc_arg(prg, source, "z", 0, &cp->a_z)) || (err = desc_arg(prg, source, "alpha", 0, &cp->a_alpha)) ||
             (err = desc_arg(prg, source, "beta", 0, &cp->a_beta)) ||
             (err = desc_count(prg, source, "n", &cp->a_n, &cp->n_literal)));
    } else if (!strcmp(family, "reduce")) {
      cp->family = FAM_REDUCE;
      cp->is_reduce = 1;
      (void)((err = desc_arg(prg, source, "x", 1, &cp->a_x)) || (err = desc_arg(prg, source, "y", 0, &cp->a_y)) ||
             (err = desc_arg(prg, source, "out", 1, &cp->a_out)) ||
             (err = desc_count(prg, source, "n", &cp->a_n, &cp->n_literal)));
    } else if (!strcmp(family, "ax") || !strcmp(family, "axdot") || !strcmp(family, "axxpaydot")) {
      cp->family = !strcmp(family, "ax") ? FAM_AX : !strcmp(family, "axdot") ? FAM_AXDOT : FAM_AXXPAYDOT;
      cp->is_reduce = cp->family != FAM_AX;
      cp->beta_dev = desc_get(source, "beta_dev", num, sizeof(num)) && num[0] == '1';
      cp->ax_n = desc_get(source, "n", num, sizeof(num)) ? atoi(num) : 0;
      if (!nompk_ax_supported(cp->ax_n))
        err = nomp_log(NOMP_LOOPY_CODEGEN_FAILURE, NOMP_ERROR, "No native Ax kernel for n = %d.", cp->ax_n);
      else
        (void)((err = desc_arg(prg, source, "u", 1, &cp->a_u)) || (err = desc_arg(prg, source, "g", 1, &cp->a_g)) ||
               (err = desc_arg(prg, source, "D", 1, &cp->a_D)) || (err = desc_arg(prg, source, "w", 1, &cp->a_w)) ||
               (err = desc_arg(prg, source, "E", 1, &cp->a_E)) ||
               (cp->family != FAM_AX && (err = desc_arg(prg, source, "out", 1, &cp->a_out))) ||
               (cp->family == FAM_AXXPAYDOT && ((err = desc_arg(prg, source, "r", 1, &cp->a_r)) ||
                                                (err = desc_arg(prg, source, "beta", 1, &cp->a_beta)))));
    } else {
      err = nomp_log(NOMP_LOOPY_CODEGEN_FAILURE, NOMP_ERROR, "Unknown native kernel family \"%s\".", family);
    }
  } else {
    err = nomp_log(NOMP_LOOPY_CODEGEN_FAILURE, NOMP_ERROR, "Unknown kernel kind \"%s\".", kind);
  }
  if (err) {
    free(cp);
    return err;
  }
  prg->bptr = cp;
  return 0;
}

/* ---- run ------------------------------------------------------------------------------------------------------ */
static long int_arg(const nomp_prog_t *prg, int idx, long literal) {
  if (idx < 0) return literal;
  const nomp_arg_t *a = &prg->args[idx];
  if (a->type == NOMP_UINT) return a->size == 8 ? (long)*(unsigned long *)a->ptr : (long)*(unsigned *)a->ptr;
  return a->size == 8 ? *(long *)a->ptr : (long)*(int *)a->ptr;
}

static void *ptr_arg(const nomp_prog_t *prg, int idx) { return idx < 0 ? NULL : prg->args[idx].ptr; }


/* Is the n x n matrix at device address D centro-antisymmetric, D[a][l] == -D[n-1-a][n-1-l] bit for bit -- what every
 * differentiation matrix on symmetric nodes is?  Then the Ax kernels may take their even-odd contractions
 * (NOMPK_AX_D_ANTISYMMETRIC, include/nompk.h).  Decided from the device image itself (the host array may have changed
 * since it was copied), once per (mapping, version): one 1 KB copy.  While a graph is captured the copy runs on a
 * stream of its own (nomp_b200_graph_begin has drained the compute stream); a D that a captured kernel may have
 * written -- its version is younger than the capture -- is not looked at, the general path is always right. */
static int ax_D_antisymmetric(cuda_state_t *st, const void *D, int n, unsigned long version, int have_mem) {
  if (getenv("NOMP_B200_NO_EVEN_ODD") || !have_mem || n > 16) return 0;
  if (st->ax_sym_D == D && st->ax_sym_n == n && st->ax_sym_version == version) return st->ax_sym;
  if (st->capturing && version > st->capture_version0) return 0;
  double h[16 * 16];
  if (!st->stream_aux && cudaStreamCreateWithFlags(&st->stream_aux, cudaStreamNonBlocking) != cudaSuccess) return 0;
  if (!st->capturing && cudaStreamSynchronize(st->stream) != cudaSuccess) return 0;
  if (cudaMemcpyAsync(h, D, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st->stream_aux) != cudaSuccess ||
      cudaStreamSynchronize(st->stream_aux) != cudaSuccess)
    return 0;
  int sym = 1;
  for (int a = 0; a < n && sym; a++)
    for (int l = 0; l < n; l++)
      if (h[a * n + l] != -h[(n - 1 - a) * n + (n - 1 - l)]) { /* (NaN fails, -0.0 == 0.0 passes: both fine) */
        sym = 0;
        break;
      }
  st->ax_sym_D = D, st->ax_sym_n = n, st->ax_sym_version = version, st->ax_sym = sym;
  return sym;
}

static int cuda_knl_run(nomp_backend_t *bnd, nomp_prog_t *prg) {
  cuda_state_t *st = (cuda_state_t *)bnd->bptr;
  cuda_prog_t *cp = (cuda_prog_t *)prg->bptr;
  /* with several ranks the host-visible value must be the all-reduced one: the native reductions all-reduce their own
   * result over NVLink peer memory (fused = 1), the others only fill the device slot and the finish all-reduces it */
  void *result_host = nomp_comm_size() > 1 ? NULL : st->pinned_dev;
  nompk_peers_t peers;
  const nompk_peers_t *px = NULL;
  st->fused_allreduce = 0;
  /* nomp_b200_device_reductions: the result goes to the device copy of the (mapped) reduction variable instead of the
   * backend's slot; the kernels are the same, only the address differs */
  st->red_result_arg = prg->reduction_dev ? prg->reduction_dev : st->red_result;
  /* a captured reduction cannot hand its result to the host: nomp_run would have to wait for a stream that is only
   * recording.  (Several ranks are fine: the number of the collective call is a counter in device memory.) */
  if (st->capturing && cp->is_reduce && !prg->reduction_dev)
    return capture_forbids("A reduce clause whose variable is not device-resident (nomp_b200_device_reductions)");
  if (prg->reduction_dev) result_host = NULL; /* nobody waits for this result: no store across PCIe in the kernel's tail */
  void *const err_host = (char *)st->pinned_dev + 16;
  /* a kernel that writes a mapping must not overtake an asynchronous copy that is still reading it */
  for (unsigned i = 0; i < prg->nargs; i++)
    if (prg->args[i].mem && !prg->args[i].is_const) nomp_check(nomp_cuda_before_write((nomp_mem_t *)prg->args[i].mem));

  switch (cp->family) {
  case FAM_MAP: {
    long n = int_arg(prg, cp->a_n, cp->n_literal);
    if (n < 0) n = 0;
    check_nompk(nompk_map((nompk_map_op_t)cp->op, (nompk_dtype_t)cp->dtype, (size_t)n, ptr_arg(prg, cp->a_y),
                          ptr_arg(prg, cp->a_x), ptr_arg(prg, cp->a_z), ptr_arg(prg, cp->a_alpha),
                          ptr_arg(prg, cp->a_beta), st->stream));
    return 0;
  }
  case FAM_REDUCE: {
    long n = int_arg(prg, cp->a_n, cp->n_literal);
    if (n < 0) n = 0;
    if (nomp_comm_size() > 1 && nomp_comm_peers(&peers, err_host)) {
      px = &peers, st->fused_allreduce = 1;
      if (!prg->reduction_dev) result_host = st->pinned_dev;
    }
    check_nompk(nompk_reduce_peers((nompk_red_op_t)cp->op, (nompk_dtype_t)cp->dtype, (size_t)n, ptr_arg(prg, cp->a_x),
                                   ptr_arg(prg, cp->a_y), st->red_result_arg, result_host, ++st->host_seq, st->red_ws, px,
                                   st->stream));
    return 0;
  }
  case FAM_AX:
  case FAM_AXDOT:
  case FAM_AXXPAYDOT: {
    long E = int_arg(prg, cp->a_E, 0);
    if (E < 0) E = 0;
    const void *D = ptr_arg(prg, cp->a_D);
    /* D is staged into __constant__ memory by a tiny D2D copy; skip it while the same device image is reused
     * (the mapping's version changes on every nomp_update(TO) and whenever a kernel may have written to it) */
    const nomp_mem_t *dm = (const nomp_mem_t *)prg->args[cp->a_D].mem;
    const unsigned long version = dm ? dm->version : 0;
    unsigned flags = 0;
    /* (while a graph is being captured the copy is recorded with the launch and the cache state is left alone) */
    if (dm && !st->capturing && st->ax_D == D && st->ax_n == cp->ax_n && st->ax_D_version == version) flags = NOMPK_AX_D_CACHED;
    if (ax_D_antisymmetric(st, D, cp->ax_n, version, dm != NULL)) flags |= NOMPK_AX_D_ANTISYMMETRIC;
    /* a rank without elements launches nothing: it joins through the stand-alone all-reduce kernel of the finish,
     * which speaks the same protocol on the same buffers */
    if (cp->family != FAM_AX && E > 0 && nomp_comm_size() > 1 && nomp_comm_peers(&peers, err_host)) {
      px = &peers, st->fused_allreduce = 1;
      if (!prg->reduction_dev) result_host = st->pinned_dev;
    }
    if (cp->family == FAM_AXXPAYDOT) {
      /* p <- r + beta p in front of the operator; beta is the caller's host scalar or a scalar in device memory */
      const double beta = cp->beta_dev ? 0.0 : *(const double *)prg->args[cp->a_beta].ptr;
      check_nompk(nompk_ax_xpay_dot_peers_f64(cp->ax_n, (size_t)E, (double *)ptr_arg(prg, cp->a_u),
                                              (const double *)ptr_arg(prg, cp->a_r), beta,
                                              cp->beta_dev ? (const double *)ptr_arg(prg, cp->a_beta) : NULL,
                                              (const double *)ptr_arg(prg, cp->a_g), (const double *)D,
                                              (double *)ptr_arg(prg, cp->a_w), (double *)st->red_result_arg,
                                              (double *)result_host, ++st->host_seq, st->red_ws, px, flags, st->stream));
    } else if (cp->family == FAM_AXDOT)
      check_nompk(nompk_ax_dot_peers_f64(cp->ax_n, (size_t)E, (const double *)ptr_arg(prg, cp->a_u),
                                         (const double *)ptr_arg(prg, cp->a_g), (const double *)D,
                                         (double *)ptr_arg(prg, cp->a_w), (double *)st->red_result_arg,
                                         (double *)result_host, ++st->host_seq, st->red_ws, px, flags,
                                         st->stream));
    else
      check_nompk(nompk_ax_f64(cp->ax_n, (size_t)E, (const double *)ptr_arg(prg, cp->a_u),
                               (const double *)ptr_arg(prg, cp->a_g), (const double *)D,
                               (double *)ptr_arg(prg, cp->a_w), flags, st->stream));
    if (!st->capturing) st->ax_D = D, st->ax_n = cp->ax_n, st->ax_D_version = version;
    else st->ax_D = NULL; /* a replay overwrites the staged copy at a time this cache cannot know */
    return 0;
  }
  case FAM_NVRTC: {
    st->nv_peers = NULL, st->nv_rank = 0, st->nv_world = 1, st->nv_cseq_dev = NULL, st->nv_err_host = err_host;
    if (cp->is_reduce && cp->has_peers && nomp_comm_size() > 1 && nomp_comm_peers(&peers, err_host)) {
      st->nv_peers = (void *)peers.peer_xchg, st->nv_rank = peers.rank, st->nv_world = peers.world;
      st->nv_cseq_dev = peers.seq_dev;
      st->fused_allreduce = 1;
      if (!prg->reduction_dev) result_host = st->pinned_dev;
    }
    st->red_result_host = result_host;
    if (cp->is_reduce) ++st->host_seq;
    void *vargs[NOMP_MAX_KERNEL_ARGS_SIZE + 10];
    for (int i = 0; i < cp->nparams; i++) {
      int s = cp->param_slot[i];
      if (s == SLOT_WS) vargs[i] = &st->red_ws;
      else if (s == SLOT_RESULT) vargs[i] = &st->red_result_arg;
      else if (s == SLOT_RESULT_HOST) vargs[i] = &st->red_result_host;
      else if (s == SLOT_SEQ) vargs[i] = &st->host_seq;
      else if (s == SLOT_PEERS) vargs[i] = &st->nv_peers;
      else if (s == SLOT_RANK) vargs[i] = &st->nv_rank;
      else if (s == SLOT_WORLD) vargs[i] = &st->nv_world;
      else if (s == SLOT_CSEQ) vargs[i] = &st->nv_cseq_dev;
      else if (s == SLOT_ERR) vargs[i] = &st->nv_err_host;
      else if (prg->args[s].type == NOMP_PTR) vargs[i] = &prg->args[s].ptr; /* device pointer by value */
      else vargs[i] = prg->args[s].ptr;                                       /* the caller's scalar */
    }
    const size_t *g = prg->global, *l = prg->local;
    if (g[0] == 0 || g[1] == 0 || g[2] == 0) return 0; /* empty iteration space */
    check_driver(drv.LaunchKernel(cp->function, (unsigned)g[0], (unsigned)g[1], (unsigned)g[2], (unsigned)l[0],
                                  (unsigned)l[1], (unsigned)l[2], 0, (CUstream)st->stream, vargs, NULL));
    st->nvrtc_launches++;
    return 0;
  }
  }
  return 0;
}

/* Backend half of the reduce finish (the type bookkeeping is in src/reduction.c): all-reduce the device scalar over the
 * ranks if there are several, then wait for the kernel to publish {value, sequence number} to the host. */
int nomp_cuda_reduction_finish(nomp_backend_t *bnd, nomp_prog_t *prg, int dtype, size_t size) {
  cuda_state_t *st = (cuda_state_t *)bnd->bptr;
  if (prg->reduction_dev) {
    /* the result stays on the device (nomp_b200_device_reductions): all-reduce it in place if the kernel has not done
     * so itself, and return without waiting -- the next kernel on the stream reads it from device memory */
    int published = 0;
    if (nomp_comm_size() > 1 && !st->fused_allreduce)
      nomp_check(nomp_comm_allreduce(prg->reduction_dev, dtype, (int)prg->reduction_op, NULL, 0, (char *)st->pinned_dev + 16,
                                     st->stream, &published));
    return 0;
  }
  if (nomp_comm_size() > 1 && !st->fused_allreduce) {
    /* all-reduce the device scalar in place; the NVLink one-shot kernel also publishes {value, seq} to the host,
     * the NCCL fallback needs an explicit 8-byte copy */
    int published = 0;
    nomp_check(nomp_comm_allreduce(st->red_result, dtype, (int)prg->reduction_op, st->pinned_dev, st->host_seq,
                                   (char *)st->pinned_dev + 16, st->stream, &published));
    if (!published) {
      check_runtime(cudaMemcpyAsync(st->pinned_host, st->red_result, 8, cudaMemcpyDeviceToHost, st->stream));
      check_runtime(cudaStreamSynchronize(st->stream));
      memcpy(prg->reduction_ptr, st->pinned_host, size);
      return 0;
    }
  }
  /* The reduce clause's result is valid on the host when nomp_run returns (reference
   * tests/nomp-api-500-impl.h:29-34 read it with no nomp_sync in between).  The kernel stores the value and then the
   * sequence number into mapped pinned memory; spinning on the sequence number is a few microseconds faster than
   * cudaStreamSynchronize.  The stream is polled now and then so that a failed launch cannot hang the caller. */
  volatile unsigned long long *seq = (volatile unsigned long long *)((char *)st->pinned_host + 8);
  for (unsigned long spins = 1; *seq != st->host_seq; spins++) {
    if ((spins & 0xfff) == 0) {
      cudaError_t q = cudaStreamQuery(st->stream);
      if (q == cudaSuccess) {
        if (*seq == st->host_seq) break;
        return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "reduction", "result was not published");
      }
      if (q != cudaErrorNotReady)
        return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "runtime", cudaGetErrorName(q));
    }
    __builtin_ia32_pause();
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  if (*(volatile unsigned long long *)((char *)st->pinned_host + 16) != 0) { /* set by the NVLink all-reduce kernel */
    *(volatile unsigned long long *)((char *)st->pinned_host + 16) = 0;
    return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "all-reduce",
                    "a rank did not join the reduction within 20 s");
  }
  memcpy(prg->reduction_ptr, st->pinned_host, size);
  return 0;
}

static int cuda_knl_free(nomp_prog_t *prg) {
  cuda_prog_t *cp = (cuda_prog_t *)prg->bptr;
  if (cp == NULL) return 0;
  if (cp->family == FAM_NVRTC && cp->module && drv.resolved) check_driver(drv.ModuleUnload(cp->module));
  free(cp);
  prg->bptr = NULL;
  return 0;
}

static int cuda_sync(nomp_backend_t *bnd) {
  cuda_state_t *st = (cuda_state_t *)bnd->bptr;
  if (st->capturing) return capture_forbids("nomp_sync");
  check_runtime(cudaStreamSynchronize(st->stream));
  if (st->async_used) {
    check_runtime(cudaStreamSynchronize(st->stream_h2d));
    check_runtime(cudaStreamSynchronize(st->stream_d2h));
  }
  /* a fused all-reduce whose result stays on the device has nobody waiting for it: a rank that never joined is
   * reported here */
  volatile unsigned long long *late = (volatile unsigned long long *)((char *)st->pinned_host + 16);
  if (*late != 0) {
    *late = 0;
    return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "all-reduce",
                    "a rank did not join the reduction within 20 s");
  }
  return 0;
}

/* Asynchronous variant of update() for the two copy directions (include/nomp-b200.h: nomp_b200_update_async).
 * H2D copies run on their own stream, D2H copies on another, so that with pinned host memory both PCIe directions
 * are busy at once while kernels run on the compute stream.  Ordering: a copy starts only after every kernel issued
 * so far (it may read or write the range); kernels issued later wait for the H2D copies issued so far.  The caller
 * must not touch the host range until nomp_sync(). */
int nomp_cuda_update_async(nomp_backend_t *bnd, nomp_mem_t *m, nomp_map_direction_t op, size_t start, size_t end,
                           size_t usize) {
  cuda_state_t *st = (cuda_state_t *)bnd->bptr;
  if (st->capturing) return capture_forbids("nomp_b200_update_async");
  if (!st->stream_h2d) {
    check_runtime(cudaStreamCreateWithFlags(&st->stream_h2d, cudaStreamNonBlocking));
    check_runtime(cudaStreamCreateWithFlags(&st->stream_d2h, cudaStreamNonBlocking));
    check_runtime(cudaEventCreateWithFlags(&st->ev_compute, cudaEventDisableTiming));
    check_runtime(cudaEventCreateWithFlags(&st->ev_h2d, cudaEventDisableTiming));
    check_runtime(cudaEventCreateWithFlags(&st->ev_d2h, cudaEventDisableTiming));
  }
  st->async_used = 1;
  if (m->transfers++ == 1) pin_host_range(m);
  cudaStream_t cs = op == NOMP_TO ? st->stream_h2d : st->stream_d2h;
  check_runtime(cudaEventRecord(st->ev_compute, st->stream));
  check_runtime(cudaStreamWaitEvent(cs, st->ev_compute, 0));
  char *dev = (char *)m->bptr + NOMP_MEM_OFFSET(start - m->idx0, usize);
  char *host = (char *)m->hptr + NOMP_MEM_OFFSET(start, usize);
  const size_t bytes = NOMP_MEM_BYTES(start, end, usize);
  if (op == NOMP_TO) {
    check_runtime(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, cs));
    check_runtime(cudaEventRecord(st->ev_h2d, cs));
    check_runtime(cudaStreamWaitEvent(st->stream, st->ev_h2d, 0));
    m->version = nomp_next_version();
  } else {
    /* later work on the compute stream that WRITES this mapping waits for the copy (nomp_cuda_before_write); work that
     * only reads it, or touches other mappings, overlaps with it */
    check_runtime(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, cs));
    check_runtime(cudaEventRecord(st->ev_d2h, cs));
    m->d2h_ticket = ++st->d2h_issued;
  }
  return 0;
}

static int cuda_finalize(nomp_backend_t *bnd) {
  cuda_state_t *st = (cuda_state_t *)bnd->bptr;
  if (st == NULL) return 0;
  nomp_comm_finalize();
  for (int i = 0; i < NOMP_MAX_GRAPHS; i++)
    if (st->graphs[i]) cudaGraphExecDestroy(st->graphs[i]);
  if (st->stream) {
    cudaStreamSynchronize(st->stream);
    cudaStreamDestroy(st->stream);
  }
  if (st->stream_h2d) {
    cudaStreamSynchronize(st->stream_h2d), cudaStreamSynchronize(st->stream_d2h);
    cudaStreamDestroy(st->stream_h2d), cudaStreamDestroy(st->stream_d2h);
    cudaEventDestroy(st->ev_compute), cudaEventDestroy(st->ev_h2d), cudaEventDestroy(st->ev_d2h);
  }
  if (st->stream_aux) cudaStreamDestroy(st->stream_aux);
  if (st->pinned_host) cudaFreeHost(st->pinned_host);
  if (g_state == st) g_state = NULL;
  free(st);
  bnd->bptr = NULL;
  return 0;
}

static int cuda_device_query(nomp_backend_t *bnd, cuda_state_t *st) {
  check_runtime(cudaGetDeviceProperties(&st->prop, st->device));
  nomp_py_dict_set_str(bnd->py_context, "device::name", st->prop.name);
  nomp_py_dict_set_str(bnd->py_context, "device::vendor", "NVIDIA");
  char arch[32];
  snprintf(arch, sizeof(arch), "sm_%d%d", st->prop.major, st->prop.minor);
  nomp_py_dict_set_str(bnd->py_context, "device::arch", arch);
  int driver = 0;
  check_runtime(cudaDriverGetVersion(&driver));
  nomp_py_dict_set_long(bnd->py_context, "device::driver", driver);
  nomp_py_dict_set_long(bnd->py_context, "device::max_threads_per_block", st->prop.maxThreadsPerBlock);
  nomp_py_dict_set_long(bnd->py_context, "device::multiprocessor_count", st->prop.multiProcessorCount);
  return 0;
}

int cuda_init(nomp_backend_t *bnd, int platform, int device) {
  (void)platform; /* validated by the core, meaningless for CUDA (reference unified-cuda-hip-impl.h:219) */
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess)
    return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "runtime", cudaGetErrorName(e));
  if (device < 0 || device >= count)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, ERR_STR_USER_DEVICE_IS_INVALID, device);
  check_runtime(cudaSetDevice(device));
  check_runtime(cudaFree(0));

  cuda_state_t *st = nomp_calloc(cuda_state_t, 1);
  st->device = device;
  int err = cuda_device_query(bnd, st);
  if (!err) {
    e = cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaHostAlloc(&st->pinned_host, 64, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&st->pinned_dev, st->pinned_host, 0);
    if (e != cudaSuccess)
      err = nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "runtime", cudaGetErrorName(e));
  }
  if (!err) err = nomp_comm_init(device);
  if (err) {
    if (st->stream) cudaStreamDestroy(st->stream);
    if (st->pinned_host) cudaFreeHost(st->pinned_host);
    free(st);
    return err;
  }
  memset(st->pinned_host, 0, 64);

  bnd->bptr = st;
  bnd->update = cuda_update;
  bnd->knl_build = cuda_knl_build;
  bnd->knl_run = cuda_knl_run;
  bnd->knl_free = cuda_knl_free;
  bnd->sync = cuda_sync;
  bnd->finalize = cuda_finalize;
  g_state = st;
  return 0;
}

/* ---- include/nomp-b200.h ---------------------------------------------------------------------------------------- */
NOMP_EXPORT void *nomp_b200_stream(void) { return g_state ? (void *)g_state->stream : NULL; }

NOMP_EXPORT unsigned long long nomp_b200_launch_count(void) {
  return nompk_launch_count() + (g_state ? g_state->nvrtc_launches : 0);
}

/* ---- CUDA graphs of nomp_run sequences (include/nomp-b200.h) ---------------------------------------------------------- */
/* Between graph_begin and graph_end the backend stream captures: every nomp_run records its launches
 * (cuLaunchKernel and the runtime launches of libnompk alike) instead of executing them.  Scalar arguments are frozen
 * into the graph with the values they have during capture; pointers are frozen as device addresses, so the mappings
 * must outlive the graph.  What varies between replays therefore lives in device memory: reduction results
 * (nomp_b200_device_reductions) and the scalars kernels read as alpha[0]. */
NOMP_EXPORT int nomp_b200_graph_begin(void) {
  cuda_state_t *st = g_state;
  if (!st) return nomp_log(NOMP_INITIALIZE_FAILURE, NOMP_ERROR, "libnomp is not initialized.");
  if (st->capturing) return capture_forbids("nomp_b200_graph_begin");
  check_runtime(cudaStreamSynchronize(st->stream));
  if (st->async_used) { /* nothing recorded in the graph may depend on work outside it */
    check_runtime(cudaStreamSynchronize(st->stream_h2d));
    check_runtime(cudaStreamSynchronize(st->stream_d2h));
    st->d2h_waited = st->d2h_issued;
  }
  st->capture_version0 = nomp_next_version();
  check_runtime(cudaStreamBeginCapture(st->stream, cudaStreamCaptureModeRelaxed));
  st->capturing = 1;
  return 0;
}

NOMP_EXPORT int nomp_b200_graph_end(int *graph) {
  cuda_state_t *st = g_state;
  if (!st || !st->capturing || !graph)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "nomp_b200_graph_end without nomp_b200_graph_begin.");
  st->capturing = 0;
  cudaGraph_t g = NULL;
  check_runtime(cudaStreamEndCapture(st->stream, &g));
  int slot = -1;
  for (int i = 0; i < NOMP_MAX_GRAPHS && slot < 0; i++)
    if (!st->graphs[i]) slot = i;
  cudaError_t e = slot < 0 ? cudaErrorMemoryAllocation : cudaGraphInstantiate(&st->graphs[slot], g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) {
    if (slot >= 0) st->graphs[slot] = NULL;
    return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, ERR_STR_CUDA_FAILURE, "graph", slot < 0 ? "too many graphs" : cudaGetErrorName(e));
  }
  *graph = slot;
  return 0;
}

NOMP_EXPORT int nomp_b200_graph_launch(int graph) {
  cuda_state_t *st = g_state;
  if (!st || graph < 0 || graph >= NOMP_MAX_GRAPHS || !st->graphs[graph] || st->capturing)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Graph id %d passed to nomp_b200_graph_launch is not valid.", graph);
  check_runtime(cudaGraphLaunch(st->graphs[graph], st->stream));
  st->ax_D = NULL; /* a captured Ax launch re-stages ITS derivative matrix: the cached one is no longer what is staged */
  return 0;
}

NOMP_EXPORT int nomp_b200_graph_free(int graph) {
  cuda_state_t *st = g_state;
  if (!st || graph < 0 || graph >= NOMP_MAX_GRAPHS || !st->graphs[graph])
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Graph id %d passed to nomp_b200_graph_free is not valid.", graph);
  check_runtime(cudaStreamSynchronize(st->stream));
  check_runtime(cudaGraphExecDestroy(st->graphs[graph]));
  st->graphs[graph] = NULL;
  return 0;
}
