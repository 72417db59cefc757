/* TEST DOUBLE -- not part of the product, never linked into it, never shipped.
 *
 * A stand-in for the CUDA runtime, NVRTC and the five driver entry points libnomp.so uses, so that the RUNTIME of
 * libnomp_b200 (src/nomp.c, src/loopy.c, src/reduction.c, backends/cuda.c: mappings, argument marshalling, the jit
 * bridge, launch-size expressions, the reduce finish, error paths) can be exercised by the CPU test tier.  It plays the
 * part pocl plays in the reference's CI ("testing without a GPU" = its OpenCL backend on a CPU OpenCL platform,
 * reference .github/workflows/ci.yml:60-77): the product is unchanged and unaware of it -- tests/test_hostdev_cpu.py
 * LD_PRELOADs this library into the reference's own nomp-api test programs (oracle/_ref/tests), nothing else does.
 * Without the preload, on a machine without a GPU, nomp_init() fails with NOMP_CUDA_FAILURE as it must.
 *
 *   "device" memory   = host memory (cudaMalloc -> posix_memalign, copies -> memmove); streams run at once.
 *   NVRTC             = tests/hostdev/compile_kernel.py: the generated CUDA source is compiled with g++ against the
 *                       cooperative emulator of tests/cuda_emulation.py (one coroutine per thread, real barriers,
 *                       shuffles, tickets); the "CUBIN" is the path of the resulting shared object.
 *   driver            = cuModuleLoadData -> dlopen, cuLaunchKernel -> the object's launcher (grid, block, void **params).
 *   libnompk launches = cudaLaunchKernel fails (there is no device code to run here); tests/hostdev/fake_nompk.c
 *                       interposes the C ABI of include/nompk.h instead.
 *   several ranks     = NOMP_HOSTDEV_SHARED=1: allocations are POSIX shared-memory objects and the cudaIpc* calls map
 *                       them into the peer processes; tests/hostdev/fake_nccl.c stands in for libnccl.so.2.
 */
#define _GNU_SOURCE
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <nvrtc.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#define EXPORT __attribute__((visibility("default")))

/* ---- runtime: devices ------------------------------------------------------------------------------------------------ */
static int n_devices(void) {
  const char *e = getenv("NOMP_HOSTDEV_DEVICES");
  return e ? atoi(e) : 1;
}

EXPORT cudaError_t cudaGetDeviceCount(int *count) {
  *count = n_devices();
  return cudaSuccess;
}
EXPORT cudaError_t cudaSetDevice(int device) { return device >= 0 && device < n_devices() ? cudaSuccess : cudaErrorInvalidDevice; }
EXPORT cudaError_t cudaGetDevice(int *device) {
  *device = 0;
  return cudaSuccess;
}
EXPORT cudaError_t cudaDriverGetVersion(int *v) {
  *v = 12090;
  return cudaSuccess;
}
EXPORT cudaError_t cudaGetDeviceProperties(struct cudaDeviceProp *prop, int device) {
  (void)device;
  memset(prop, 0, sizeof(*prop));
  snprintf(prop->name, sizeof(prop->name), "hostdev (CUDA runtime test double)");
  prop->major = 10, prop->minor = 0; /* the bridge is told sm_100, as on a B200 */
  prop->maxThreadsPerBlock = 1024;
  prop->multiProcessorCount = 148;
  prop->warpSize = 32;
  prop->sharedMemPerBlock = 48 * 1024;
  prop->totalGlobalMem = (size_t)1 << 34;
  return cudaSuccess;
}
EXPORT cudaError_t cudaDeviceGetAttribute(int *value, enum cudaDeviceAttr attr, int device) {
  (void)device;
  *value = attr == cudaDevAttrMultiProcessorCount ? 148 : 0;
  return cudaSuccess;
}
EXPORT cudaError_t cudaDeviceSynchronize(void);
EXPORT cudaError_t cudaGetLastError(void) { return cudaSuccess; }
EXPORT cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
EXPORT const char *cudaGetErrorName(cudaError_t e) {
  switch (e) {
  case cudaSuccess: return "cudaSuccess";
  case cudaErrorInvalidDevice: return "cudaErrorInvalidDevice";
  case cudaErrorMemoryAllocation: return "cudaErrorMemoryAllocation";
  case cudaErrorNotSupported: return "cudaErrorNotSupported";
  case cudaErrorInvalidValue: return "cudaErrorInvalidValue";
  case cudaErrorInvalidDeviceFunction: return "cudaErrorInvalidDeviceFunction";
  case cudaErrorStreamCaptureUnsupported: return "cudaErrorStreamCaptureUnsupported";
  case cudaErrorLaunchFailure: return "cudaErrorLaunchFailure";
  default: return "cudaErrorUnknown";
  }
}
EXPORT const char *cudaGetErrorString(cudaError_t e) { return cudaGetErrorName(e); }

/* ---- runtime: memory ------------------------------------------------------------------------------------------------- */
/* With NOMP_HOSTDEV_SHARED=1 (multi-rank tests: one process per rank) every "device" allocation is a POSIX shared-memory
 * object, so that cudaIpcGetMemHandle / cudaIpcOpenMemHandle can map it into a peer process -- host memory standing in
 * for NVLink peer memory.  The handle carries the object's name and size. */
typedef struct {
  void *ptr;
  size_t bytes;
  char name[48];
  int foreign; /* mapped through cudaIpcOpenMemHandle */
} shared_alloc_t;
static shared_alloc_t shared_allocs[4096];
static int n_shared_allocs = 0;

static int shared_mode(void) {
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("NOMP_HOSTDEV_SHARED");
    mode = e && e[0] == '1';
  }
  return mode;
}

static shared_alloc_t *find_shared(const void *p) {
  for (int i = 0; i < n_shared_allocs; i++)
    if (shared_allocs[i].ptr == p) return &shared_allocs[i];
  return NULL;
}

static void *map_shared(const char *name, size_t bytes, int create) {
  const int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
  if (fd < 0) return NULL;
  if (create && ftruncate(fd, (off_t)bytes) != 0) {
    close(fd);
    shm_unlink(name);
    return NULL;
  }
  void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  return p == MAP_FAILED ? NULL : p;
}

EXPORT cudaError_t cudaMalloc(void **p, size_t bytes) {
  if (bytes == 0) bytes = 1;
  if (shared_mode()) {
    static unsigned serial = 0;
    if (n_shared_allocs == 4096) return cudaErrorMemoryAllocation;
    shared_alloc_t *a = &shared_allocs[n_shared_allocs];
    snprintf(a->name, sizeof(a->name), "/nomp-hostdev-%ld-%u", (long)getpid(), serial++);
    a->bytes = (bytes + 4095) & ~(size_t)4095;
    a->ptr = map_shared(a->name, a->bytes, 1);
    a->foreign = 0;
    if (!a->ptr) return cudaErrorMemoryAllocation;
    n_shared_allocs++;
    *p = a->ptr;
  } else if (posix_memalign(p, 256, bytes)) {
    return cudaErrorMemoryAllocation;
  }
  memset(*p, 0xA5, bytes); /* fresh device memory is not zero */
  return cudaSuccess;
}
EXPORT cudaError_t cudaFree(void *p) {
  shared_alloc_t *a = p ? find_shared(p) : NULL;
  if (a) {
    munmap(a->ptr, a->bytes);
    if (!a->foreign) shm_unlink(a->name);
    *a = shared_allocs[--n_shared_allocs];
  } else if (!shared_mode()) {
    free(p);
  }
  return cudaSuccess;
}
__attribute__((destructor)) static void unlink_shared(void) { /* a rank that exits without freeing must not leave objects behind */
  for (int i = 0; i < n_shared_allocs; i++)
    if (!shared_allocs[i].foreign) shm_unlink(shared_allocs[i].name);
}
EXPORT cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned flags) {
  (void)flags;
  return posix_memalign(p, 256, bytes ? bytes : 1) ? cudaErrorMemoryAllocation : cudaSuccess;
}
EXPORT cudaError_t cudaFreeHost(void *p) {
  free(p);
  return cudaSuccess;
}
EXPORT cudaError_t cudaHostGetDevicePointer(void **dev, void *host, unsigned flags) {
  (void)flags;
  *dev = host;
  return cudaSuccess;
}
EXPORT cudaError_t cudaHostRegister(void *p, size_t bytes, unsigned flags) {
  (void)p, (void)bytes, (void)flags;
  return cudaSuccess;
}
EXPORT cudaError_t cudaHostUnregister(void *p) {
  (void)p;
  return cudaSuccess;
}
EXPORT cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, enum cudaMemcpyKind kind) {
  (void)kind;
  memmove(dst, src, bytes);
  return cudaSuccess;
}
EXPORT cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, enum cudaMemcpyKind kind, cudaStream_t s) {
  (void)s;
  return cudaMemcpy(dst, src, bytes, kind);
}
EXPORT cudaError_t cudaMemset(void *p, int value, size_t bytes) {
  memset(p, value, bytes);
  return cudaSuccess;
}
EXPORT cudaError_t cudaMemsetAsync(void *p, int value, size_t bytes, cudaStream_t s) {
  (void)s;
  return cudaMemset(p, value, bytes);
}
EXPORT cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
  shared_alloc_t *a = find_shared(p);
  if (!a || a->foreign) return cudaErrorNotSupported;
  memset(h, 0, sizeof(*h));
  memcpy(h->reserved, &a->bytes, sizeof(size_t));
  memcpy(h->reserved + sizeof(size_t), a->name, sizeof(a->name));
  return cudaSuccess;
}
EXPORT cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags) {
  (void)flags;
  if (!shared_mode() || n_shared_allocs == 4096) return cudaErrorNotSupported;
  shared_alloc_t *a = &shared_allocs[n_shared_allocs];
  memcpy(&a->bytes, h.reserved, sizeof(size_t));
  memcpy(a->name, h.reserved + sizeof(size_t), sizeof(a->name));
  a->name[sizeof(a->name) - 1] = '\0';
  a->ptr = map_shared(a->name, a->bytes, 0);
  a->foreign = 1;
  if (!a->ptr) return cudaErrorInvalidValue;
  n_shared_allocs++;
  *p = a->ptr;
  return cudaSuccess;
}
EXPORT cudaError_t cudaIpcCloseMemHandle(void *p) {
  shared_alloc_t *a = find_shared(p);
  if (!a || !a->foreign) return cudaErrorInvalidValue;
  return cudaFree(p);
}

/* ---- stream capture and graphs: launches are recorded (arguments copied, as CUDA copies kernel parameters) and replayed ---- */
typedef struct {
  int (*fn)(void *blob);
  void *blob;
} recorded_t;
typedef struct {
  recorded_t *items;
  int n, cap;
} recording_t;
static recording_t *capturing = NULL;

/* used by fake_nompk.c and by cuLaunchKernel below: 1 = recorded (do not execute now) */
__attribute__((visibility("default"))) int nomp_hostdev_record(int (*fn)(void *), const void *blob, size_t bytes) {
  if (!capturing) return 0;
  if (capturing->n == capturing->cap) {
    capturing->cap = capturing->cap ? 2 * capturing->cap : 16;
    capturing->items = realloc(capturing->items, sizeof(recorded_t) * (size_t)capturing->cap);
  }
  void *copy = malloc(bytes);
  memcpy(copy, blob, bytes);
  capturing->items[capturing->n].fn = fn, capturing->items[capturing->n].blob = copy;
  capturing->n++;
  return 1;
}

EXPORT cudaError_t cudaStreamBeginCapture(cudaStream_t s, enum cudaStreamCaptureMode mode) {
  (void)s, (void)mode;
  if (capturing) return cudaErrorInvalidValue;
  capturing = calloc(1, sizeof(*capturing));
  return cudaSuccess;
}
EXPORT cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *graph) {
  (void)s;
  if (!capturing) return cudaErrorInvalidValue;
  *graph = (cudaGraph_t)capturing;
  capturing = NULL;
  return cudaSuccess;
}
EXPORT cudaError_t cudaGraphInstantiate(cudaGraphExec_t *exec, cudaGraph_t graph, unsigned long long flags) {
  (void)flags;
  recording_t *src = (recording_t *)graph, *dst = calloc(1, sizeof(*dst));
  dst->n = dst->cap = src->n;
  dst->items = malloc(sizeof(recorded_t) * (size_t)(src->n ? src->n : 1));
  memcpy(dst->items, src->items, sizeof(recorded_t) * (size_t)src->n); /* the blobs are shared: freed with the exec */
  *exec = (cudaGraphExec_t)dst;
  return cudaSuccess;
}
EXPORT cudaError_t cudaGraphDestroy(cudaGraph_t graph) {
  recording_t *r = (recording_t *)graph;
  free(r->items);
  free(r);
  return cudaSuccess;
}
EXPORT cudaError_t cudaGraphExecDestroy(cudaGraphExec_t exec) {
  recording_t *r = (recording_t *)exec;
  for (int i = 0; i < r->n; i++) free(r->items[i].blob);
  free(r->items);
  free(r);
  return cudaSuccess;
}
EXPORT cudaError_t cudaGraphLaunch(cudaGraphExec_t exec, cudaStream_t s) {
  (void)s;
  recording_t *r = (recording_t *)exec;
  if (capturing) return cudaErrorInvalidValue;
  for (int i = 0; i < r->n; i++)
    if (r->items[i].fn(r->items[i].blob)) return cudaErrorLaunchFailure;
  return cudaSuccess;
}
/* everything that waits for, or is ordered outside, a capturing stream is an error in CUDA: make it one here too */
#define REFUSE_WHILE_CAPTURING()                                                                                      \
  do {                                                                                                                 \
    if (capturing) return cudaErrorStreamCaptureUnsupported;                                                          \
  } while (0)

/* ---- runtime: streams and events (everything has completed by the time a call returns) ---------------------------------- */
EXPORT cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags) {
  (void)flags;
  *s = (cudaStream_t)calloc(1, 8);
  return cudaSuccess;
}
EXPORT cudaError_t cudaStreamDestroy(cudaStream_t s) {
  free(s);
  return cudaSuccess;
}
EXPORT cudaError_t cudaStreamSynchronize(cudaStream_t s) {
  (void)s;
  REFUSE_WHILE_CAPTURING();
  return cudaSuccess;
}
EXPORT cudaError_t cudaStreamQuery(cudaStream_t s) {
  (void)s;
  REFUSE_WHILE_CAPTURING();
  return cudaSuccess;
}
EXPORT cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags) {
  (void)s, (void)e, (void)flags;
  return cudaSuccess;
}
EXPORT cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags) {
  (void)flags;
  *e = (cudaEvent_t)calloc(1, 8);
  return cudaSuccess;
}
EXPORT cudaError_t cudaEventDestroy(cudaEvent_t e) {
  free(e);
  return cudaSuccess;
}
EXPORT cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  (void)e, (void)s;
  return cudaSuccess;
}

/* ---- runtime: launches of compiled device code cannot run here -------------------------------------------------------------- */
EXPORT cudaError_t cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t smem, cudaStream_t s) {
  (void)func, (void)grid, (void)block, (void)args, (void)smem, (void)s;
  fprintf(stderr, "hostdev: cudaLaunchKernel of compiled device code (libnompk) -- there is no GPU behind this test double\n");
  return cudaErrorInvalidDeviceFunction;
}
EXPORT cudaError_t cudaMemcpyToSymbolAsync(const void *symbol, const void *src, size_t bytes, size_t offset,
                                          enum cudaMemcpyKind kind, cudaStream_t s) {
  (void)symbol, (void)src, (void)bytes, (void)offset, (void)kind, (void)s;
  return cudaErrorInvalidDeviceFunction;
}

EXPORT cudaError_t cudaDeviceSynchronize(void) {
  REFUSE_WHILE_CAPTURING();
  return cudaSuccess;
}

/* ---- NVRTC -------------------------------------------------------------------------------------------------------------------- */
typedef struct {
  char *src, *name, *log, *image;
} fake_prog_t;

EXPORT nvrtcResult nvrtcVersion(int *major, int *minor) {
  *major = 12, *minor = 9;
  return NVRTC_SUCCESS;
}
EXPORT const char *nvrtcGetErrorString(nvrtcResult r) { return r == NVRTC_SUCCESS ? "NVRTC_SUCCESS" : "NVRTC_ERROR_COMPILATION"; }
EXPORT nvrtcResult nvrtcCreateProgram(nvrtcProgram *prog, const char *src, const char *name, int nh, const char *const *h,
                                      const char *const *in) {
  (void)nh, (void)h, (void)in;
  fake_prog_t *p = calloc(1, sizeof(*p));
  p->src = strdup(src), p->name = strdup(name ? name : "kernel");
  *prog = (nvrtcProgram)p;
  return NVRTC_SUCCESS;
}
EXPORT nvrtcResult nvrtcDestroyProgram(nvrtcProgram *prog) {
  fake_prog_t *p = (fake_prog_t *)*prog;
  if (p) free(p->src), free(p->name), free(p->log), free(p->image), free(p);
  *prog = NULL;
  return NVRTC_SUCCESS;
}

static char *read_file(const char *path) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return strdup("");
  char *buf = calloc(1, 1 << 16);
  size_t n = fread(buf, 1, (1 << 16) - 1, fp);
  buf[n] = '\0';
  fclose(fp);
  return buf;
}

EXPORT nvrtcResult nvrtcCompileProgram(nvrtcProgram prog, int nopt, const char *const *opt) {
  (void)nopt, (void)opt;
  fake_prog_t *p = (fake_prog_t *)prog;
  const char *dir = getenv("NOMP_HOSTDEV_DIR"), *py = getenv("NOMP_HOSTDEV_PYTHON"), *helper = getenv("NOMP_HOSTDEV_COMPILER");
  if (!dir || !py || !helper) {
    p->log = strdup("hostdev: NOMP_HOSTDEV_DIR / NOMP_HOSTDEV_PYTHON / NOMP_HOSTDEV_COMPILER are not set");
    return NVRTC_ERROR_COMPILATION;
  }
  static unsigned serial = 0;
  char base[1024], cmd[8 * 1024];
  snprintf(base, sizeof(base), "%s/k%ld_%u", dir, (long)getpid(), serial++);
  snprintf(cmd, sizeof(cmd), "%s.cu", base);
  FILE *fp = fopen(cmd, "w");
  if (!fp) {
    p->log = strdup("hostdev: cannot write the kernel source");
    return NVRTC_ERROR_COMPILATION;
  }
  fputs(p->src, fp);
  fclose(fp);
  /* the embedded interpreter's environment must not leak into the helper */
  snprintf(cmd, sizeof(cmd), "env -u PYTHONHOME -u PYTHONPATH '%s' '%s' '%s.cu' '%s.so' > '%s.log' 2>&1", py, helper, base, base, base);
  const int rc = system(cmd);
  snprintf(cmd, sizeof(cmd), "%s.log", base);
  p->log = read_file(cmd);
  if (rc != 0) return NVRTC_ERROR_COMPILATION;
  p->image = malloc(strlen(base) + 16);
  sprintf(p->image, "HOSTDEV:%s.so", base);
  return NVRTC_SUCCESS;
}
EXPORT nvrtcResult nvrtcGetProgramLogSize(nvrtcProgram prog, size_t *size) {
  fake_prog_t *p = (fake_prog_t *)prog;
  *size = (p->log ? strlen(p->log) : 0) + 1;
  return NVRTC_SUCCESS;
}
EXPORT nvrtcResult nvrtcGetProgramLog(nvrtcProgram prog, char *log) {
  fake_prog_t *p = (fake_prog_t *)prog;
  strcpy(log, p->log ? p->log : "");
  return NVRTC_SUCCESS;
}
EXPORT nvrtcResult nvrtcGetCUBINSize(nvrtcProgram prog, size_t *size) {
  fake_prog_t *p = (fake_prog_t *)prog;
  if (!p->image) return NVRTC_ERROR_INVALID_PROGRAM;
  *size = strlen(p->image) + 1;
  return NVRTC_SUCCESS;
}
EXPORT nvrtcResult nvrtcGetCUBIN(nvrtcProgram prog, char *cubin) {
  fake_prog_t *p = (fake_prog_t *)prog;
  if (!p->image) return NVRTC_ERROR_INVALID_PROGRAM;
  strcpy(cubin, p->image);
  return NVRTC_SUCCESS;
}

/* ---- driver entry points ----------------------------------------------------------------------------------------------------------- */
typedef int (*hostdev_launch_t)(unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void **);

static CUresult fake_cuModuleLoadData(CUmodule *module, const void *image) {
  if (strncmp((const char *)image, "HOSTDEV:", 8)) return CUDA_ERROR_INVALID_IMAGE;
  void *h = dlopen((const char *)image + 8, RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    fprintf(stderr, "hostdev: %s\n", dlerror());
    return CUDA_ERROR_INVALID_IMAGE;
  }
  *module = (CUmodule)h;
  return CUDA_SUCCESS;
}
static CUresult fake_cuModuleGetFunction(CUfunction *f, CUmodule module, const char *name) {
  const char *const *have = (const char *const *)dlsym((void *)module, "nomp_hostdev_kernel_name");
  void *launch = dlsym((void *)module, "nomp_emu_launch");
  if (!have || !launch || strcmp(*have, name)) return CUDA_ERROR_NOT_FOUND;
  *f = (CUfunction)launch;
  return CUDA_SUCCESS;
}
static CUresult fake_cuModuleUnload(CUmodule module) { return dlclose((void *)module) ? CUDA_ERROR_INVALID_HANDLE : CUDA_SUCCESS; }
/* a launch as a graph node: the function, its dimensions and a COPY of every parameter (sizes from the compiled object) */
typedef struct {
  hostdev_launch_t launch;
  unsigned dims[6];
  int nparams;
  size_t offset[80];
  unsigned char data[80 * 16];
} launch_node_t;

static int replay_launch(void *blob) {
  launch_node_t *n = (launch_node_t *)blob;
  void *params[80];
  for (int i = 0; i < n->nparams; i++) params[i] = n->data + n->offset[i];
  return n->launch(n->dims[0], n->dims[1], n->dims[2], n->dims[3], n->dims[4], n->dims[5], params);
}

static CUresult fake_cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                    unsigned smem, CUstream s, void **params, void **extra) {
  (void)smem, (void)s, (void)extra;
  if (bx * by * bz > 1024 || bx * by * bz == 0) return CUDA_ERROR_INVALID_VALUE;
  if (capturing) {
    Dl_info info;
    if (!dladdr((void *)f, &info)) return CUDA_ERROR_INVALID_HANDLE;
    void *lib = dlopen(info.dli_fname, RTLD_NOW | RTLD_NOLOAD);
    const int *count = lib ? (const int *)dlsym(lib, "nomp_hostdev_param_count") : NULL;
    const size_t *sizes = lib ? (const size_t *)dlsym(lib, "nomp_hostdev_param_sizes") : NULL;
    if (!count || !sizes || *count > 80) return CUDA_ERROR_INVALID_HANDLE;
    launch_node_t node;
    memset(&node, 0, sizeof(node));
    node.launch = (hostdev_launch_t)f, node.nparams = *count;
    const unsigned dims[6] = {gx, gy, gz, bx, by, bz};
    memcpy(node.dims, dims, sizeof(dims));
    size_t off = 0;
    for (int i = 0; i < *count; i++) {
      if (sizes[i] > 16) return CUDA_ERROR_INVALID_VALUE;
      node.offset[i] = off;
      memcpy(node.data + off, params[i], sizes[i]);
      off += 16;
    }
    if (lib) dlclose(lib);
    nomp_hostdev_record(replay_launch, &node, sizeof(node));
    return CUDA_SUCCESS;
  }
  return ((hostdev_launch_t)f)(gx, gy, gz, bx, by, bz, params) ? CUDA_ERROR_LAUNCH_FAILED : CUDA_SUCCESS;
}
static CUresult fake_cuGetErrorName(CUresult r, const char **name) {
  *name = r == CUDA_SUCCESS             ? "CUDA_SUCCESS"
          : r == CUDA_ERROR_INVALID_IMAGE ? "CUDA_ERROR_INVALID_IMAGE"
          : r == CUDA_ERROR_NOT_FOUND     ? "CUDA_ERROR_NOT_FOUND"
          : r == CUDA_ERROR_LAUNCH_FAILED ? "CUDA_ERROR_LAUNCH_FAILED"
          : r == CUDA_ERROR_INVALID_VALUE ? "CUDA_ERROR_INVALID_VALUE"
                                          : "CUDA_ERROR_UNKNOWN";
  return CUDA_SUCCESS;
}

EXPORT cudaError_t cudaGetDriverEntryPoint(const char *symbol, void **fn, unsigned long long flags,
                                          enum cudaDriverEntryPointQueryResult *status) {
  (void)flags;
  *fn = !strcmp(symbol, "cuModuleLoadData")      ? (void *)fake_cuModuleLoadData
        : !strcmp(symbol, "cuModuleGetFunction") ? (void *)fake_cuModuleGetFunction
        : !strcmp(symbol, "cuModuleUnload")      ? (void *)fake_cuModuleUnload
        : !strcmp(symbol, "cuLaunchKernel")      ? (void *)fake_cuLaunchKernel
        : !strcmp(symbol, "cuGetErrorName")      ? (void *)fake_cuGetErrorName
                                                 : NULL;
  if (status) *status = *fn ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
  return cudaSuccess;
}
