/* NCCL communicator for multi-GPU runs (one process per GPU).
 *
 * The reference has no collective of any kind.  The only exchange on the north-star path is the
 * scalar of a reduce clause, e.g. the dot products of a CG iteration: each rank reduces its slice in one kernel
 * and this file all-reduces the 4/8-byte result in place on the backend stream (NVLink 5 / NVSwitch).
 *
 * Bootstrap without MPI: NOMP_COMM_SIZE, NOMP_COMM_RANK and NOMP_COMM_ID_FILE.  Rank 0 creates the ncclUniqueId and
 * publishes it by writing <file>.tmp and renaming it to <file>; the other ranks poll for <file>.  The caller must
 * pick a path that is fresh for every job (bench.py derives it from MASTER_PORT and a broadcast token).
 * NCCL is dlopen()ed on first use, so single-GPU programs and the nomp-api tests never load it.
 *
 * Fast path: an 8-byte allreduce through NCCL costs a collective launch plus protocol latency (tens of microseconds),
 * which is the whole latency budget of a CG step at 8 GPUs.  When every rank can map its peers' memory
 * (CUDA IPC over NVLink / NVSwitch) the communicator also sets up the exchange buffers of libnompk's one-shot
 * all-reduce kernel (include/nompk.h: nompk_allreduce_scalar) and uses it instead; NCCL stays as the fallback and is
 * selected with NOMP_COMM_ALLREDUCE=nccl.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <time.h>
#include <unistd.h>

#include "nomp-aux.h"
#include "nomp-impl.h"
#include "nompk.h"

static struct {
  void *handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char *(*GetErrorString)(ncclResult_t);
} nccl;

static ncclComm_t comm = NULL;
static int comm_rank = 0, comm_size = 1;
static char comm_path[PATH_MAX + 1]; /* NOMP_COMM_ID_FILE of this communicator */

/* NVLink one-shot all-reduce state */
static struct {
  int enabled;
  void *mine;          /* this rank's exchange buffer (cudaMalloc) */
  void *peers[64];     /* every rank's buffer as mapped into this process */
  void **table_dev;    /* device copy of peers[] */
  unsigned long long *seq_dev; /* number of the last collective call, in DEVICE memory: the kernels count (nompk.h,
                                  nompk_peers_t), so a captured launch can be replayed and ranks need no host agreement */
} p2p;

#define check_nccl(call)                                                                                         \
  do {                                                                                                           \
    ncclResult_t r_ = (call);                                                                                    \
    if (r_ != ncclSuccess)                                                                                       \
      return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "NCCL failure: %s.", nccl.GetErrorString(r_));              \
  } while (0)

static int load_nccl(void) {
  if (nccl.handle) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (unsigned i = 0; i < 2 && !nccl.handle; i++) nccl.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!nccl.handle) return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "NCCL failure: cannot load libnccl.so.2 (%s).", dlerror());
#define RESOLVE(field, sym)                                                                                      \
  if (!(*(void **)&nccl.field = dlsym(nccl.handle, sym)))                                                        \
    return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "NCCL failure: symbol %s not found.", sym);
  RESOLVE(GetUniqueId, "ncclGetUniqueId")
  RESOLVE(CommInitRank, "ncclCommInitRank")
  RESOLVE(AllReduce, "ncclAllReduce")
  RESOLVE(CommDestroy, "ncclCommDestroy")
  RESOLVE(GetErrorString, "ncclGetErrorString")
#undef RESOLVE
  return 0;
}

static int env_int(const char *name, int fallback) {
  const char *v = getenv(name);
  if (!v) return fallback;
  return nomp_str_toui(v, NOMP_MAX_BUFFER_SIZE);
}

static void p2p_setup(const char *path, int rank, int size);

int nomp_comm_rank(void) { return comm_rank; }
int nomp_comm_size(void) { return comm_size; }
NOMP_EXPORT int nomp_b200_comm_rank(void) { return comm_rank; }
NOMP_EXPORT int nomp_b200_comm_size(void) { return comm_size; }

/* Rank 0 publishes `bytes` bytes through the file `path` (write <path>.tmp, then rename: readers never see a partial
 * file); every other rank polls for the file (up to 120 s) and reads them.  Exported so that the rendezvous can be
 * tested without GPUs (tests/test_multirank_cpu.py). */
NOMP_EXPORT int nomp_b200_exchange_blob(const char *path, int rank, void *blob, size_t bytes) {
  if (rank == 0) {
    char *tmp = nomp_str_cat(2, PATH_MAX, path, ".tmp");
    FILE *f = fopen(tmp, "wb");
    int ok = f && fwrite(blob, bytes, 1, f) == 1;
    if (f) ok = (fclose(f) == 0) && ok;
    ok = ok && rename(tmp, path) == 0;
    free(tmp);
    if (!ok) return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Cannot publish the NCCL id through \"%s\".", path);
    return 0;
  }
  for (int tries = 0; tries < 12000; tries++) {
    FILE *f = fopen(path, "rb");
    if (f) {
      int got = fread(blob, bytes, 1, f) == 1;
      fclose(f);
      if (got) return 0;
    }
    struct timespec ts = {0, 10 * 1000 * 1000};
    nanosleep(&ts, NULL);
  }
  return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Timed out waiting for the NCCL id in \"%s\".", path);
}

int nomp_comm_init(int device) {
  (void)device;
  comm_rank = 0, comm_size = 1;
  const int size = env_int("NOMP_COMM_SIZE", 1);
  if (size == 1) return 0;
  const int rank = env_int("NOMP_COMM_RANK", -1);
  const char *path = getenv("NOMP_COMM_ID_FILE");
  if (size < 1 || rank < 0 || rank >= size || !path || !path[0])
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR,
                    "NOMP_COMM_SIZE > 1 needs a valid NOMP_COMM_RANK and NOMP_COMM_ID_FILE.");
  nomp_check(load_nccl());

  ncclUniqueId id;
  if (rank == 0) check_nccl(nccl.GetUniqueId(&id));
  nomp_check(nomp_b200_exchange_blob(path, rank, &id, sizeof(id)));
  check_nccl(nccl.CommInitRank(&comm, size, id, rank));
  comm_rank = rank, comm_size = size;
  snprintf(comm_path, sizeof(comm_path), "%s", path);

  /* optional NVLink one-shot path: enabled only if EVERY rank managed to map every peer (min over ranks via NCCL) */
  p2p_setup(path, rank, size);
  int *flag_dev = NULL;
  int flag = p2p.enabled;
  if (cudaMalloc((void **)&flag_dev, sizeof(int)) == cudaSuccess) {
    cudaMemcpy(flag_dev, &flag, sizeof(int), cudaMemcpyHostToDevice);
    ncclResult_t r = nccl.AllReduce(flag_dev, flag_dev, 1, ncclInt32, ncclMin, comm, 0);
    cudaDeviceSynchronize();
    if (r == ncclSuccess) cudaMemcpy(&flag, flag_dev, sizeof(int), cudaMemcpyDeviceToHost);
    else flag = 0;
    cudaFree(flag_dev);
  } else {
    flag = 0;
  }
  p2p.enabled = flag;
  /* every rank has read what it needed (the all-reduce above is also a barrier): remove the rendezvous files so that a
   * later job, or a re-init in this process, that reuses the path cannot pick up a stale id or stale IPC handles */
  char name[PATH_MAX + 64];
  snprintf(name, sizeof(name), "%s.ipc.%d", path, rank);
  unlink(name);
  if (rank == 0) unlink(path);
  nomp_info("multi-GPU reductions use %s", p2p.enabled ? "the NVLink one-shot all-reduce kernel" : "ncclAllReduce");
  return 0;
}

/* Exchange CUDA IPC handles of the per-rank exchange buffers through files next to the id file and map the peers'.
 * Any failure simply leaves the NCCL path in place. */
static void p2p_setup(const char *path, int rank, int size) {
  memset(&p2p, 0, sizeof(p2p));
  const char *mode = getenv("NOMP_COMM_ALLREDUCE");
  if ((mode && !strcmp(mode, "nccl")) || size > 64) return;
  const size_t bytes = nompk_allreduce_xchg_bytes(size);
  if (cudaMalloc(&p2p.mine, bytes) != cudaSuccess) return;
  cudaMemset(p2p.mine, 0, bytes);
  cudaDeviceSynchronize();
  cudaIpcMemHandle_t handle;
  int ok = cudaIpcGetMemHandle(&handle, p2p.mine) == cudaSuccess;
  /* all-gather of (ok, handle): rank r publishes "<path>.ipc.<r>", everybody reads everybody */
  struct { int ok; cudaIpcMemHandle_t h; } rec, all[64];
  rec.ok = ok, rec.h = handle;
  char name[PATH_MAX];
  snprintf(name, sizeof(name), "%s.ipc.%d", path, rank);
  if (nomp_b200_exchange_blob(name, 0, &rec, sizeof(rec)) > 0) ok = 0;
  for (int r = 0; r < size; r++) {
    snprintf(name, sizeof(name), "%s.ipc.%d", path, r);
    if (r == rank) all[r] = rec;
    else if (nomp_b200_exchange_blob(name, 1, &all[r], sizeof(all[r])) > 0) all[r].ok = 0;
    ok = ok && all[r].ok;
  }
  for (int r = 0; ok && r < size; r++) {
    if (r == rank) p2p.peers[r] = p2p.mine;
    else if (cudaIpcOpenMemHandle(&p2p.peers[r], all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
  }
  if (ok && cudaMalloc((void **)&p2p.table_dev, sizeof(void *) * (size_t)size) == cudaSuccess &&
      cudaMemcpy(p2p.table_dev, p2p.peers, sizeof(void *) * (size_t)size, cudaMemcpyHostToDevice) == cudaSuccess &&
      cudaMalloc((void **)&p2p.seq_dev, sizeof(unsigned long long)) == cudaSuccess &&
      cudaMemset(p2p.seq_dev, 0, sizeof(unsigned long long)) == cudaSuccess)
    p2p.enabled = 1;
  cudaDeviceSynchronize();
  cudaGetLastError(); /* clear sticky-free errors of the probing above */
  /* every rank must agree: a rank that failed keeps enabled = 0 and the others would wait for it forever, so
   * nomp_comm_init() takes the minimum over the ranks with one NCCL all-reduce (which is also the barrier after which
   * the rendezvous files are removed) */
}

/* Every rank publishes its record as "<id file>.<tag>.<rank>" and reads everybody else's.  The files are removed by
 * their writers after a barrier, so a tag can only be used once per communicator. */
int nomp_comm_allgather(const char *tag, const void *mine, void *all, size_t bytes) {
  if (comm_size == 1) {
    memcpy(all, mine, bytes);
    return 0;
  }
  char name[PATH_MAX + 128];
  void *copy = malloc(bytes);
  memcpy(copy, mine, bytes);
  snprintf(name, sizeof(name), "%s.%s.%d", comm_path, tag, comm_rank);
  int err = nomp_b200_exchange_blob(name, 0, copy, bytes);
  free(copy);
  for (int r = 0; r < comm_size && !err; r++) {
    char *slot = (char *)all + (size_t)r * bytes;
    if (r == comm_rank) {
      memcpy(slot, mine, bytes);
      continue;
    }
    snprintf(name, sizeof(name), "%s.%s.%d", comm_path, tag, r);
    err = nomp_b200_exchange_blob(name, 1, slot, bytes);
  }
  if (!err) err = nomp_comm_barrier();
  snprintf(name, sizeof(name), "%s.%s.%d", comm_path, tag, comm_rank);
  unlink(name);
  return err;
}

/* Host-level barrier: one NCCL all-reduce on the default stream, waited for. */
int nomp_comm_barrier(void) {
  if (comm_size == 1) return 0;
  static int *token = NULL;
  if (token == NULL && cudaMalloc((void **)&token, sizeof(int)) != cudaSuccess)
    return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "NCCL failure: no memory for the barrier token.");
  cudaMemset(token, 0, sizeof(int));
  check_nccl(nccl.AllReduce(token, token, 1, ncclInt32, ncclSum, comm, 0));
  if (cudaStreamSynchronize(0) != cudaSuccess)
    return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "NCCL failure: barrier did not complete.");
  return 0;
}

int nomp_comm_finalize(void) {
  if (p2p.mine) {
    for (int r = 0; r < comm_size; r++)
      if (r != comm_rank && p2p.peers[r]) cudaIpcCloseMemHandle(p2p.peers[r]);
    cudaFree(p2p.mine);
    if (p2p.table_dev) cudaFree(p2p.table_dev);
    if (p2p.seq_dev) cudaFree(p2p.seq_dev);
    memset(&p2p, 0, sizeof(p2p));
  }
  if (comm) {
    nccl.CommDestroy(comm);
    comm = NULL;
  }
  comm_rank = 0, comm_size = 1;
  return 0;
}

NOMP_EXPORT int nomp_b200_comm_uses_nvlink_kernel(void) { return p2p.enabled; }

/* Peer description for a reduction kernel that all-reduces its own result (nompk_reduce_peers, nompk_ax_dot_peers_f64).
 * The number of the collective call is a counter in device memory that the kernels advance themselves.  Returns 0 when
 * the kernels cannot do it (one rank, no peer access, more ranks than a warp has lanes, or NOMP_COMM_FUSED=0) and the
 * caller must use nomp_comm_allreduce after the kernel. */
int nomp_comm_peers(void *peers_, void *error_host_mapped) {
  nompk_peers_t *peers = (nompk_peers_t *)peers_;
  static int fused = -1;
  if (fused < 0) {
    const char *e = getenv("NOMP_COMM_FUSED");
    fused = !(e && e[0] == '0');
  }
  if (comm_size == 1 || !p2p.enabled || !fused || comm_size > 32) return 0;
  peers->peer_xchg = (void *const *)p2p.table_dev;
  peers->rank = comm_rank, peers->world = comm_size;
  peers->seq = 0, peers->seq_dev = p2p.seq_dev;
  peers->error_host_mapped = (unsigned long long *)error_host_mapped;
  return 1;
}

int nomp_comm_allreduce(void *dev_scalar, int dtype, int op, void *result_host_mapped, unsigned long long host_seq,
                        void *error_host_mapped, void *stream, int *published) {
  *published = 0;
  if (comm_size == 1) return 0;
  if (p2p.enabled) {
    const nompk_peers_t peers = {(void *const *)p2p.table_dev, comm_rank, comm_size, 0, p2p.seq_dev,
                                 (unsigned long long *)error_host_mapped};
    int rc = nompk_allreduce_scalar_peers((nompk_red_op_t)op, (nompk_dtype_t)dtype, dev_scalar, result_host_mapped, host_seq,
                                          &peers, stream);
    if (rc != NOMPK_OK)
      return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "CUDA kernel library failure: %s.", nompk_last_error());
    *published = 1;
    return 0;
  }
  ncclDataType_t dt;
  switch (dtype) {
  case NOMPK_I32: dt = ncclInt32; break;
  case NOMPK_U32: dt = ncclUint32; break;
  case NOMPK_I64: dt = ncclInt64; break;
  case NOMPK_U64: dt = ncclUint64; break;
  case NOMPK_F32: dt = ncclFloat32; break;
  default: dt = ncclFloat64; break;
  }
  ncclRedOp_t rop = op == NOMP_PROD ? ncclProd : op == NOMP_MIN ? ncclMin : op == NOMP_MAX ? ncclMax : ncclSum;
  check_nccl(nccl.AllReduce(dev_scalar, dev_scalar, 1, dt, rop, comm, (cudaStream_t)stream));
  return 0;
}
