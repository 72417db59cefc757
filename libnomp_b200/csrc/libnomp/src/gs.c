/* Gather-scatter handles of the extension API (include/nomp-b200.h: nomp_b200_gs_setup / nomp_b200_gs /
 * nomp_b200_gs_free).  SURVEY.md section 8 row f-2: the direct-stiffness summation either side of the local Ax.
 *
 * This file is host plumbing only: it copies the numbering to the device, lets libnompk build the groups
 * (include/nompk.h: nompk_gs_*), and -- with several ranks -- shows every rank the others' distinct ids and exchange
 * buffers through CUDA IPC mappings (NVLink / NVSwitch peer memory).  The small records needed for that (IPC handles,
 * segment offsets) travel through files next to the communicator's id file (comm.c: nomp_comm_allgather).  Once set
 * up, nomp_b200_gs() is two kernel launches on the backend stream and no host synchronisation.
 */
#include <cuda_runtime.h>

#include "nomp-b200.h"
#include "nomp-impl.h"
#include "nompk.h"

typedef struct {
  nompk_gs_t *gs;
  size_t n;
  void *xchg;                   /* this rank's exchange buffer */
  void *peers[64];              /* the peers' buffers as mapped here (NULL when not a neighbour) */
  unsigned long long *err_host; /* mapped pinned word: non-zero after a peer failed to arrive */
  unsigned long long *err_dev;
} gs_handle_t;

static gs_handle_t *handles = NULL;
static unsigned n_handles = 0;
static unsigned n_setups = 0; /* tags the rendezvous files of each collective setup; same sequence on every rank */

#define gs_cuda(call)                                                                                            \
  do {                                                                                                           \
    cudaError_t e_ = (call);                                                                                     \
    if (e_ != cudaSuccess) {                                                                                     \
      err = nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "CUDA %s failure: %s.", "gather-scatter setup", cudaGetErrorString(e_)); \
      goto done;                                                                                                 \
    }                                                                                                            \
  } while (0)

#define gs_nompk(call)                                                                                           \
  do {                                                                                                           \
    if ((call) != NOMPK_OK) {                                                                                    \
      err = nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "CUDA kernel library failure: %s.", nompk_last_error());     \
      goto done;                                                                                                 \
    }                                                                                                            \
  } while (0)

static void release(gs_handle_t *h) {
  const int rank = nomp_comm_rank(), size = nomp_comm_size();
  for (int r = 0; r < size && r < 64; r++)
    if (r != rank && h->peers[r]) cudaIpcCloseMemHandle(h->peers[r]);
  if (h->gs) nompk_gs_destroy(h->gs);
  if (h->xchg) cudaFree(h->xchg);
  if (h->err_host) cudaFreeHost(h->err_host);
  memset(h, 0, sizeof(*h));
}

typedef struct {
  int ok;
  size_t count;
  cudaIpcMemHandle_t mem;
} ids_record_t;

typedef struct {
  int ok, has_buffer;
  cudaIpcMemHandle_t mem;
  size_t total;       /* ids this rank shares with all its peers: the stride of the two slots of its buffer */
  size_t offsets[64]; /* where each peer's segment starts inside a slot */
} xchg_record_t;

NOMP_EXPORT int nomp_b200_gs_setup(int *handle, const long long *ids, size_t n) {
  void *stream = nomp_b200_stream();
  if (stream == NULL) return nomp_log(NOMP_INITIALIZE_FAILURE, NOMP_ERROR, "libnomp is not initialized.");
  if (handle == NULL || (n > 0 && ids == NULL))
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "nomp_b200_gs_setup needs a handle and %zu ids.", n);
  const int rank = nomp_comm_rank(), size = nomp_comm_size();
  if (size > 64) return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Gather-scatter supports at most 64 ranks.");

  int err = 0;
  gs_handle_t h;
  memset(&h, 0, sizeof(h));
  h.n = n;
  long long *ids_dev = NULL;
  void *opened[64] = {NULL};
  char tag[64];
  const unsigned setup = n_setups++;

  gs_cuda(cudaMalloc((void **)&ids_dev, (n ? n : 1) * sizeof(long long)));
  gs_cuda(cudaMemcpyAsync(ids_dev, ids, n * sizeof(long long), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  gs_nompk(nompk_gs_create(ids_dev, n, &h.gs, stream));
  cudaFree(ids_dev), ids_dev = NULL;
  gs_cuda(cudaHostAlloc((void **)&h.err_host, sizeof(*h.err_host), cudaHostAllocMapped));
  *h.err_host = 0;
  gs_cuda(cudaHostGetDevicePointer((void **)&h.err_dev, h.err_host, 0));

  size_t xchg_bytes = 0;
  if (size == 1) {
    gs_nompk(nompk_gs_finalize_setup(h.gs, 0, 1, &xchg_bytes, stream));
  } else {
    /* 1. everybody looks at everybody's distinct ids (peer memory, read in place) */
    const long long *mine = NULL;
    ids_record_t rec, all[64];
    memset(&rec, 0, sizeof(rec));
    gs_nompk(nompk_gs_unique(h.gs, &mine, &rec.count));
    rec.ok = rec.count == 0 || cudaIpcGetMemHandle(&rec.mem, (void *)mine) == cudaSuccess;
    snprintf(tag, sizeof(tag), "gs%u.ids", setup);
    if ((err = nomp_comm_allgather(tag, &rec, all, sizeof(rec)))) goto done;
    for (int r = 0; r < size; r++) {
      if (!all[r].ok) {
        err = nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "Gather-scatter across ranks needs CUDA IPC (rank %d has none).", r);
        goto done;
      }
    }
    for (int r = 0; r < size; r++) {
      if (r == rank || all[r].count == 0 || rec.count == 0) continue;
      gs_cuda(cudaIpcOpenMemHandle(&opened[r], all[r].mem, cudaIpcMemLazyEnablePeerAccess));
      gs_nompk(nompk_gs_match_peer(h.gs, r, size, (const long long *)opened[r], all[r].count, NULL, stream));
    }
    gs_nompk(nompk_gs_finalize_setup(h.gs, rank, size, &xchg_bytes, stream));
    /* nobody may free its ids before everybody has read them */
    if ((err = nomp_comm_barrier())) goto done;
    for (int r = 0; r < size; r++)
      if (opened[r]) cudaIpcCloseMemHandle(opened[r]), opened[r] = NULL;

    /* 2. exchange buffers: allocate, publish, map the neighbours' */
    xchg_record_t xr, xall[64];
    size_t counts[64];
    memset(&xr, 0, sizeof(xr));
    xr.ok = 1;
    gs_nompk(nompk_gs_recv_offsets(h.gs, xr.offsets, counts));
    for (int r = 0; r < size; r++) xr.total += counts[r];
    if (xchg_bytes > 0) {
      gs_cuda(cudaMalloc(&h.xchg, xchg_bytes));
      gs_cuda(cudaMemset(h.xchg, 0, xchg_bytes));
      gs_cuda(cudaDeviceSynchronize());
      xr.has_buffer = 1;
      xr.ok = cudaIpcGetMemHandle(&xr.mem, h.xchg) == cudaSuccess;
    }
    snprintf(tag, sizeof(tag), "gs%u.xchg", setup);
    if ((err = nomp_comm_allgather(tag, &xr, xall, sizeof(xr)))) goto done;
    size_t send_offsets[64] = {0}, peer_totals[64] = {0};
    for (int r = 0; r < size; r++) {
      if (!xall[r].ok) {
        err = nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "Gather-scatter across ranks needs CUDA IPC (rank %d has none).", r);
        goto done;
      }
      send_offsets[r] = xall[r].offsets[rank], peer_totals[r] = xall[r].total;
      if (r == rank) h.peers[r] = h.xchg;
      else if (counts[r] > 0) {
        if (!xall[r].has_buffer) {
          err = nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Ranks %d and %d disagree on their shared ids.", rank, r);
          goto done;
        }
        gs_cuda(cudaIpcOpenMemHandle(&h.peers[r], xall[r].mem, cudaIpcMemLazyEnablePeerAccess));
      }
    }
    if (xchg_bytes > 0) gs_nompk(nompk_gs_connect(h.gs, (void *const *)h.peers, send_offsets, peer_totals, stream));
    if ((err = nomp_comm_barrier())) goto done;
  }

  {
    unsigned slot = 0;
    while (slot < n_handles && handles[slot].gs) slot++;
    if (slot == n_handles) {
      handles = nomp_realloc(handles, gs_handle_t, n_handles + 8);
      memset(handles + n_handles, 0, 8 * sizeof(gs_handle_t));
      n_handles += 8;
    }
    handles[slot] = h;
    *handle = (int)slot;
  }
  return 0;

done:
  for (int r = 0; r < 64; r++)
    if (opened[r]) cudaIpcCloseMemHandle(opened[r]);
  if (ids_dev) cudaFree(ids_dev);
  release(&h);
  return err;
}

static gs_handle_t *lookup(int handle) {
  if (handle < 0 || (unsigned)handle >= n_handles || handles[handle].gs == NULL) return NULL;
  return &handles[handle];
}

NOMP_EXPORT int nomp_b200_gs(int handle, void *ptr, size_t unit_size, int type, const char *op) {
  gs_handle_t *h = lookup(handle);
  if (h == NULL)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Gather-scatter handle %d is not valid.", handle);
  if (*h->err_host)
    return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "Gather-scatter call %llu timed out waiting for a peer rank.", *h->err_host);
  nomp_mem_t *m = nomp_lookup_mem(ptr);
  if (m == NULL) return nomp_log(NOMP_USER_MAP_PTR_IS_INVALID, NOMP_ERROR, ERR_STR_USER_MAP_PTR_IS_INVALID, ptr);
  if (m->usize != unit_size || m->idx0 != 0 || m->idx1 < h->n)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR,
                    "Gather-scatter needs elements [0, %zu) of size %zu mapped; the mapping holds [%zu, %zu) of size %zu.",
                    h->n, unit_size, m->idx0, m->idx1, m->usize);
  int dtype = -1;
  if (type == NOMP_FLOAT) dtype = unit_size == 8 ? NOMPK_F64 : unit_size == 4 ? NOMPK_F32 : -1;
  else if (type == NOMP_INT) dtype = unit_size == 8 ? NOMPK_I64 : unit_size == 4 ? NOMPK_I32 : -1;
  else if (type == NOMP_UINT) dtype = unit_size == 8 ? NOMPK_U64 : unit_size == 4 ? NOMPK_U32 : -1;
  int rop = -1;
  if (op && !strcmp(op, "+")) rop = NOMPK_RED_SUM;
  else if (op && !strcmp(op, "*")) rop = NOMPK_RED_PROD;
  else if (op && !strcmp(op, "min")) rop = NOMPK_RED_MIN;
  else if (op && !strcmp(op, "max")) rop = NOMPK_RED_MAX;
  if (dtype < 0 || rop < 0)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR,
                    "Gather-scatter needs a 4- or 8-byte NOMP_INT / NOMP_UINT / NOMP_FLOAT and one of \"+\", \"*\", \"min\", \"max\".");
  nomp_check(nomp_cuda_before_write(m)); /* an asynchronous copy out of v may still be reading it */
  if (nompk_gs_apply(h->gs, (nompk_red_op_t)rop, (nompk_dtype_t)dtype, m->bptr, h->err_dev, nomp_b200_stream()) != NOMPK_OK)
    return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "CUDA kernel library failure: %s.", nompk_last_error());
  m->version = nomp_next_version();
  return 0;
}

NOMP_EXPORT int nomp_b200_gs_info(int handle, size_t out[8]) {
  gs_handle_t *h = lookup(handle);
  if (h == NULL || out == NULL)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Gather-scatter handle %d is not valid.", handle);
  nompk_gs_stats(h->gs, out);
  return 0;
}

NOMP_EXPORT int nomp_b200_gs_free(int handle) {
  gs_handle_t *h = lookup(handle);
  if (h == NULL)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Gather-scatter handle %d is not valid.", handle);
  cudaStreamSynchronize((cudaStream_t)nomp_b200_stream());
  /* a peer may still be storing into this rank's exchange buffer, or reading its flags */
  int err = nomp_comm_size() > 1 ? nomp_comm_barrier() : 0;
  release(h);
  return err;
}

/* nomp_sync: the asynchronous kernels report a peer that never arrived through the mapped error word */
int nomp_gs_check(void) {
  for (unsigned i = 0; i < n_handles; i++)
    if (handles[i].gs && *handles[i].err_host)
      return nomp_log(NOMP_CUDA_FAILURE, NOMP_ERROR, "Gather-scatter call %llu timed out waiting for a peer rank.",
                      *handles[i].err_host);
  return 0;
}

/* nomp_finalize: drop whatever the program did not free (collective when ranks > 1, like nomp_finalize itself) */
void nomp_gs_finalize(void) {
  for (unsigned i = 0; i < n_handles; i++)
    if (handles[i].gs) nomp_b200_gs_free((int)i);
  free(handles);
  handles = NULL, n_handles = 0, n_setups = 0;
}
