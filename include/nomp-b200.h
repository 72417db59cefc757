/* nomp-b200.h -- additions of the B200 implementation to libnomp.so's exported surface.  None of them is needed by
 * a program written against nomp.h; they exist for measurement (bench.py records CUDA events on the backend's own
 * stream) and for multi-GPU runs.
 *
 * Multi-GPU (one process per GPU).  The reference is one process <-> one device (reference src/nomp.c:5,
 * backends/unified-cuda-hip-impl.h:228) and has no collective.  Here nomp_init() joins an NCCL communicator when
 * NOMP_COMM_SIZE > 1 (with NOMP_COMM_RANK and NOMP_COMM_ID_FILE, a path on a file system all ranks share, e.g.
 * /dev/shm, through which rank 0 publishes the ncclUniqueId).  Each rank maps and loops over its own slice of the
 * vectors / elements; the ONLY collective is an ncclAllReduce of the scalar a reduce clause produces, issued on
 * the backend stream right after the single-pass device reduction. */
#ifndef LIBNOMP_B200_EXT_H_
#define LIBNOMP_B200_EXT_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cudaStream_t (as void*) on which this runtime issues copies and kernels; NULL before nomp_init(). */
void *nomp_b200_stream(void);
/* Asynchronous nomp_update for an EXISTING mapping, op = NOMP_TO or NOMP_FROM.  Host-to-device and device-to-host
 * copies run on two dedicated streams (both PCIe directions at once when the host memory is pinned); a copy starts
 * after all kernels issued so far, kernels issued later see the data of the NOMP_TO copies issued so far.  The host
 * range must not be touched until nomp_sync() returns.  (nomp_update itself stays blocking, as in the reference.) */
int nomp_b200_update_async(void *ptr, size_t start_index, size_t end_index, size_t unit_size, int op);
/* Device address that kernels receive for host pointer `hptr` (the address of host element 0), or NULL. */
void *nomp_b200_device_ptr(void *hptr);
/* Device-resident reduction results (off by default; returns the previous setting, a negative argument only queries).
 * While on, a reduce clause whose variable is a MAPPED host address (nomp_update) leaves its result -- all-reduced over
 * the ranks -- in the device copy of that variable, and nomp_run() returns without waiting for it: the host copy is
 * not touched until nomp_update(NOMP_FROM).  Kernels read such scalars from device memory as `alpha[0]` (pointer
 * arguments; the map and reduce skeletons keep their vectorised schedules for loop-invariant reads), so a whole
 * iteration of a solver can be enqueued without a host round trip.  A reduction variable that is not mapped behaves as
 * in the reference (result on the host when nomp_run returns, reference tests/nomp-api-500-impl.h:29-34). */
int nomp_b200_device_reductions(int enable);
/* CUDA graphs of nomp_run sequences.  Between graph_begin and graph_end every nomp_run is RECORDED on the backend stream
 * instead of executed (run the sequence once before capturing it: kernels load lazily, on their first launch); graph_launch replays the recording, asynchronously like nomp_run.  Integer
 * and floating-point arguments are frozen with the values they have during capture, pointers as the device addresses of
 * their mappings (which must outlive the graph) -- so what changes from one replay to the next lives in device memory:
 * reduction results (nomp_b200_device_reductions must be on for captured reduce clauses) and scalars read as alpha[0].
 * While capturing, nomp_update / nomp_sync / nomp_b200_update_async and reduce clauses that deliver to the host are
 * refused (NOMP_USER_INPUT_IS_INVALID); reduce clauses on more than one rank are refused as well (the call number of the
 * fused all-reduce would be frozen).  Up to 64 graphs; nomp_finalize releases them. */
int nomp_b200_graph_begin(void);
int nomp_b200_graph_end(int *graph);
int nomp_b200_graph_launch(int graph);
int nomp_b200_graph_free(int graph);
/* Kernels launched by this runtime since load: NVRTC-built kernels + libnompk launches. */
unsigned long long nomp_b200_launch_count(void);
/* Rank / size of the NCCL communicator (0 / 1 when single-process). */
int nomp_b200_comm_rank(void);
int nomp_b200_comm_size(void);
/* 1 if multi-GPU reduce clauses use libnompk's NVLink one-shot all-reduce kernel, 0 if they use ncclAllReduce. */
int nomp_b200_comm_uses_nvlink_kernel(void);
/* File rendezvous used to distribute the ncclUniqueId: rank 0 publishes `bytes` bytes through `path`, the other
 * ranks wait for the file and read them.  0 on success, a log id otherwise. */
int nomp_b200_exchange_blob(const char *path, int rank, void *blob, size_t bytes);
/* "kind=native family=map ..." descriptor of program `id` (valid until nomp_finalize), or NULL. */
const char *nomp_b200_prog_info(int id);
/* Gather-scatter (direct stiffness summation, the gs_setup / gs_op pair of Nekbone and gslib; not in the reference).
 * ids[i] > 0 is the global number of local degree of freedom i (ids <= 0 do not take part).  After nomp_b200_gs()
 * every element of the mapped vector holds the combination of all elements -- on this GPU and on every other rank of
 * the communicator -- that carry the same id.  Setup and free are collective (every rank calls them, in the same
 * order); nomp_b200_gs() is asynchronous like nomp_run(): kernels on the backend stream, partial results exchanged by
 * stores into the peers' memory over NVLink (needs CUDA peer access between the ranks' GPUs), no host
 * synchronisation.  `ptr` must be mapped from element 0 over at least n elements of `unit_size` bytes; `type` is
 * NOMP_INT, NOMP_UINT or NOMP_FLOAT; `op` is "+", "*", "min" or "max".  Copies are combined in ascending local index,
 * ranks in ascending rank order: results are deterministic and bit-identical on every rank.
 * nomp_b200_gs_info: {n, distinct ids, groups, copies in groups, groups shared with other ranks, (group, rank) pairs,
 * neighbour ranks, ids shared with other ranks}. */
int nomp_b200_gs_setup(int *handle, const long long *ids, size_t n);
int nomp_b200_gs(int handle, void *ptr, size_t unit_size, int type, const char *op);
int nomp_b200_gs_info(int handle, size_t info[8]);
int nomp_b200_gs_free(int handle);
/* On-disk JIT cache (directory $NOMP_JIT_CACHE_DIR, else $XDG_CACHE_HOME/libnomp_b200, else ~/.cache/libnomp_b200;
 * NOMP_JIT_CACHE=0 disables it).  Counters since load: {programs served from the cache, programs built by the
 * bridge and stored, CUBINs loaded from the cache, CUBINs compiled by NVRTC and stored}.  A transform or annotation
 * script is part of the key by its own text only: scripts that import other user modules need NOMP_JIT_CACHE=0 (or
 * a fresh directory) when those modules change. */
void nomp_b200_jit_cache_stats(unsigned long long counters[4]);
/* The entry store behind the cache, for diagnostics and tests: the directory in use (NULL when the cache is off;
 * re-reads the environment), and raw put / get of one entry "<hex>.<ext>".  get returns 1 when there is no intact entry
 * (missing, truncated, or its checksum trailer does not match). */
const char *nomp_b200_jit_cache_dir(void);
int nomp_b200_jit_cache_put(const char *hex, const char *ext, const void *data, size_t size);
int nomp_b200_jit_cache_get(const char *hex, const char *ext, void *buf, size_t cap, size_t *size);
/* The hash the cache keys are made of (SHA-256, lower-case hex, NUL-terminated); exported for the tests. */
void nomp_b200_sha256_hex(const void *data, size_t n, char hex[65]);

#ifdef __cplusplus
}
#endif

#endif /* LIBNOMP_B200_EXT_H_ */
