/*
 * nompk.h -- C ABI of libnompk, the hand-written sm_100a kernel library that sits under
 * libnomp's CUDA backend.
 *
 * libnomp (the reference, nomp-org/libnomp) has no hand-written device code: every kernel is a
 * string printed by loopy and compiled by NVRTC (reference backends/unified-cuda-hip-impl.h:96-141,
 * launched at :143-158).  This library is what replaces that generated code for the three loop
 * families that dominate spectral-element runs.  The bridge (libnomp_b200/python/nomp_bridge) recognises
 * the family at nomp_jit() time and the backend's knl_run() (reference include/nomp-impl.h:213-237)
 * calls one of the entry points below instead of cuLaunchKernel().
 *
 * Conventions
 *   - plain C: pointers are DEVICE pointers unless the name ends in _host; sizes are element counts;
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function is asynchronous on `stream` and returns 0 on success, or a negative NOMPK_E*
 *     code; nompk_last_error() gives the text of the last failure on the calling thread;
 *   - no entry point allocates device memory; workspaces are caller-owned.
 */
#ifndef NOMPK_H_
#define NOMPK_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NOMPK_VERSION 200

/* error codes */
#define NOMPK_OK 0
#define NOMPK_EINVAL (-1)   /* bad argument (unsupported dtype/op/n, NULL pointer, ...) */
#define NOMPK_ECUDA (-2)    /* a CUDA runtime call or a launch failed */
#define NOMPK_EUNSUPPORTED (-3)

/* Scalar element types.  These are the six types of the reference test matrix
 * (reference tests/nomp-generate-tests.h:1-49): int, unsigned, long, unsigned long, float, double. */
typedef enum {
  NOMPK_I32 = 0,
  NOMPK_U32 = 1,
  NOMPK_I64 = 2,
  NOMPK_U64 = 3,
  NOMPK_F32 = 4,
  NOMPK_F64 = 5
} nompk_dtype_t;

/* Elementwise map family.  Each op is one canonical loop body, written the way the reference tests
 * spell it (reference tests/nomp-api-200-impl.h:36-40, :64-68, :92-96; tests/nomp-api-600-impl.h:36-40).
 * y is read-modify-write, x and z are read-only, alpha/beta are HOST scalars of the element type. */
typedef enum {
  NOMPK_MAP_ADD   = 0, /* y[i] = y[i] + x[i]                 "a[i] += b[i]"          */
  NOMPK_MAP_SUB   = 1, /* y[i] = y[i] - x[i]                 "a[i] -= b[i]"          */
  NOMPK_MAP_MUL   = 2, /* y[i] = y[i] * x[i]                 "a[i] *= b[i]"          */
  NOMPK_MAP_AXPY  = 3, /* y[i] = y[i] + alpha * x[i]         "a[i] += alpha * b[i]"  */
  NOMPK_MAP_XPAY  = 4, /* y[i] = x[i] + alpha * y[i]         "p[i] = r[i] + beta * p[i]" */
  NOMPK_MAP_AXPBY = 5, /* y[i] = alpha * x[i] + beta * y[i]                          */
  NOMPK_MAP_SCALE = 6, /* y[i] = alpha * y[i]                                        */
  NOMPK_MAP_COPY  = 7, /* y[i] = x[i]                                                */
  NOMPK_MAP_FILL  = 8, /* y[i] = alpha                                               */
  NOMPK_MAP_ADD3  = 9, /* y[i] = x[i] + z[i]                 "c[i] = a[i] + b[i]"    */
  NOMPK_MAP_OP_COUNT
} nompk_map_op_t;

/* Reduction family.  SUM and PROD are the reference's two operators (reference src/reduction.c:3-22,
 * include/nomp-impl.h:99-102); MIN and MAX are additions named by the north star.  With y == NULL the
 * reduced value is x[i]; with y != NULL it is x[i] * y[i] (dot product for SUM). */
typedef enum {
  NOMPK_RED_SUM  = 0,
  NOMPK_RED_PROD = 1,
  NOMPK_RED_MIN  = 2,
  NOMPK_RED_MAX  = 3,
  NOMPK_RED_OP_COUNT
} nompk_red_op_t;

int nompk_version(void);
const char *nompk_last_error(void);
size_t nompk_dtype_size(nompk_dtype_t dt);

/* y <- op(y, x, z, alpha, beta) over n elements.  Unused operands may be NULL.
 * Replaces: the loopy-generated one-element-per-thread map kernel launched by
 * reference backends/unified-cuda-hip-impl.h:143-158 for the loops of tests 200/205/600.
 * fp results are bit-identical to the serial C loop compiled without FMA contraction. */
int nompk_map(nompk_map_op_t op, nompk_dtype_t dt, size_t n, void *y, const void *x, const void *z,
              const void *alpha_host, const void *beta_host, void *stream);

/* Bytes of device workspace nompk_reduce() needs (partials of two levels + ticket counters, 536 KiB).  The workspace
 * must be zero-filled once before first use; the kernels leave it ready for the next call on the same stream.
 * nompk_reduce_workspace_layout() returns the byte offsets {global ticket, group tickets, level-2 partials, level-1
 * partials}; generated reduction kernels (python/nomp_bridge/reduction.py) follow the same layout. */
size_t nompk_reduce_workspace_bytes(void);
void nompk_reduce_workspace_layout(size_t offsets[4]);

/* result[0] <- reduce_op over i of (y ? x[i]*y[i] : x[i]),  i in [0,n).  n == 0 writes the identity.
 * Single pass: per-thread accumulators, warp-shuffle tree, one partial per CTA, and two levels of atomic tickets fold
 * the partials with a fixed association (deterministic run to run).  `result` is a device
 * pointer (8-byte aligned); if result_host_mapped != NULL (device address of 16 bytes of pinned, mapped host memory)
 * the value is also stored at its bytes [0,8) and then host_seq at bytes [8,16), so the host needs no D2H copy and
 * may spin on the sequence number instead of synchronising the stream.
 * Replaces: loopy's per-block tree (reference python/reduction.py:30-134) + the D2H of all partials and
 * the serial host loop in reference src/reduction.c:33-88. */
int nompk_reduce(nompk_red_op_t op, nompk_dtype_t dt, size_t n, const void *x, const void *y,
                 void *result, void *result_host_mapped, unsigned long long host_seq, void *workspace,
                 void *stream);

/* nompk_reduce fused with the all-reduce of its result over `world` ranks (one GPU each): the CTA that finishes the
 * grid-wide fold stores the value straight into every peer's exchange buffer and folds the peers' values in rank
 * order -- the collective costs no launch of its own.  `peers` describes the exchange buffers exactly as for
 * nompk_allreduce_scalar below (same buffers, same sequence numbering; at most 32 ranks; NULL or world <= 1 = plain
 * nompk_reduce); result / result_host_mapped receive the all-reduced value, the host block is 24 bytes (error word). */
typedef struct {
  void *const *peer_xchg; /* DEVICE array of `world` pointers to the ranks' exchange buffers as mapped here */
  int rank, world;
  unsigned long long seq; /* number of this collective call: 1, 2, 3, ... in the same order on all ranks */
  /* Call number kept in DEVICE memory instead (non-NULL: `seq` is ignored): the kernel that finishes the reduction
   * takes *seq_dev + 1 as the number of the call and stores it back.  Nothing about the collective is then a launch
   * parameter, so a launch recorded in a CUDA graph can be replayed on every rank (include/nomp-b200.h:
   * nomp_b200_graph_*), and the host keeps no count.  One counter per rank, zero-initialised, touched only by
   * kernels on one stream. */
  unsigned long long *seq_dev;
  /* Device address of 8 bytes of mapped pinned host memory (may be NULL): receives the call number if a peer did not
   * arrive within 20 s, also when no result_host_mapped block was given (results that stay on the device). */
  unsigned long long *error_host_mapped;
} nompk_peers_t;
int nompk_reduce_peers(nompk_red_op_t op, nompk_dtype_t dt, size_t n, const void *x, const void *y, void *result,
                       void *result_host_mapped, unsigned long long host_seq, void *workspace,
                       const nompk_peers_t *peers, void *stream);

/* One-shot all-reduce of one scalar across `world` processes (one GPU each) over NVLink peer memory.
 * `value` (device, 8-byte aligned) holds this rank's contribution and receives the result, identical bit for bit on
 * every rank (contributions are folded in rank order).  peer_xchg is a DEVICE array of `world` pointers: entry r is
 * rank r's exchange buffer of nompk_allreduce_xchg_bytes(world) zero-initialised bytes, as mapped into THIS process
 * (own allocation for r == rank, cudaIpcOpenMemHandle otherwise).  `seq` numbers the collective calls: 1, 2, 3, ...
 * in the same order on all ranks.  result_host_mapped / host_seq as in nompk_reduce, except that the host block is
 * 24 bytes: if a peer does not arrive within 20 s the kernel gives up (no GPU hang) and stores `seq` at bytes [16,24).
 * Replaces: nothing in the reference (it has no collective); in this implementation it replaces ncclAllReduce for the
 * 4/8-byte result of a reduce clause (NCCL stays as the fallback when peer access is unavailable). */
size_t nompk_allreduce_xchg_bytes(int world);
int nompk_allreduce_scalar(nompk_red_op_t op, nompk_dtype_t dt, void *value, void *result_host_mapped,
                           unsigned long long host_seq, void *const *peer_xchg, int rank, int world,
                           unsigned long long seq, void *stream);
/* The same with the whole peer description (call number in device memory, error word): see nompk_peers_t. */
int nompk_allreduce_scalar_peers(nompk_red_op_t op, nompk_dtype_t dt, void *value, void *result_host_mapped,
                                 unsigned long long host_seq, const nompk_peers_t *peers, void *stream);

/* Local Poisson operator on E hexahedral spectral elements with n = N+1 points per direction, fp64:
 *   w_e = D^T_r (g1 Dr u + g2 Ds u + g3 Dt u) + D^T_s (g2 Dr u + g4 Ds u + g5 Dt u)
 *       + D^T_t (g3 Dr u + g5 Ds u + g6 Dt u)                         (Nekbone ax_e / CEED BK5 form)
 * Layouts: u, w  double[E][n][n][n]  (i fastest);  g  double[E][6][n][n][n]  (g1..g6 = G11,G12,G13,G22,G23,G33);
 *          D  double[n][n] row-major, D[a][l] = d phi_l / dx at node a.
 * The reference has no such operator (reference tests/sem.py:10-36 only tags loops); the canonical kernel
 * string that the bridge maps onto this entry point is in libnomp_b200/python/nomp_bridge/families.py.
 * Hand-written for n in {6, 8, 10, 12} (N = 5, 7, 9, 11); other n return NOMPK_EUNSUPPORTED and the bridge falls back to
 * NVRTC.  D is staged into __constant__ memory on `stream` unless NOMPK_AX_D_CACHED is set, which asserts
 * that the previous call with the same n used identical D values (and the same flags).
 * NOMPK_AX_D_ANTISYMMETRIC: the caller vouches that D[a][l] == -D[n-1-a][n-1-l] bit for bit (a differentiation matrix on
 * symmetric nodes: every Gauss-Lobatto-Legendre D is).  The six contractions then take their even-odd form, n^2/2 + 2n
 * instead of n^2 operations per line (n = 8 and 10: all six stages, 5 - 10 % faster; n = 12: four of them, 6 %; ignored for
 * n = 6).  Sums
 * are taken in another order: the result differs from the general path in the last bits (1e-16 relative), identically
 * in nompk_ax_f64, nompk_ax_dot*_f64 and nompk_ax_xpay_dot_peers_f64.  libnomp's backend decides it by reading D. */
#define NOMPK_AX_D_CACHED 1
#define NOMPK_AX_D_ANTISYMMETRIC 2
int nompk_ax_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w,
                 unsigned flags, void *stream);
/* The same operator fused with the dot product  result = u . (A u)  (the p.Ap of a conjugate-gradient iteration):
 * one pass over u and g instead of Ax followed by a 16 B/DOF dot kernel.  Evaluated in its energy form
 * sum over points of (ur wr + us ws + ut wt), which equals sum u_i w_i in exact arithmetic and needs no extra load.
 * result / result_host_mapped / host_seq / workspace as in nompk_reduce (same workspace, same publication protocol,
 * deterministic for a given grid). */
int nompk_ax_dot_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w, double *result,
                     double *result_host_mapped, unsigned long long host_seq, void *workspace, unsigned flags,
                     void *stream);
/* ... and with the all-reduce of u . (A u) over the ranks fused in as well (see nompk_reduce_peers). */
int nompk_ax_dot_peers_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w, double *result,
                           double *result_host_mapped, unsigned long long host_seq, void *workspace,
                           const nompk_peers_t *peers, unsigned flags, void *stream);
/* The first kernel of a conjugate-gradient iteration in one launch: the direction update in front of the operator,
 *   p <- r + beta p   (in place; multiply, then add -- the roundings of nompk_map(NOMPK_MAP_XPAY)),
 *   w <- A p,   result <- p . w   (energy form, all-reduced over `peers` like nompk_ax_dot_peers_f64),
 * 80 algorithmic B/DOF instead of 24 (map) + 64 (Ax) and one launch less.  beta is the host value, or beta_dev[0] from
 * device memory when beta_dev != NULL (a scalar that a previous kernel left there).  p and r: double[E][n][n][n], 16-byte
 * aligned.  The lane that loads a pair of p for the operator is the one that updates it; nothing else reads p.
 * Not in the reference (it has no Ax). */
int nompk_ax_xpay_dot_peers_f64(int n, size_t E, double *p, const double *r, double beta, const double *beta_dev,
                                const double *g, const double *D, double *w, double *result, double *result_host_mapped,
                                unsigned long long host_seq, void *workspace, const nompk_peers_t *peers, unsigned flags,
                                void *stream);
int nompk_ax_supported(int n);
/* Variant selector for benchmarking/profiling (0 = default). */
int nompk_ax_set_variant(int variant);

/* Gather-scatter (direct stiffness summation; gs_setup / gs_op of Nekbone and gslib): after nompk_gs_apply every local
 * degree of freedom v[i] holds the combination (op) of all degrees of freedom -- on this GPU and on every peer rank --
 * whose global id equals ids[i].  ids <= 0 do not take part.  The reference has no such operator; it is the step either
 * side of the local Ax in a spectral-element solve.
 *
 * Setup (all arrays are DEVICE arrays; every call synchronises `stream`):
 *   nompk_gs_create          sorts the n ids (n < 2^32 - 1) and finds the distinct ones
 *   nompk_gs_unique          ascending distinct ids of this rank, to be shown to the peers
 *   nompk_gs_match_peer      once per peer rank: intersects with that peer's distinct ids (any pointer this device
 *                            can read: a local copy or a CUDA-IPC mapping); *n_shared = number of shared ids
 *   nompk_gs_finalize_setup  builds the groups; *xchg_bytes = size of the exchange buffer this rank must allocate
 *                            (zero-filled; 0 when no id is shared with a peer)
 *   nompk_gs_recv_offsets    offsets[r] / counts[r]: where rank r's segment starts in MY exchange buffer (in values)
 *   nompk_gs_connect         peer_xchg[r] = rank r's exchange buffer as mapped into this process (own buffer at
 *                            [rank]); send_offsets[r] = what rank r reported as ITS offsets[my rank];
 *                            peer_totals[r] = the sum of rank r's counts (the stride of the two slots in ITS buffer)
 * Single-GPU use: create -> finalize_setup(0, 1) -> apply.
 *
 * nompk_gs_apply: asynchronous on `stream`, two launches (one when nothing is shared with a peer).  Copies of one id
 * are combined in ascending local index, ranks in ascending rank order: the result is deterministic and bit-identical
 * on every rank.  Partial results travel by stores into the peers' exchange buffers over NVLink; a peer that does not
 * arrive within 20 s makes the kernel give up (no GPU hang) and store the call number at *error_host_mapped (device
 * address of mapped pinned host memory, may be NULL).  All ranks must call apply in the same order. */
typedef struct nompk_gs nompk_gs_t;
int nompk_gs_create(const long long *ids, size_t n, nompk_gs_t **gs, void *stream);
int nompk_gs_unique(const nompk_gs_t *gs, const long long **ids, size_t *count);
int nompk_gs_match_peer(nompk_gs_t *gs, int peer, int world, const long long *peer_ids, size_t peer_count,
                        size_t *n_shared, void *stream);
int nompk_gs_finalize_setup(nompk_gs_t *gs, int rank, int world, size_t *xchg_bytes, void *stream);
int nompk_gs_recv_offsets(const nompk_gs_t *gs, size_t *offsets, size_t *counts);
int nompk_gs_connect(nompk_gs_t *gs, void *const *peer_xchg, const size_t *send_offsets, const size_t *peer_totals,
                     void *stream);
int nompk_gs_apply(nompk_gs_t *gs, nompk_red_op_t op, nompk_dtype_t dt, void *v, unsigned long long *error_host_mapped,
                   void *stream);
/* {n, distinct ids, groups, copies in groups, groups shared with peers, (group, peer) pairs, neighbours, shared ids} */
int nompk_gs_stats(const nompk_gs_t *gs, size_t out[8]);
void nompk_gs_destroy(nompk_gs_t *gs);

/* Launch bookkeeping used by bench.py's "gpu_launches" field: number of kernels this library has
 * launched since load (monotonic, process-wide). */
unsigned long long nompk_launch_count(void);

#ifdef __cplusplus
}
#endif

#endif /* NOMPK_H_ */
