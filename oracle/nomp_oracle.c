/*
 * nomp_oracle.c -- CPU restatement of the libnomp hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (libnomp.so / libnompk.so) never links, loads or calls it and has no CPU fallback.
 *
 * What is restated, and from where (paths relative to the reference tree):
 *   - kernel semantics = "the C loop of the kernel string, run serially, statements in program order"
 *     (python/loopy_api.py:769-821 builds the kernel with seq_dependencies=True from exactly that loop;
 *     accepted operators python/loopy_api.py:24-48).  The map functions below are the loops of
 *     tests/nomp-api-200-impl.h:36-40 (+=), :64-68 (-=), :92-96 (*=) and of the north-star axpy.
 *   - reduce clause = sum / product of the right-hand side only, the incoming value of the accumulator is
 *     ignored and the result OVERWRITES the output (python/reduction.py:61-100 drops the lhs;
 *     src/reduction.c:3-22: identities 0 / 1, wrap-around integer arithmetic; type selection by
 *     (domain, size == 4) src/reduction.c:44-85).
 *   - Ax: NOT in the reference (tests/sem.py:10-36 only tags loops).  Restated from the definition in
 *     SURVEY.md 8(a-17) / include/nompk.h (Nekbone ax_e: local_grad3 -> geometric factors -> local_grad3_t).
 *     PARITY UNPINNED by the reference for this function; tests pin it with analytic properties instead, among them
 *     a known answer that owes nothing to any implementation (tests/ax_closed_form.py: the boundary fluxes of a
 *     harmonic polynomial on a sheared element with a full constant metric, zero at interior nodes).
 *   - gather-scatter: NOT in the reference either.  Restated from the definition of gslib's gs_op as Nekbone uses
 *     it (every copy of a global id receives the combination of all copies; ids <= 0 do not take part), with the
 *     association order include/nompk.h documents.  PARITY UNPINNED by the reference; pinned by closed forms
 *     (multiplicity counts of a box mesh, idempotence of min/max, conservation of the sum) and by the independent
 *     numpy restatement behind tests/golden/gs_cases.json.
 *
 * Pinning: tests/test_oracle.py checks these functions against every closed-form golden value the
 * reference tests hold for the path (tests/nomp-api-200/205/500/600-impl.h, listed in SURVEY.md 8c).
 *
 * Build: oracle/Makefile, `gcc -O2 -ffp-contract=off` (no FMA contraction: keep the C roundings).
 * The *_mt variants are the same loops split statically over all host cores with pthreads (this image has no
 * libgomp, so no `#pragma omp`): the stated stand-in for the reference's OpenCL-on-pocl CPU path, whose
 * pthread driver does the same thing and which cannot be installed here.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORACLE_API __attribute__((visibility("default")))

/* dtype codes shared with include/nompk.h */
enum { O_I32 = 0, O_U32 = 1, O_I64 = 2, O_U64 = 3, O_F32 = 4, O_F64 = 5 };
/* map ops shared with include/nompk.h */
enum { M_ADD = 0, M_SUB, M_MUL, M_AXPY, M_XPAY, M_AXPBY, M_SCALE, M_COPY, M_FILL, M_ADD3 };
/* reduce ops shared with include/nompk.h */
enum { R_SUM = 0, R_PROD, R_MIN, R_MAX };

/* ------------------------------------------------------------------------------------------------ */
/* static-partition parallel for over [0, n) on all online cores                                    */
/* ------------------------------------------------------------------------------------------------ */
#define ORACLE_MAX_THREADS 256
static int g_threads = 0;

ORACLE_API int oracle_num_threads(void) {
  if (g_threads <= 0) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    g_threads = c < 1 ? 1 : (c > ORACLE_MAX_THREADS ? ORACLE_MAX_THREADS : (int)c);
  }
  return g_threads;
}

ORACLE_API void oracle_set_num_threads(int t) { g_threads = t < 1 ? 1 : (t > ORACLE_MAX_THREADS ? ORACLE_MAX_THREADS : t); }

typedef void (*range_fn)(size_t lo, size_t hi, int tid, void *ctx);
typedef struct { range_fn fn; size_t lo, hi; int tid; void *ctx; } range_task;

static void *range_trampoline(void *p) {
  range_task *t = (range_task *)p;
  t->fn(t->lo, t->hi, t->tid, t->ctx);
  return NULL;
}

static void parallel_for(size_t n, range_fn fn, void *ctx) {
  const int nt = oracle_num_threads();
  pthread_t th[ORACLE_MAX_THREADS];
  range_task task[ORACLE_MAX_THREADS];
  for (int t = 0; t < nt; t++) {
    task[t].fn = fn, task[t].ctx = ctx, task[t].tid = t;
    task[t].lo = n * (size_t)t / nt, task[t].hi = n * (size_t)(t + 1) / nt;
    if (t > 0) pthread_create(&th[t], NULL, range_trampoline, &task[t]);
  }
  range_trampoline(&task[0]);
  for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------------------------------ */
/* maps                                                                                             */
/* ------------------------------------------------------------------------------------------------ */
#define MAP_BODY(T)                                                                                 \
  switch (op) {                                                                                     \
  case M_ADD: for (size_t i = 0; i < n; i++) y[i] = y[i] + x[i]; break;                             \
  case M_SUB: for (size_t i = 0; i < n; i++) y[i] = y[i] - x[i]; break;                             \
  case M_MUL: for (size_t i = 0; i < n; i++) y[i] = y[i] * x[i]; break;                             \
  case M_AXPY: for (size_t i = 0; i < n; i++) y[i] = y[i] + alpha * x[i]; break;                    \
  case M_XPAY: for (size_t i = 0; i < n; i++) y[i] = x[i] + alpha * y[i]; break;                    \
  case M_AXPBY: for (size_t i = 0; i < n; i++) y[i] = alpha * x[i] + beta * y[i]; break;            \
  case M_SCALE: for (size_t i = 0; i < n; i++) y[i] = alpha * y[i]; break;                          \
  case M_COPY: for (size_t i = 0; i < n; i++) y[i] = x[i]; break;                                   \
  case M_FILL: for (size_t i = 0; i < n; i++) y[i] = alpha; break;                                  \
  case M_ADD3: for (size_t i = 0; i < n; i++) y[i] = x[i] + z[i]; break;                            \
  default: return -1;                                                                               \
  }                                                                                                 \
  return 0;

#define MAP_FN(NAME, T)                                                                             \
  static int NAME(int op, size_t n, T *y, const T *x, const T *z, T alpha, T beta) { MAP_BODY(T) }

/* Integer kernels are evaluated in the unsigned type of the same width: identical bits, no signed-overflow UB. */
MAP_FN(map_u32, uint32_t)
MAP_FN(map_u64, uint64_t)
MAP_FN(map_f32, float)
MAP_FN(map_f64, double)

ORACLE_API int oracle_map(int op, int dtype, size_t n, void *y, const void *x, const void *z,
                          const void *alpha, const void *beta) {
  switch (dtype) {
  case O_I32:
  case O_U32: return map_u32(op, n, y, x, z, alpha ? *(const uint32_t *)alpha : 1u, beta ? *(const uint32_t *)beta : 1u);
  case O_I64:
  case O_U64: return map_u64(op, n, y, x, z, alpha ? *(const uint64_t *)alpha : 1u, beta ? *(const uint64_t *)beta : 1u);
  case O_F32: return map_f32(op, n, y, x, z, alpha ? *(const float *)alpha : 1.f, beta ? *(const float *)beta : 1.f);
  case O_F64: return map_f64(op, n, y, x, z, alpha ? *(const double *)alpha : 1.0, beta ? *(const double *)beta : 1.0);
  }
  return -1;
}

/* `a[i] += alpha * b[i]` on all host cores: the multi-threaded stand-in named in BASELINE.md section 3. */
typedef struct { double alpha; const double *x; double *y; } axpy_ctx;
static void axpy_range(size_t lo, size_t hi, int tid, void *p) {
  (void)tid;
  axpy_ctx *c = (axpy_ctx *)p;
  for (size_t i = lo; i < hi; i++) c->y[i] = c->y[i] + c->alpha * c->x[i];
}
ORACLE_API void oracle_axpy_f64_mt(size_t n, double alpha, const double *x, double *y) {
  axpy_ctx c = {alpha, x, y};
  parallel_for(n, axpy_range, &c);
}

/* ------------------------------------------------------------------------------------------------ */
/* reductions: the serial loop of the kernel string, result overwrites *out                          */
/* ------------------------------------------------------------------------------------------------ */
#define RED_FN(NAME, T, LO, HI)                                                                     \
  static void NAME(int op, size_t n, const T *x, const T *y, T *out) {                              \
    T acc;                                                                                          \
    switch (op) {                                                                                   \
    case R_SUM: acc = 0; for (size_t i = 0; i < n; i++) acc += y ? (T)(x[i] * y[i]) : x[i]; break;  \
    case R_PROD: acc = 1; for (size_t i = 0; i < n; i++) acc *= y ? (T)(x[i] * y[i]) : x[i]; break; \
    case R_MIN: acc = HI; for (size_t i = 0; i < n; i++) { T v = y ? (T)(x[i] * y[i]) : x[i]; if (v < acc) acc = v; } break; \
    default: acc = LO; for (size_t i = 0; i < n; i++) { T v = y ? (T)(x[i] * y[i]) : x[i]; if (v > acc) acc = v; } break; \
    }                                                                                               \
    *out = acc;                                                                                     \
  }

RED_FN(red_i32, int32_t, INT32_MIN, INT32_MAX)
RED_FN(red_u32, uint32_t, 0u, UINT32_MAX)
RED_FN(red_i64, int64_t, INT64_MIN, INT64_MAX)
RED_FN(red_u64, uint64_t, 0u, UINT64_MAX)
RED_FN(red_f32, float, -INFINITY, INFINITY)
RED_FN(red_f64, double, -INFINITY, INFINITY)

ORACLE_API int oracle_reduce(int op, int dtype, size_t n, const void *x, const void *y, void *out) {
  switch (dtype) {
  case O_I32:
    /* sums/products wrap: evaluate in unsigned (same bits) */
    if (op == R_SUM || op == R_PROD) red_u32(op, n, x, y, out); else red_i32(op, n, x, y, out);
    return 0;
  case O_U32: red_u32(op, n, x, y, out); return 0;
  case O_I64:
    if (op == R_SUM || op == R_PROD) red_u64(op, n, x, y, out); else red_i64(op, n, x, y, out);
    return 0;
  case O_U64: red_u64(op, n, x, y, out); return 0;
  case O_F32: red_f32(op, n, x, y, out); return 0;
  case O_F64: red_f64(op, n, x, y, out); return 0;
  }
  return -1;
}

/* Compensated fp64 sum / dot (Neumaier on long double): the yardstick for random data, where the serial
 * loop itself is only good to ~sqrt(n) ulp. */
ORACLE_API double oracle_sum_f64_compensated(size_t n, const double *x, const double *y) {
  long double s = 0.0L, c = 0.0L;
  for (size_t i = 0; i < n; i++) {
    long double v = y ? (long double)x[i] * (long double)y[i] : (long double)x[i];
    long double t = s + v;
    if (fabsl(s) >= fabsl(v)) c += (s - t) + v; else c += (v - t) + s;
    s = t;
  }
  return (double)(s + c);
}

/* all host cores: per-thread serial partial sums, combined in thread order */
typedef struct { const void *x, *y; double fpart[ORACLE_MAX_THREADS]; uint64_t ipart[ORACLE_MAX_THREADS]; } sum_ctx;
static void sum_f64_range(size_t lo, size_t hi, int tid, void *p) {
  sum_ctx *c = (sum_ctx *)p;
  const double *x = c->x, *y = c->y;
  double acc = 0.0;
  for (size_t i = lo; i < hi; i++) acc += y ? x[i] * y[i] : x[i];
  c->fpart[tid] = acc;
}
static void sum_i64_range(size_t lo, size_t hi, int tid, void *p) {
  sum_ctx *c = (sum_ctx *)p;
  const uint64_t *x = c->x, *y = c->y;
  uint64_t acc = 0;
  for (size_t i = lo; i < hi; i++) acc += y ? x[i] * y[i] : x[i];
  c->ipart[tid] = acc;
}
ORACLE_API double oracle_sum_f64_mt(size_t n, const double *x, const double *y) {
  static sum_ctx c;
  c.x = x, c.y = y;
  parallel_for(n, sum_f64_range, &c);
  double total = 0.0;
  for (int t = 0; t < oracle_num_threads(); t++) total += c.fpart[t];
  return total;
}
ORACLE_API int64_t oracle_sum_i64_mt(size_t n, const int64_t *x, const int64_t *y) {
  static sum_ctx c;
  c.x = x, c.y = y;
  parallel_for(n, sum_i64_range, &c);
  uint64_t total = 0;
  for (int t = 0; t < oracle_num_threads(); t++) total += c.ipart[t];
  return (int64_t)total;
}

/* ------------------------------------------------------------------------------------------------ */
/* Ax: local Poisson operator, layouts as in include/nompk.h                                         */
/*   u, w: [E][n][n][n] (i fastest)   g: [E][6][n^3]   D: [n][n] row-major                            */
/* ------------------------------------------------------------------------------------------------ */
#define AX_MAX_N 16

#define AX_ELEMENT(REAL)                                                                            \
  REAL ur[AX_MAX_N * AX_MAX_N * AX_MAX_N], us[AX_MAX_N * AX_MAX_N * AX_MAX_N],                      \
      ut[AX_MAX_N * AX_MAX_N * AX_MAX_N];                                                           \
  const int n2 = n * n, n3 = n * n * n;                                                             \
  const double *ue = u + (size_t)e * n3, *ge = g + (size_t)e * 6 * n3;                              \
  double *we = w + (size_t)e * n3;                                                                  \
  for (int k = 0; k < n; k++)                                                                       \
    for (int j = 0; j < n; j++)                                                                     \
      for (int i = 0; i < n; i++) {                                                                 \
        REAL r = 0, s = 0, t = 0;                                                                   \
        for (int l = 0; l < n; l++) {                                                               \
          r += (REAL)D[i * n + l] * (REAL)ue[k * n2 + j * n + l];                                   \
          s += (REAL)D[j * n + l] * (REAL)ue[k * n2 + l * n + i];                                   \
          t += (REAL)D[k * n + l] * (REAL)ue[l * n2 + j * n + i];                                   \
        }                                                                                           \
        const int id = k * n2 + j * n + i;                                                          \
        const REAL g1 = ge[id], g2 = ge[n3 + id], g3 = ge[2 * n3 + id], g4 = ge[3 * n3 + id],       \
                   g5 = ge[4 * n3 + id], g6 = ge[5 * n3 + id];                                      \
        ur[id] = g1 * r + g2 * s + g3 * t;                                                          \
        us[id] = g2 * r + g4 * s + g5 * t;                                                          \
        ut[id] = g3 * r + g5 * s + g6 * t;                                                          \
      }                                                                                             \
  for (int k = 0; k < n; k++)                                                                       \
    for (int j = 0; j < n; j++)                                                                     \
      for (int i = 0; i < n; i++) {                                                                 \
        REAL acc = 0;                                                                               \
        for (int l = 0; l < n; l++) {                                                               \
          acc += (REAL)D[l * n + i] * ur[k * n2 + j * n + l];                                       \
          acc += (REAL)D[l * n + j] * us[k * n2 + l * n + i];                                       \
          acc += (REAL)D[l * n + k] * ut[l * n2 + j * n + i];                                       \
        }                                                                                           \
        we[k * n2 + j * n + i] = (double)acc;                                                       \
      }

static void ax_element_f64(int n, size_t e, const double *u, const double *g, const double *D, double *w) {
  AX_ELEMENT(double)
}

static void ax_element_ld(int n, size_t e, const double *u, const double *g, const double *D, double *w) {
  AX_ELEMENT(long double)
}

/* plain fp64 arithmetic, serial over elements */
ORACLE_API int oracle_ax_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w) {
  if (n < 2 || n > AX_MAX_N) return -1;
  for (size_t e = 0; e < E; e++) ax_element_f64(n, e, u, g, D, w);
  return 0;
}

/* long double accumulation (the compensated yardstick for random data) */
ORACLE_API int oracle_ax_f64_extended(int n, size_t E, const double *u, const double *g, const double *D,
                                      double *w) {
  if (n < 2 || n > AX_MAX_N) return -1;
  for (size_t e = 0; e < E; e++) ax_element_ld(n, e, u, g, D, w);
  return 0;
}

/* plain fp64, elements spread over all host cores */
typedef struct { int n; const double *u, *g, *D; double *w; } ax_ctx;
static void ax_range(size_t lo, size_t hi, int tid, void *p) {
  (void)tid;
  ax_ctx *c = (ax_ctx *)p;
  for (size_t e = lo; e < hi; e++) ax_element_f64(c->n, e, c->u, c->g, c->D, c->w);
}
ORACLE_API int oracle_ax_f64_mt(int n, size_t E, const double *u, const double *g, const double *D, double *w) {
  if (n < 2 || n > AX_MAX_N) return -1;
  ax_ctx c = {n, u, g, D, w};
  parallel_for(E, ax_range, &c);
  return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* deterministic synthetic data                                                      */
/* ------------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

/* counter-based: value i depends only on (seed, i), so any slice can be generated independently */
static inline uint64_t mix_at(uint64_t seed, uint64_t i) {
  uint64_t s = seed + i * 0x9e3779b97f4a7c15ull;
  return splitmix64(&s);
}

/* Set X: small integers stored as doubles in [lo, hi]: every fp64 sum/product order is exact. */
typedef struct { void *a; uint64_t seed; size_t first; int lo, hi; double flo, fhi; } fill_ctx;
static void fill_int_range(size_t b, size_t e, int tid, void *p) {
  (void)tid;
  fill_ctx *c = (fill_ctx *)p;
  double *a = c->a;
  const uint64_t span = (uint64_t)(c->hi - c->lo + 1);
  for (size_t i = b; i < e; i++) a[i] = (double)(c->lo + (int)(mix_at(c->seed, c->first + i) % span));
}
ORACLE_API void oracle_fill_int_f64(double *a, size_t n, uint64_t seed, int lo, int hi, size_t first) {
  fill_ctx c = {a, seed, first, lo, hi, 0, 0};
  parallel_for(n, fill_int_range, &c);
}

/* Set R: uniform doubles in [lo, hi). */
static void fill_uniform_range(size_t b, size_t e, int tid, void *p) {
  (void)tid;
  fill_ctx *c = (fill_ctx *)p;
  double *a = c->a;
  for (size_t i = b; i < e; i++) {
    const double r = (double)(mix_at(c->seed, c->first + i) >> 11) * (1.0 / 9007199254740992.0);
    a[i] = c->flo + (c->fhi - c->flo) * r;
  }
}
ORACLE_API void oracle_fill_uniform_f64(double *a, size_t n, uint64_t seed, double lo, double hi, size_t first) {
  fill_ctx c = {a, seed, first, 0, 0, lo, hi};
  parallel_for(n, fill_uniform_range, &c);
}

/* full-range int64 (wrap-around sums are associative -> bit-exact in any order) */
static void fill_i64_range(size_t b, size_t e, int tid, void *p) {
  (void)tid;
  fill_ctx *c = (fill_ctx *)p;
  int64_t *a = c->a;
  for (size_t i = b; i < e; i++) a[i] = (int64_t)mix_at(c->seed, c->first + i);
}
ORACLE_API void oracle_fill_i64(int64_t *a, size_t n, uint64_t seed, size_t first) {
  fill_ctx c = {a, seed, first, 0, 0, 0, 0};
  parallel_for(n, fill_i64_range, &c);
}

/* Gauss-Lobatto-Legendre nodes and the n x n derivative matrix D[a][l] = l_l'(x_a) (row-major). */
static double legendre(int N, double x, double *dP) {
  double p0 = 1.0, p1 = x;
  if (N == 0) { if (dP) *dP = 0.0; return 1.0; }
  for (int k = 2; k <= N; k++) {
    const double pk = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
    p0 = p1; p1 = pk;
  }
  if (dP) *dP = N * (x * p1 - p0) / (x * x - 1.0);
  return p1;
}

ORACLE_API int oracle_gll_derivative(int n, double *D, double *nodes_out) {
  if (n < 2 || n > AX_MAX_N) return -1;
  const int N = n - 1;
  double x[AX_MAX_N];
  x[0] = -1.0; x[N] = 1.0;
  for (int i = 1; i < N; i++) {
    double xi = -cos(M_PI * i / N);           /* Chebyshev-Lobatto start */
    for (int it = 0; it < 100; it++) {        /* Newton on (1-x^2) P_N'(x), i.e. on q(x) = P_{N-1}(x) - x P_N(x) ... */
      double dP; const double P = legendre(N, xi, &dP);
      /* P_N'' from the Legendre ODE: (1-x^2) P'' - 2x P' + N(N+1) P = 0 */
      const double d2P = (2.0 * xi * dP - N * (N + 1.0) * P) / (1.0 - xi * xi);
      const double dx = dP / d2P;
      xi -= dx;
      if (fabs(dx) < 1e-16) break;
    }
    x[i] = xi;
  }
  for (int a = 0; a < n; a++)
    for (int l = 0; l < n; l++) {
      double v;
      if (a == l) {
        if (a == 0) v = -N * (N + 1.0) / 4.0;
        else if (a == N) v = N * (N + 1.0) / 4.0;
        else v = 0.0;
      } else {
        v = legendre(N, x[a], NULL) / (legendre(N, x[l], NULL) * (x[a] - x[l]));
      }
      D[a * n + l] = v;
    }
  if (nodes_out) memcpy(nodes_out, x, sizeof(double) * n);
  return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* gather-scatter (direct stiffness summation)                                                      */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { long long id; size_t idx; } gs_key;
static int gs_key_cmp(const void *a, const void *b) {
  const gs_key *x = a, *y = b;
  if (x->id != y->id) return x->id < y->id ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

#define GS_FN(NAME, T)                                                                              \
  static T NAME##_op(int op, T a, T b) {                                                            \
    switch (op) {                                                                                   \
    case R_SUM: return (T)(a + b);                                                                  \
    case R_PROD: return (T)(a * b);                                                                 \
    case R_MIN: return b < a ? b : a;                                                               \
    default: return b > a ? b : a;                                                                  \
    }                                                                                               \
  }                                                                                                 \
  static void NAME(int op, const gs_key *k, size_t lo, size_t hi, T *v, int nseg, const size_t *seg) { \
    /* copies [lo, hi) of one id, ascending index: fold inside each segment (rank), then the segment \
     * partials in segment order */                                                                 \
    T total = 0;                                                                                    \
    int have_total = 0;                                                                             \
    size_t i = lo;                                                                                  \
    for (int s = 0; s < nseg && i < hi; s++) {                                                      \
      if (k[i].idx >= seg[s + 1]) continue;                                                         \
      T part = v[k[i].idx];                                                                         \
      for (i++; i < hi && k[i].idx < seg[s + 1]; i++) part = NAME##_op(op, part, v[k[i].idx]);      \
      total = have_total ? NAME##_op(op, total, part) : part;                                       \
      have_total = 1;                                                                               \
    }                                                                                               \
    for (i = lo; i < hi; i++) v[k[i].idx] = total;                                                  \
  }

GS_FN(gs_i32, int32_t)
GS_FN(gs_u32, uint32_t)
GS_FN(gs_i64, int64_t)
GS_FN(gs_u64, uint64_t)
GS_FN(gs_f32, float)
GS_FN(gs_f64, double)

/* v[i] <- op over { v[j] : ids[j] == ids[i] } for ids[i] > 0.  The n values are the concatenation of nseg
 * segments (one per rank), seg[0] = 0 <= ... <= seg[nseg] = n; nseg = 1 is the single-GPU case. */
ORACLE_API int oracle_gs(int op, int dtype, const long long *ids, size_t n, void *v, int nseg, const size_t *seg) {
  const size_t one[2] = {0, n};
  if (nseg <= 0 || seg == NULL) nseg = 1, seg = one;
  gs_key *k = malloc((n ? n : 1) * sizeof(*k));
  if (!k) return -1;
  size_t m = 0;
  for (size_t i = 0; i < n; i++)
    if (ids[i] > 0) k[m].id = ids[i], k[m].idx = i, m++;
  qsort(k, m, sizeof(*k), gs_key_cmp);
  int err = 0;
  for (size_t lo = 0; lo < m && !err;) {
    size_t hi = lo + 1;
    while (hi < m && k[hi].id == k[lo].id) hi++;
    if (hi - lo > 1) {
      switch (dtype) {
      case O_I32: if (op == R_SUM || op == R_PROD) gs_u32(op, k, lo, hi, v, nseg, seg); else gs_i32(op, k, lo, hi, v, nseg, seg); break;
      case O_U32: gs_u32(op, k, lo, hi, v, nseg, seg); break;
      case O_I64: if (op == R_SUM || op == R_PROD) gs_u64(op, k, lo, hi, v, nseg, seg); else gs_i64(op, k, lo, hi, v, nseg, seg); break;
      case O_U64: gs_u64(op, k, lo, hi, v, nseg, seg); break;
      case O_F32: gs_f32(op, k, lo, hi, v, nseg, seg); break;
      case O_F64: gs_f64(op, k, lo, hi, v, nseg, seg); break;
      default: err = -1;
      }
    }
    lo = hi;
  }
  free(k);
  return err;
}
