/* Conjugate gradients on the local (block-diagonal) spectral-element Poisson operator, written against the public
 * libnomp API the way nompcc-generated code calls it (nomp_init / nomp_update / nomp_jit / nomp_run / nomp_finalize).
 *
 * Per iteration three kernels, 136 B/DOF instead of the 168 B/DOF of the textbook sequence (1 Ax, 2 dots, 3 axpys):
 *     w = A p   and   pAp = p.w              one launch: the canonical Ax + dot kernel string -> nompk_ax_dot_f64
 *     x += a p; r -= a w; rr = r.r           one launch: elementwise updates fused with the reduce clause (NVRTC skeleton)
 *     p = r + b p                            one launch: nompk_map (XPAY)
 * With NOMP_COMM_SIZE > 1 every rank owns E elements of a larger mesh; the two dot products are all-reduced by the
 * runtime, nothing else changes (the local operator needs no halo exchange).
 *
 *   usage: cg_poisson [E [n [max_iter [tol [host|fused|device|device3|device_fused|graph [check_every]]]]]]  + the usual --nomp-* flags
 *          (prints one JSON object per line)
 *
 * "device_fused" is "fused" with the scalars in device memory: xpay + Ax + dot reading beta[0], the update with alpha
 * folded in, and a one-thread kernel for beta -- three launches, 128 B/DOF, no round trip.
 * "fused" folds the direction update into the operator (the canonical xpay + Ax + dot kernel string ->
 * nompk_ax_xpay_dot_peers_f64): two launches per iteration and 128 instead of 136 B/DOF, scalars on the host.
 * "device" keeps every scalar of the iteration in device memory (include/nomp-b200.h: nomp_b200_device_reductions):
 * the two dot products leave their results in mapped variables, alpha and beta are computed by one-iteration kernels,
 * the updates read them as alpha[0] / beta[0] -- five launches per iteration and no host round trip; the host fetches
 * the residual every `check_every` iterations (default 10) to test for convergence.
 * "device3" folds alpha and beta into their consumers (x[i] += (rr[0] / pap[0]) * p[i], ...) and swaps the roles of the
 * two residual scalars on the host instead of copying one to the other: three launches per iteration, the same number as
 * "host", and still no round trip (it prints no per-iteration trace).
 * "graph" is "device3" captured once as a CUDA graph of two iterations (the residual scalars swap names every iteration) and
 * replayed (include/nomp-b200.h: nomp_b200_graph_*): six launches per cudaGraphLaunch, on any number of ranks -- the fused
 * all-reduce counts its calls in device memory, so nothing in the graph is specific to one iteration.
 */
#define _POSIX_C_SOURCE 200809L
#define _DEFAULT_SOURCE
#include "sem_common.h"
#include "nomp-b200.h"

/* counter-based generator shared with the tests (splitmix64) */
static double uniform(unsigned long long seed, unsigned long long i) {
  unsigned long long z = seed + (i + 1) * 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

static const char *UPDATE_SRC =
    "void cg_update(double *x, double *r, const double *p, const double *w, double alpha, int N, double *rr) {\n"
    "  for (int i = 0; i < N; i++) { x[i] += alpha * p[i]; r[i] -= alpha * w[i]; rr[0] += r[i] * r[i]; }\n}\n";
static const char *XPAY_SRC =
    "void cg_direction(double *p, const double *r, double beta, int N) { for (int i = 0; i < N; i++) p[i] = r[i] + beta * p[i]; }\n";
static const char *DOT_SRC =
    "void cg_dot(const double *a, const double *b, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i] * b[i]; }\n";

/* the same iteration with its scalars in device memory */
static const char *ALPHA_SRC =
    "void cg_alpha(double *alpha, const double *rr, const double *pap, double *trace, int slot) {\n"
    "  for (int i = 0; i < 1; i++) { alpha[i] = rr[i] / pap[i]; trace[3 * slot] = pap[i]; trace[3 * slot + 1] = alpha[i]; }\n}\n";
static const char *UPDATE_DEV_SRC =
    "void cg_update_d(double *x, double *r, const double *p, const double *w, const double *alpha, int N, double *rr_new) {\n"
    "  for (int i = 0; i < N; i++) { x[i] += alpha[0] * p[i]; r[i] -= alpha[0] * w[i]; rr_new[0] += r[i] * r[i]; }\n}\n";
static const char *BETA_SRC =
    "void cg_beta(double *beta, double *rr, const double *rr_new, double *trace, int slot) {\n"
    "  for (int i = 0; i < 1; i++) { beta[i] = rr_new[i] / rr[i]; rr[i] = rr_new[i]; trace[3 * slot + 2] = rr_new[i]; }\n}\n";
static const char *XPAY_DEV_SRC =
    "void cg_direction_d(double *p, const double *r, const double *beta, int N) { for (int i = 0; i < N; i++) p[i] = r[i] + beta[0] * p[i]; }\n";

static const char *UPDATE_DEV3_SRC =
    "void cg_update_3(double *x, double *r, const double *p, const double *w, const double *rr, const double *pap, int N,\n"
    "                 double *rr_new) {\n"
    "  for (int i = 0; i < N; i++) { x[i] += (rr[0] / pap[0]) * p[i]; r[i] -= (rr[0] / pap[0]) * w[i]; rr_new[0] += r[i] * r[i]; }\n}\n";
static const char *XPAY_DEV3_SRC =
    "void cg_direction_3(double *p, const double *r, const double *rr_new, const double *rr, int N) {\n"
    "  for (int i = 0; i < N; i++) p[i] = r[i] + (rr_new[0] / rr[0]) * p[i];\n}\n";

static const char *BETA2_SRC =
    "void cg_beta2(double *beta, double *rr, const double *rr_new) {\n"
    "  for (int i = 0; i < 1; i++) { beta[i] = rr_new[i] / rr[i]; rr[i] = rr_new[i]; }\n}\n";

int main(int argc, const char **argv) {
  int E = 1024, n = 8, max_iter = 200, device_scalars = 0, check_every = 10, fused = 0;
  double tol = 1e-10;
  int pos = 0;
  for (int i = 1; i < argc; i++) {
    if (!strncmp(argv[i], "--nomp", 6)) { i++; continue; }
    if (pos == 0) E = atoi(argv[i]);
    else if (pos == 1) n = atoi(argv[i]);
    else if (pos == 2) max_iter = atoi(argv[i]);
    else if (pos == 3) tol = atof(argv[i]);
    else if (pos == 4)
      device_scalars = !strcmp(argv[i], "device") ? 1 : !strcmp(argv[i], "device3") ? 3 : !strcmp(argv[i], "device_fused") ? 4 :
                       !strcmp(argv[i], "graph") ? 5 : 0,
      fused = !strcmp(argv[i], "fused");
    else if (pos == 5) check_every = atoi(argv[i]) > 0 ? atoi(argv[i]) : 1;
    pos++;
  }
  CHECK(nomp_init(argc, argv));
  const char *rank_s = getenv("NOMP_COMM_RANK");
  const int rank = rank_s ? atoi(rank_s) : 0;
  const size_t n3 = (size_t)n * n * n, N = (size_t)E * n3;
  const int Ni = (int)N;
  double *x = calloc(N, 8), *r = calloc(N, 8), *p = calloc(N, 8), *w = calloc(N, 8), *xt = calloc(N, 8);
  double *g = calloc(6 * N, 8), *D = calloc((size_t)n * n, 8);
  gll_derivative(n, D);
  const unsigned long long off = (unsigned long long)rank * N;
  for (size_t i = 0; i < N; i++) xt[i] = uniform(11, off + i) - 0.5;
  for (size_t e = 0; e < (size_t)E; e++)      /* symmetric positive definite metric: dominant diagonal entries */
    for (int f = 0; f < 6; f++)
      for (size_t q = 0; q < n3; q++) {
        const size_t i = (e * 6 + f) * n3 + q;
        const double v = uniform(13, 6 * off + i);
        g[i] = (f == 0 || f == 3 || f == 5) ? 1.0 + 0.5 * v : 0.2 * (v - 0.5);
      }

  double *arrays[] = {x, r, p, w, xt};
  for (int a = 0; a < 5; a++) CHECK(nomp_update(arrays[a], 0, N, 8, NOMP_TO));
  CHECK(nomp_update(g, 0, 6 * N, 8, NOMP_TO));
  CHECK(nomp_update(D, 0, (size_t)n * n, 8, NOMP_TO));

  const char *none[1] = {NULL};
  const char *red_pap[4] = {"reduce", "pap", "+", NULL}, *red_rr[4] = {"reduce", "rr", "+", NULL}, *red_s[4] = {"reduce", "s", "+", NULL};
  int id_ax = -1, id_axdot = -1, id_upd = -1, id_dir = -1, id_dot = -1;
  CHECK(nomp_jit(&id_ax, AX_SRC, none, 6, "w", sizeof(double), NOMP_PTR, "u", sizeof(double), NOMP_PTR, "g", sizeof(double), NOMP_PTR,
                 "D", sizeof(double), NOMP_PTR, "E", sizeof(int), NOMP_INT, "n", sizeof(int), NOMP_INT | NOMP_JIT, &n));
  CHECK(nomp_jit(&id_axdot, AX_DOT_SRC, red_pap, 7, "w", sizeof(double), NOMP_PTR, "u", sizeof(double), NOMP_PTR, "g", sizeof(double),
                 NOMP_PTR, "D", sizeof(double), NOMP_PTR, "E", sizeof(int), NOMP_INT, "n", sizeof(int), NOMP_INT | NOMP_JIT, &n, "pap",
                 sizeof(double), NOMP_FLOAT));
  CHECK(nomp_jit(&id_upd, UPDATE_SRC, red_rr, 7, "x", sizeof(double), NOMP_PTR, "r", sizeof(double), NOMP_PTR, "p", sizeof(double),
                 NOMP_PTR, "w", sizeof(double), NOMP_PTR, "alpha", sizeof(double), NOMP_FLOAT, "N", sizeof(int), NOMP_INT, "rr",
                 sizeof(double), NOMP_FLOAT));
  CHECK(nomp_jit(&id_dir, XPAY_SRC, none, 4, "p", sizeof(double), NOMP_PTR, "r", sizeof(double), NOMP_PTR, "beta", sizeof(double),
                 NOMP_FLOAT, "N", sizeof(int), NOMP_INT));
  CHECK(nomp_jit(&id_dot, DOT_SRC, red_s, 4, "a", sizeof(double), NOMP_PTR, "b", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT,
                 "s", sizeof(double), NOMP_FLOAT));

  /* b = A x_true  ->  r = b (x0 = 0), p = r */
  CHECK(nomp_run(id_ax, r, xt, g, D, &E));
  double one_beta = 0.0;
  CHECK(nomp_run(id_dir, p, r, &one_beta, &Ni));
  double rr = 0, rr0, pap = 0;
  CHECK(nomp_run(id_dot, r, r, &Ni, &rr));
  rr0 = rr;
  printf("{\"E_per_rank\": %d, \"n\": %d, \"dof_per_rank\": %zu, \"rr0\": %.17g}\n", E, n, N, rr0);

  /* scalars of the device-resident variant: one mapped double each; trace = {pAp, alpha, rr} of the first iterations */
  double *pap_d = calloc(1, 8), *alpha_d = calloc(1, 8), *beta_d = calloc(1, 8), *rr_d = calloc(1, 8), *rrn_d = calloc(1, 8);
  double *trace = calloc(18, 8);
  double *scalars[] = {pap_d, alpha_d, beta_d, rr_d, rrn_d};
  int id_alpha = -1, id_updd = -1, id_beta = -1, id_dird = -1, id_upd3 = -1, id_dir3 = -1, id_fused_dev = -1, id_beta2 = -1;
  if (device_scalars) {
    const char *red_rrn[4] = {"reduce", "rr_new", "+", NULL};
    rr_d[0] = rr;
    for (int a = 0; a < 5; a++) CHECK(nomp_update(scalars[a], 0, 1, 8, NOMP_TO));
    CHECK(nomp_update(trace, 0, 18, 8, NOMP_TO));
    CHECK(nomp_jit(&id_alpha, ALPHA_SRC, none, 5, "alpha", sizeof(double), NOMP_PTR, "rr", sizeof(double), NOMP_PTR, "pap",
                   sizeof(double), NOMP_PTR, "trace", sizeof(double), NOMP_PTR, "slot", sizeof(int), NOMP_INT));
    CHECK(nomp_jit(&id_updd, UPDATE_DEV_SRC, red_rrn, 7, "x", sizeof(double), NOMP_PTR, "r", sizeof(double), NOMP_PTR, "p",
                   sizeof(double), NOMP_PTR, "w", sizeof(double), NOMP_PTR, "alpha", sizeof(double), NOMP_PTR, "N", sizeof(int),
                   NOMP_INT, "rr_new", sizeof(double), NOMP_FLOAT));
    CHECK(nomp_jit(&id_beta, BETA_SRC, none, 5, "beta", sizeof(double), NOMP_PTR, "rr", sizeof(double), NOMP_PTR, "rr_new",
                   sizeof(double), NOMP_PTR, "trace", sizeof(double), NOMP_PTR, "slot", sizeof(int), NOMP_INT));
    CHECK(nomp_jit(&id_dird, XPAY_DEV_SRC, none, 4, "p", sizeof(double), NOMP_PTR, "r", sizeof(double), NOMP_PTR, "beta",
                   sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT));
    CHECK(nomp_jit(&id_upd3, UPDATE_DEV3_SRC, red_rrn, 8, "x", sizeof(double), NOMP_PTR, "r", sizeof(double), NOMP_PTR, "p",
                   sizeof(double), NOMP_PTR, "w", sizeof(double), NOMP_PTR, "rr", sizeof(double), NOMP_PTR, "pap", sizeof(double),
                   NOMP_PTR, "N", sizeof(int), NOMP_INT, "rr_new", sizeof(double), NOMP_FLOAT));
    CHECK(nomp_jit(&id_dir3, XPAY_DEV3_SRC, none, 5, "p", sizeof(double), NOMP_PTR, "r", sizeof(double), NOMP_PTR, "rr_new",
                   sizeof(double), NOMP_PTR, "rr", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT));
    if (device_scalars == 4) { /* the direction update inside the operator kernel, beta read from device memory */
      CHECK(nomp_jit(&id_fused_dev, AX_XPAY_DOT_DEV_SRC, red_pap, 9, "w", sizeof(double), NOMP_PTR, "u", sizeof(double), NOMP_PTR,
                     "res", sizeof(double), NOMP_PTR, "g", sizeof(double), NOMP_PTR, "D", sizeof(double), NOMP_PTR, "beta",
                     sizeof(double), NOMP_PTR, "E", sizeof(int), NOMP_INT, "n", sizeof(int), NOMP_INT | NOMP_JIT, &n, "pap",
                     sizeof(double), NOMP_FLOAT));
      CHECK(nomp_jit(&id_beta2, BETA2_SRC, none, 3, "beta", sizeof(double), NOMP_PTR, "rr", sizeof(double), NOMP_PTR, "rr_new",
                     sizeof(double), NOMP_PTR));
      memset(p, 0, N * sizeof(double)); /* p <- r + 0 p in the first iteration */
      CHECK(nomp_update(p, 0, N, 8, NOMP_TO));
    }
    nomp_b200_device_reductions(1);
  }

  CHECK(nomp_sync());
  double t0 = now_s();
  int it = 0;
  if (device_scalars == 5) { /* two iterations recorded once, then replayed */
    int graph = -1;
    CHECK(nomp_run(id_axdot, w, p, g, D, &E, pap_d)); /* loads every kernel before the capture (lazy module loading) ... */
    CHECK(nomp_run(id_upd3, x, r, p, w, rr_d, pap_d, &Ni, rrn_d));
    CHECK(nomp_run(id_dir3, p, r, rrn_d, rr_d, &Ni));
    it = 1;                                           /* ... and is the first iteration */
    CHECK(nomp_b200_graph_begin());
    CHECK(nomp_run(id_axdot, w, p, g, D, &E, pap_d));
    CHECK(nomp_run(id_upd3, x, r, p, w, rrn_d, pap_d, &Ni, rr_d));
    CHECK(nomp_run(id_dir3, p, r, rr_d, rrn_d, &Ni));
    CHECK(nomp_run(id_axdot, w, p, g, D, &E, pap_d));
    CHECK(nomp_run(id_upd3, x, r, p, w, rr_d, pap_d, &Ni, rrn_d));
    CHECK(nomp_run(id_dir3, p, r, rrn_d, rr_d, &Ni));
    CHECK(nomp_b200_graph_end(&graph));
    CHECK(nomp_sync());
    t0 = now_s();
    const int every = check_every > 1 ? check_every / 2 : 1;
    for (int replays = 0; it + 2 <= max_iter && rr > tol * tol * rr0; it += 2) {
      CHECK(nomp_b200_graph_launch(graph));
      if (++replays % every == 0 || it + 4 > max_iter) { /* after an odd number of iterations the current residual is rr_new */
        CHECK(nomp_update(rrn_d, 0, 1, 8, NOMP_FROM));
        rr = rrn_d[0];
      }
    }
    CHECK(nomp_sync());
    CHECK(nomp_b200_graph_free(graph));
  }
  for (; device_scalars && device_scalars != 5 && it < max_iter && rr > tol * tol * rr0; it++) {
    if (it == 1) {
      CHECK(nomp_sync());
      t0 = now_s();
    }
    const int slot = it < 5 ? it : 5; /* iterations beyond the fifth share a scratch slot of the trace */
    if (device_scalars == 4) {
      CHECK(nomp_run(id_fused_dev, w, p, r, g, D, beta_d, &E, pap_d));
      CHECK(nomp_run(id_upd3, x, r, p, w, rr_d, pap_d, &Ni, rrn_d));
      CHECK(nomp_run(id_beta2, beta_d, rr_d, rrn_d));
    } else if (device_scalars == 3) {
      CHECK(nomp_run(id_axdot, w, p, g, D, &E, pap_d));
      CHECK(nomp_run(id_upd3, x, r, p, w, rr_d, pap_d, &Ni, rrn_d));
      CHECK(nomp_run(id_dir3, p, r, rrn_d, rr_d, &Ni));
      double *swap = rr_d; /* the new residual becomes the current one: a host-side change of names, no copy */
      rr_d = rrn_d, rrn_d = swap;
    } else {
      CHECK(nomp_run(id_axdot, w, p, g, D, &E, pap_d));
      CHECK(nomp_run(id_alpha, alpha_d, rr_d, pap_d, trace, &slot));
      CHECK(nomp_run(id_updd, x, r, p, w, alpha_d, &Ni, rrn_d));
      CHECK(nomp_run(id_beta, beta_d, rr_d, rrn_d, trace, &slot));
      CHECK(nomp_run(id_dird, p, r, beta_d, &Ni));
    }
    if ((it + 1) % check_every == 0 || it + 1 == max_iter) { /* the only host round trip */
      CHECK(nomp_update(rr_d, 0, 1, 8, NOMP_FROM));
      rr = rr_d[0];
    }
  }
  if (device_scalars) {
    nomp_b200_device_reductions(0);
    CHECK(nomp_update(trace, 0, 18, 8, NOMP_FROM));
    for (int i = 0; device_scalars == 1 && i < 5 && i < it; i++)
      printf("{\"iter\": %d, \"pAp\": %.17g, \"alpha\": %.17g, \"rr\": %.17g}\n", i, trace[3 * i], trace[3 * i + 1], trace[3 * i + 2]);
  }
  if (fused) { /* p <- r + beta p is the first thing the operator kernel does: start from p = 0, beta = 0 */
    int id_fused = -1;
    CHECK(nomp_jit(&id_fused, AX_XPAY_DOT_SRC, red_pap, 9, "w", sizeof(double), NOMP_PTR, "u", sizeof(double), NOMP_PTR, "res",
                   sizeof(double), NOMP_PTR, "g", sizeof(double), NOMP_PTR, "D", sizeof(double), NOMP_PTR, "beta", sizeof(double),
                   NOMP_FLOAT, "E", sizeof(int), NOMP_INT, "n", sizeof(int), NOMP_INT | NOMP_JIT, &n, "pap", sizeof(double),
                   NOMP_FLOAT));
    memset(p, 0, N * sizeof(double));
    CHECK(nomp_update(p, 0, N, 8, NOMP_TO));
    double beta = 0.0;
    CHECK(nomp_sync());
    t0 = now_s();
    for (; it < max_iter && rr > tol * tol * rr0; it++) {
      if (it == 1) {
        CHECK(nomp_sync());
        t0 = now_s();
      }
      CHECK(nomp_run(id_fused, w, p, r, g, D, &beta, &E, &pap));
      const double alpha = rr / pap;
      double rr_new = 0;
      CHECK(nomp_run(id_upd, x, r, p, w, &alpha, &Ni, &rr_new));
      beta = rr_new / rr;
      if (it < 5) printf("{\"iter\": %d, \"pAp\": %.17g, \"alpha\": %.17g, \"rr\": %.17g}\n", it, pap, alpha, rr_new);
      rr = rr_new;
    }
  }
  for (; !fused && !device_scalars && it < max_iter && rr > tol * tol * rr0; it++) {
    if (it == 1) { /* the first iteration loads every kernel (lazy module loading): time from the second one */
      CHECK(nomp_sync());
      t0 = now_s();
    }
    CHECK(nomp_run(id_axdot, w, p, g, D, &E, &pap));
    const double alpha = rr / pap;
    double rr_new = 0;
    CHECK(nomp_run(id_upd, x, r, p, w, &alpha, &Ni, &rr_new));
    const double beta = rr_new / rr;
    CHECK(nomp_run(id_dir, p, r, &beta, &Ni));
    if (it < 5) printf("{\"iter\": %d, \"pAp\": %.17g, \"alpha\": %.17g, \"rr\": %.17g}\n", it, pap, alpha, rr_new);
    rr = rr_new;
  }
  CHECK(nomp_sync());
  const double dt = now_s() - t0;

  /* true residual b - A x, and the error in the A-seminorm's range: ||A (x - x_true)|| */
  CHECK(nomp_run(id_ax, w, x, g, D, &E));        /* w = A x */
  CHECK(nomp_run(id_ax, p, xt, g, D, &E));       /* p = b   */
  double minus_one = -1.0, res2 = 0;
  const char *axpy_src = "void cg_axpy(double *y, const double *x, double a, int N) { for (int i = 0; i < N; i++) y[i] += a * x[i]; }\n";
  int id_axpy = -1;
  CHECK(nomp_jit(&id_axpy, axpy_src, none, 4, "y", sizeof(double), NOMP_PTR, "x", sizeof(double), NOMP_PTR, "a", sizeof(double),
                 NOMP_FLOAT, "N", sizeof(int), NOMP_INT));
  CHECK(nomp_run(id_axpy, p, w, &minus_one, &Ni)); /* p = b - A x */
  CHECK(nomp_run(id_dot, p, p, &Ni, &res2));
  printf("{\"iterations\": %d, \"rr_final\": %.17g, \"true_residual_rel\": %.3e, \"seconds\": %.6f, \"ms_per_iter\": %.4f, "
         "\"GDOF_per_s_per_rank\": %.2f, \"bytes_per_dof\": %d, \"scalars\": \"%s\"}\n",
         it, rr, sqrt(res2 / rr0), dt, dt / (it > 1 ? it - 1 : 1) * 1e3, it > 1 ? (double)N * (it - 1) / dt / 1e9 : 0.0, fused || device_scalars == 4 ? 128 : 136,
         device_scalars == 5 ? "graph" : device_scalars == 4 ? "device_fused" : device_scalars == 3 ? "device3" : device_scalars ? "device" : fused ? "fused" : "host");
  CHECK(nomp_finalize());
  return sqrt(res2 / rr0) < 1e-6 ? 0 : 2;
}
