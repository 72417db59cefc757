/* A complete spectral-element Poisson solve on the public libnomp API plus the gather-scatter extension:
 *
 *     -Laplace(u) = f  in [0,1]^3,  u = 0 on the boundary,  f = 3 pi^2 sin(pi x) sin(pi y) sin(pi z)
 *
 * on a box of ex * ey * ez hexahedral elements with n = N + 1 Gauss-Lobatto-Legendre points per direction, solved by
 * conjugate gradients on the assembled operator  M Q Q^T A_local  (A_local: the canonical Ax kernel; Q Q^T: direct
 * stiffness summation, nomp_b200_gs; M: Dirichlet mask).  Vectors are stored element by element (shared points are
 * stored once per element that touches them, all copies equal), so inner products weight every copy by
 * c = 1 / multiplicity.  The discrete solution is compared with the exact one, sin(pi x) sin(pi y) sin(pi z).
 *
 * Per iteration, four launches (five with more than one rank) and about 164 B/DOF:
 *     w = A p,  pAp = p . w              nompk_ax_dot_f64 (p is continuous, so the unweighted local sum is the global one)
 *     w = Q Q^T w                         nompk_gs_apply: interface values travel through NVLink peer memory
 *     x += a p;  r -= a M w;  rr = r.r|c  one launch: elementwise updates fused with the weighted reduce clause
 *     p = r + b p                         nompk_map (XPAY)
 * With NOMP_COMM_SIZE = k every rank owns ez / k layers of elements (ez must be a multiple of k); the dot products
 * are all-reduced by the runtime and the gather-scatter exchanges the interface planes.
 *
 *   usage: poisson_box [ex [ey [ez [n [max_iter [tol]]]]]]  + the usual --nomp-* flags   (one JSON object per line)
 */
#define _POSIX_C_SOURCE 200809L
#define _DEFAULT_SOURCE
#include "sem_common.h"

#include "nomp-b200.h"

static const char *UPDATE_SRC =
    "void pcg_update(double *x, double *r, const double *p, const double *w, const double *mask, const double *c,\n"
    "                double alpha, int N, double *rr) {\n"
    "  for (int i = 0; i < N; i++) {\n"
    "    x[i] += alpha * p[i];\n"
    "    r[i] -= alpha * (mask[i] * w[i]);\n"
    "    rr[0] += r[i] * r[i] * c[i];\n"
    "  }\n}\n";
static const char *XPAY_SRC =
    "void pcg_direction(double *p, const double *r, double beta, int N) { for (int i = 0; i < N; i++) p[i] = r[i] + beta * p[i]; }\n";
static const char *WDOT_SRC =
    "void pcg_wdot(const double *a, const double *b, const double *c, int N, double *s) {\n"
    "  for (int i = 0; i < N; i++) s[0] += a[i] * b[i] * c[i];\n}\n";
static const char *MUL_SRC = "void pcg_mask(double *a, const double *b, int N) { for (int i = 0; i < N; i++) a[i] *= b[i]; }\n";
static const char *ERR_SRC =
    "void pcg_err(const double *x, const double *u, int N, double *m) {\n"
    "  for (int i = 0; i < N; i++) { double d = (x[i] - u[i]) * (x[i] - u[i]); m[0] = (d > m[0]) ? d : m[0]; }\n}\n";

int main(int argc, const char **argv) {
  int dims[4] = {8, 8, 8, 8}, max_iter = 500, pos = 0;
  double tol = 1e-10;
  for (int i = 1; i < argc; i++) {
    if (!strncmp(argv[i], "--nomp", 6)) { i++; continue; }
    if (pos < 4) dims[pos] = atoi(argv[i]);
    else if (pos == 4) max_iter = atoi(argv[i]);
    else if (pos == 5) tol = atof(argv[i]);
    pos++;
  }
  const int ex = dims[0], ey = dims[1], ez = dims[2], n = dims[3], Np = n - 1;
  CHECK(nomp_init(argc, argv));
  const int rank = nomp_b200_comm_rank(), world = nomp_b200_comm_size();
  if (ex < 1 || ey < 1 || ez < 1 || ez % world || n < 2 || n > 16) {
    fprintf(stderr, "need ex, ey, ez >= 1, ez a multiple of the %d ranks, 2 <= n <= 16\n", world);
    return 1;
  }
  const int ezl = ez / world, z0 = rank * ezl;      /* this rank's layers of elements */
  int E = ex * ey * ezl;
  const size_t n3 = (size_t)n * n * n, N = (size_t)E * n3;
  const int Ni = (int)N;
  const double hx = 1.0 / ex, hy = 1.0 / ey, hz = 1.0 / ez, J = hx * hy * hz / 8.0;
  const long long px = (long long)Np * ex + 1, py = (long long)Np * ey + 1, pz = (long long)Np * ez + 1;

  double xi[16], wt[16], *D = calloc((size_t)n * n, 8);
  gll(n, xi, wt, D);
  double *x = calloc(N, 8), *r = calloc(N, 8), *p = calloc(N, 8), *w = calloc(N, 8), *ue = calloc(N, 8);
  double *mask = calloc(N, 8), *c = calloc(N, 8), *g = calloc(6 * N, 8);
  long long *ids = calloc(N, sizeof(long long));
  for (int e = 0; e < E; e++) {
    const int e_x = e % ex, e_y = (e / ex) % ey, e_z = z0 + e / (ex * ey);
    for (int k = 0; k < n; k++)
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
          const size_t q = (size_t)e * n3 + ((size_t)k * n + j) * n + i;
          const long long gx = (long long)e_x * Np + i, gy = (long long)e_y * Np + j, gz = (long long)e_z * Np + k;
          const double X = (e_x + 0.5 * (xi[i] + 1.0)) * hx, Y = (e_y + 0.5 * (xi[j] + 1.0)) * hy,
                       Z = (e_z + 0.5 * (xi[k] + 1.0)) * hz;
          const double B = wt[i] * wt[j] * wt[k] * J;
          ids[q] = 1 + gx + px * (gy + py * gz);
          mask[q] = (gx == 0 || gx == px - 1 || gy == 0 || gy == py - 1 || gz == 0 || gz == pz - 1) ? 0.0 : 1.0;
          ue[q] = sin(M_PI * X) * sin(M_PI * Y) * sin(M_PI * Z);
          r[q] = B * 3.0 * M_PI * M_PI * ue[q];     /* local load vector B f */
          c[q] = 1.0;
          const size_t gq = (size_t)e * 6 * n3 + ((size_t)k * n + j) * n + i;
          g[gq + 0 * n3] = B * 4.0 / (hx * hx), g[gq + 3 * n3] = B * 4.0 / (hy * hy), g[gq + 5 * n3] = B * 4.0 / (hz * hz);
        }
  }

  double *vectors[] = {x, r, p, w, ue, mask, c};
  for (int a = 0; a < 7; a++) CHECK(nomp_update(vectors[a], 0, N, 8, NOMP_TO));
  CHECK(nomp_update(g, 0, 6 * N, 8, NOMP_TO));
  CHECK(nomp_update(D, 0, (size_t)n * n, 8, NOMP_TO));
  int gs = -1;
  const double t_setup = now_s();
  CHECK(nomp_b200_gs_setup(&gs, ids, N));
  const double setup_s = now_s() - t_setup;
  size_t info[8];
  CHECK(nomp_b200_gs_info(gs, info));

  const char *none[1] = {NULL};
  const char *red_pap[4] = {"reduce", "pap", "+", NULL}, *red_rr[4] = {"reduce", "rr", "+", NULL},
             *red_s[4] = {"reduce", "s", "+", NULL}, *red_m[4] = {"reduce", "m", "max", NULL};
  int id_axdot = -1, id_upd = -1, id_dir = -1, id_wdot = -1, id_mul = -1, id_err = -1;
  CHECK(nomp_jit(&id_axdot, AX_DOT_SRC, red_pap, 7, "w", sizeof(double), NOMP_PTR, "u", sizeof(double), NOMP_PTR, "g", sizeof(double),
                 NOMP_PTR, "D", sizeof(double), NOMP_PTR, "E", sizeof(int), NOMP_INT, "n", sizeof(int), NOMP_INT | NOMP_JIT, &n, "pap",
                 sizeof(double), NOMP_FLOAT));
  CHECK(nomp_jit(&id_upd, UPDATE_SRC, red_rr, 9, "x", sizeof(double), NOMP_PTR, "r", sizeof(double), NOMP_PTR, "p", sizeof(double),
                 NOMP_PTR, "w", sizeof(double), NOMP_PTR, "mask", sizeof(double), NOMP_PTR, "c", sizeof(double), NOMP_PTR, "alpha",
                 sizeof(double), NOMP_FLOAT, "N", sizeof(int), NOMP_INT, "rr", sizeof(double), NOMP_FLOAT));
  CHECK(nomp_jit(&id_dir, XPAY_SRC, none, 4, "p", sizeof(double), NOMP_PTR, "r", sizeof(double), NOMP_PTR, "beta", sizeof(double),
                 NOMP_FLOAT, "N", sizeof(int), NOMP_INT));
  CHECK(nomp_jit(&id_wdot, WDOT_SRC, red_s, 5, "a", sizeof(double), NOMP_PTR, "b", sizeof(double), NOMP_PTR, "c", sizeof(double),
                 NOMP_PTR, "N", sizeof(int), NOMP_INT, "s", sizeof(double), NOMP_FLOAT));
  CHECK(nomp_jit(&id_mul, MUL_SRC, none, 3, "a", sizeof(double), NOMP_PTR, "b", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT));
  CHECK(nomp_jit(&id_err, ERR_SRC, red_m, 4, "x", sizeof(double), NOMP_PTR, "u", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT,
                 "m", sizeof(double), NOMP_FLOAT));

  /* c = 1 / multiplicity (gather-scatter of ones, inverted on the host once); b = M Q Q^T B f; x0 = 0, r = p = b */
  CHECK(nomp_b200_gs(gs, c, 8, NOMP_FLOAT, "+"));
  CHECK(nomp_update(c, 0, N, 8, NOMP_FROM));
  for (size_t i = 0; i < N; i++) c[i] = 1.0 / c[i];
  CHECK(nomp_update(c, 0, N, 8, NOMP_TO));
  CHECK(nomp_b200_gs(gs, r, 8, NOMP_FLOAT, "+"));
  CHECK(nomp_run(id_mul, r, mask, &Ni));
  double zero = 0.0, rr = 0, rr0, pap = 0;
  CHECK(nomp_run(id_dir, p, r, &zero, &Ni));
  CHECK(nomp_run(id_wdot, r, r, c, &Ni, &rr));
  rr0 = rr;
  printf("{\"rank\": %d, \"ranks\": %d, \"elements\": [%d, %d, %d], \"n\": %d, \"dof_per_rank\": %zu, \"unique_points\": %lld, "
         "\"gs_groups\": %zu, \"gs_copies\": %zu, \"gs_shared_with_ranks\": %zu, \"gs_setup_s\": %.4f, \"rr0\": %.17g}\n",
         rank, world, ex, ey, ez, n, N, px * py * pz, info[2], info[3], info[7], setup_s, rr0);

  CHECK(nomp_sync());
  double t0 = now_s();
  int it = 0;
  for (; it < max_iter && rr > tol * tol * rr0; it++) {
    if (it == 1) { /* the first iteration loads every kernel (lazy module loading): time from the second one */
      CHECK(nomp_sync());
      t0 = now_s();
    }
    CHECK(nomp_run(id_axdot, w, p, g, D, &E, &pap));
    CHECK(nomp_b200_gs(gs, w, 8, NOMP_FLOAT, "+"));
    const double alpha = rr / pap;
    double rr_new = 0;
    CHECK(nomp_run(id_upd, x, r, p, w, mask, c, &alpha, &Ni, &rr_new));
    const double beta = rr_new / rr;
    CHECK(nomp_run(id_dir, p, r, &beta, &Ni));
    if (it < 5) printf("{\"iter\": %d, \"pAp\": %.17g, \"alpha\": %.17g, \"rr\": %.17g}\n", it, pap, alpha, rr_new);
    rr = rr_new;
  }
  CHECK(nomp_sync());
  const double dt = now_s() - t0;

  /* POISSON_PHASES=1: ten more iterations with a nomp_sync() after every call, to see where an iteration's time goes
   * (the solution is left alone: alpha = 0 for these) */
  if (getenv("POISSON_PHASES")) {
    double t[5] = {0, 0, 0, 0, 0}, none_alpha = 0.0, rr_tmp = 0, one = 1.0;
    for (int rep = 0; rep < 10; rep++) {
      double a = now_s();
      CHECK(nomp_run(id_axdot, w, p, g, D, &E, &pap));
      CHECK(nomp_sync());
      double b = now_s();
      CHECK(nomp_b200_gs(gs, w, 8, NOMP_FLOAT, "+"));
      CHECK(nomp_sync());
      double c2 = now_s();
      CHECK(nomp_run(id_upd, x, r, p, w, mask, c, &none_alpha, &Ni, &rr_tmp));
      CHECK(nomp_sync());
      double d = now_s();
      CHECK(nomp_run(id_dir, p, p, &zero, &Ni));
      CHECK(nomp_sync());
      double e2 = now_s();
      t[0] += b - a, t[1] += c2 - b, t[2] += d - c2, t[3] += e2 - d;
    }
    (void)one;
    printf("{\"phase_ms\": {\"ax_dot\": %.4f, \"gather_scatter\": %.4f, \"update_rr\": %.4f, \"direction\": %.4f}}\n",
           t[0] * 100, t[1] * 100, t[2] * 100, t[3] * 100);
  }

  double err2 = 0;
  CHECK(nomp_run(id_err, x, ue, &Ni, &err2));
  printf("{\"iterations\": %d, \"rr_final\": %.17g, \"residual_rel\": %.3e, \"max_error\": %.3e, \"seconds\": %.6f, "
         "\"ms_per_iter\": %.4f, \"GDOF_per_s_per_rank\": %.2f}\n",
         it, rr, sqrt(rr / rr0), sqrt(err2), dt, dt / (it > 1 ? it - 1 : 1) * 1e3, it > 1 ? (double)N * (it - 1) / dt / 1e9 : 0.0);
  CHECK(nomp_b200_gs_free(gs));
  CHECK(nomp_finalize());
  return 0;
}
