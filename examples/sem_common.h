/* Pieces shared by the example programs: error checking, a timer, Gauss-Lobatto-Legendre nodes / weights / derivative
 * matrix, and the canonical Ax kernel strings (identical to nomp_bridge.families.AX_KERNEL_SOURCE /
 * AX_DOT_KERNEL_SOURCE, which the transform bridge maps onto libnompk's hand-written kernels). */
#ifndef LIBNOMP_B200_EXAMPLES_SEM_COMMON_H_
#define LIBNOMP_B200_EXAMPLES_SEM_COMMON_H_

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "nomp.h"

#define CHECK(x)                                                                                                  \
  do {                                                                                                            \
    int e_ = (x);                                                                                                 \
    if (e_) {                                                                                                     \
      char *s_ = nomp_get_err_str(e_);                                                                            \
      fprintf(stderr, "%s failed: %s\n", #x, s_ ? s_ : "?");                                                      \
      exit(1);                                                                                                    \
    }                                                                                                             \
  } while (0)

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* Gauss-Lobatto-Legendre nodes x, weights wt (either may be NULL) and derivative matrix D[a][l] = l_l'(x_a),
 * row-major */
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static double legendre(int N, double x, double *dP) {
  double p0 = 1.0, p1 = x;
  for (int k = 2; k <= N; k++) {
    double pk = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
    p0 = p1, p1 = pk;
  }
  if (dP) *dP = N * (x * p1 - p0) / (x * x - 1.0);
  return N == 0 ? 1.0 : p1;
}

static void gll(int n, double *nodes, double *wt, double *D) {
  const int N = n - 1;
  double x[32];
  x[0] = -1.0, x[N] = 1.0;
  for (int i = 1; i < N; i++) {
    double xi = -cos(M_PI * i / N);
    for (int it = 0; it < 100; it++) {
      double dP, P = legendre(N, xi, &dP);
      double d2P = (2.0 * xi * dP - N * (N + 1.0) * P) / (1.0 - xi * xi);
      double dx = dP / d2P;
      xi -= dx;
      if (fabs(dx) < 1e-16) break;
    }
    x[i] = xi;
  }
  for (int a = 0; a < n; a++)
    for (int l = 0; l < n; l++) {
      if (a != l) D[a * n + l] = legendre(N, x[a], NULL) / (legendre(N, x[l], NULL) * (x[a] - x[l]));
      else D[a * n + l] = a == 0 ? -N * (N + 1.0) / 4.0 : (a == N ? N * (N + 1.0) / 4.0 : 0.0);
    }
  for (int a = 0; a < n; a++) {
    const double P = legendre(N, x[a], NULL);
    if (nodes) nodes[a] = x[a];
    if (wt) wt[a] = 2.0 / (N * (N + 1.0) * P * P);
  }
}

static void gll_derivative(int n, double *D) { gll(n, NULL, NULL, D); }

/* The canonical kernel strings (identical to nomp_bridge.families.AX_KERNEL_SOURCE / AX_DOT_KERNEL_SOURCE). */
#define AX_BODY(EXTRA) AX_BODY2("", EXTRA)
#define AX_BODY2(PRE, EXTRA)                                                                                      \
  "  for (int e = 0; e < E; e++) {\n" PRE                                                                         \
  "    double ur[n][n][n];\n    double us[n][n][n];\n    double ut[n][n][n];\n"                                   \
  "    for (int k = 0; k < n; k++)\n      for (int j = 0; j < n; j++)\n        for (int i = 0; i < n; i++) {\n"   \
  "          double r = 0;\n          double s = 0;\n          double t = 0;\n"                                   \
  "          for (int l = 0; l < n; l++) {\n"                                                                     \
  "            r += D[i * n + l] * u[e * n * n * n + k * n * n + j * n + l];\n"                                   \
  "            s += D[j * n + l] * u[e * n * n * n + k * n * n + l * n + i];\n"                                   \
  "            t += D[k * n + l] * u[e * n * n * n + l * n * n + j * n + i];\n"                                   \
  "          }\n"                                                                                                 \
  "          ur[k][j][i] = g[(e * 6 + 0) * n * n * n + k * n * n + j * n + i] * r + g[(e * 6 + 1) * n * n * n + k * n * n + j * n + i] * s + g[(e * 6 + 2) * n * n * n + k * n * n + j * n + i] * t;\n" \
  "          us[k][j][i] = g[(e * 6 + 1) * n * n * n + k * n * n + j * n + i] * r + g[(e * 6 + 3) * n * n * n + k * n * n + j * n + i] * s + g[(e * 6 + 4) * n * n * n + k * n * n + j * n + i] * t;\n" \
  "          ut[k][j][i] = g[(e * 6 + 2) * n * n * n + k * n * n + j * n + i] * r + g[(e * 6 + 4) * n * n * n + k * n * n + j * n + i] * s + g[(e * 6 + 5) * n * n * n + k * n * n + j * n + i] * t;\n" \
  "        }\n"                                                                                                   \
  "    for (int k = 0; k < n; k++)\n      for (int j = 0; j < n; j++)\n        for (int i = 0; i < n; i++) {\n"   \
  "          double acc = 0;\n          for (int l = 0; l < n; l++) {\n"                                          \
  "            acc += D[l * n + i] * ur[k][j][l];\n            acc += D[l * n + j] * us[k][l][i];\n"              \
  "            acc += D[l * n + k] * ut[l][j][i];\n          }\n"                                                 \
  "          w[e * n * n * n + k * n * n + j * n + i] = acc;\n" EXTRA "        }\n  }\n}\n"

static const char *AX_SRC =
    "void nomp_ax(double *w, const double *u, const double *g, const double *D, int E, int n) {\n" AX_BODY("");
static const char *AX_DOT_SRC =
    "void nomp_ax_dot(double *w, const double *u, const double *g, const double *D, int E, int n, double *pap) {\n" AX_BODY(
        "          pap[0] += u[e * n * n * n + k * n * n + j * n + i] * acc;\n");

/* ... and nomp_bridge.families.AX_XPAY_DOT_KERNEL_SOURCE up to the names of its identifiers: u <- res + beta u first */
#define AX_POINT "e * n * n * n + k * n * n + j * n + i"
static const char *AX_XPAY_DOT_SRC =
    "void nomp_ax_xpay_dot(double *w, double *u, const double *res, const double *g, const double *D, double beta, int E, int n,"
    " double *pap) {\n" AX_BODY2(
        "    for (int k = 0; k < n; k++)\n      for (int j = 0; j < n; j++)\n        for (int i = 0; i < n; i++)\n"
        "          u[" AX_POINT "] = res[" AX_POINT "] + beta * u[" AX_POINT "];\n",
        "          pap[0] += u[e * n * n * n + k * n * n + j * n + i] * acc;\n");

/* ... and AX_XPAY_DOT_DEV_KERNEL_SOURCE: beta is a scalar in device memory, read as beta[0] */
static const char *AX_XPAY_DOT_DEV_SRC =
    "void nomp_ax_xpay_dot(double *w, double *u, const double *res, const double *g, const double *D, const double *beta, int E,"
    " int n, double *pap) {\n" AX_BODY2(
        "    for (int k = 0; k < n; k++)\n      for (int j = 0; j < n; j++)\n        for (int i = 0; i < n; i++)\n"
        "          u[" AX_POINT "] = res[" AX_POINT "] + beta[0] * u[" AX_POINT "];\n",
        "          pap[0] += u[e * n * n * n + k * n * n + j * n + i] * acc;\n");

#endif
