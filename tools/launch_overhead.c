/* BASELINE.json configs[4]: nomp_jit cache-hit + nomp_run latency for n = 1e3 .. 1e8, alternating map and reduce
 * kernels, measured from C (no Python in the loop) with clock_gettime.  Prints one JSON object per line.
 *   build:  gcc -O2 -Iinclude tools/launch_overhead.c -o libnomp_b200/build/launch_overhead -Llibnomp_b200/lib -lnomp -Wl,-rpath,$PWD/libnomp_b200/lib
 *   run:    NOMP_INSTALL_DIR=$PWD/libnomp_b200 libnomp_b200/build/launch_overhead
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "nomp.h"

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

#define CHECK(x)                                                                                                  \
  do {                                                                                                            \
    int e_ = (x);                                                                                                 \
    if (e_) {                                                                                                     \
      char *s_ = nomp_get_err_str(e_);                                                                            \
      fprintf(stderr, "%s failed: %s\n", #x, s_ ? s_ : "?");                                                      \
      exit(1);                                                                                                    \
    }                                                                                                             \
  } while (0)

static const char *MAP_SRC = "void k_add(double *a, const double *b, int N) { for (int i = 0; i < N; i++) a[i] += b[i]; }";
static const char *GEN_SRC = "void k_gen(double *a, const double *b, int N) { for (int i = 0; i < N; i++) a[i] = a[i] * b[i] + i; }";
static const char *RED_SRC = "void k_dot(const double *a, const double *b, int N, double *s) { for (int i = 0; i < N; i++) s[0] += a[i] * b[i]; }";

int main(void) {
  const char *install = getenv("NOMP_INSTALL_DIR");
  const char *argv[] = {"launch_overhead", "--nomp-backend", "cuda", "--nomp-device", "0", "--nomp-platform", "0",
                        "--nomp-verbose", "1", "--nomp-install-dir", install ? install : "."};
  CHECK(nomp_init(11, argv));

  const size_t nmax = 100000000;
  double *a = calloc(nmax, sizeof(double)), *b = calloc(nmax, sizeof(double));
  for (size_t i = 0; i < nmax; i += 4096) a[i] = 1.0, b[i] = 2.0;
  CHECK(nomp_update(a, 0, nmax, sizeof(double), NOMP_TO));
  CHECK(nomp_update(b, 0, nmax, sizeof(double), NOMP_TO));

  const char *none[1] = {NULL};
  const char *red[4] = {"reduce", "s", "+", NULL};
  int id_map = -1, id_gen = -1, id_red = -1;
  double t0 = now_us();
  CHECK(nomp_jit(&id_map, MAP_SRC, none, 3, "a", sizeof(double), NOMP_PTR, "b", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT));
  double t_jit_native = now_us() - t0;
  t0 = now_us();
  CHECK(nomp_jit(&id_gen, GEN_SRC, none, 3, "a", sizeof(double), NOMP_PTR, "b", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT));
  double t_jit_nvrtc = now_us() - t0;
  CHECK(nomp_jit(&id_red, RED_SRC, red, 4, "a", sizeof(double), NOMP_PTR, "b", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT, "s",
                 sizeof(double), NOMP_FLOAT));
  printf("{\"what\": \"nomp_jit miss\", \"native_family_us\": %.1f, \"nvrtc_us\": %.1f}\n", t_jit_native, t_jit_nvrtc);

  /* cache hit: the call the generated code makes before every nomp_run */
  const int hits = 1000000;
  t0 = now_us();
  for (int i = 0; i < hits; i++)
    CHECK(nomp_jit(&id_map, MAP_SRC, none, 3, "a", sizeof(double), NOMP_PTR, "b", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT));
  printf("{\"what\": \"nomp_jit hit\", \"ns_per_call\": %.2f}\n", (now_us() - t0) * 1e3 / hits);

  for (long n = 1000; n <= (long)nmax; n *= 10) {
    int N = (int)n;
    int reps = n <= 1000000 ? 2000 : (n <= 10000000 ? 300 : 50);
    double s = 0;
    /* warm */
    for (int i = 0; i < 20; i++) CHECK(nomp_run(id_map, a, b, &N));
    CHECK(nomp_sync());
    /* (1) issue cost: back-to-back asynchronous nomp_run, one sync at the end */
    t0 = now_us();
    for (int i = 0; i < reps; i++) {
      CHECK(nomp_jit(&id_map, MAP_SRC, none, 3, "a", sizeof(double), NOMP_PTR, "b", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT));
      CHECK(nomp_run(id_map, a, b, &N));
    }
    double issue = (now_us() - t0) / reps;
    CHECK(nomp_sync());
    double pipelined = (now_us() - t0) / reps;
    /* (2) latency: nomp_run + nomp_sync every time */
    t0 = now_us();
    for (int i = 0; i < reps; i++) {
      CHECK(nomp_run(id_map, a, b, &N));
      CHECK(nomp_sync());
    }
    double map_sync = (now_us() - t0) / reps;
    /* (3) NVRTC-built elementwise kernel, same protocol */
    for (int i = 0; i < 5; i++) CHECK(nomp_run(id_gen, a, b, &N));
    CHECK(nomp_sync());
    t0 = now_us();
    for (int i = 0; i < reps; i++) {
      CHECK(nomp_run(id_gen, a, b, &N));
      CHECK(nomp_sync());
    }
    double gen_sync = (now_us() - t0) / reps;
    /* (4) reduce clause: the result is on the host when nomp_run returns */
    for (int i = 0; i < 5; i++) CHECK(nomp_run(id_red, a, b, &N, &s));
    t0 = now_us();
    for (int i = 0; i < reps; i++) CHECK(nomp_run(id_red, a, b, &N, &s));
    double red = (now_us() - t0) / reps;
    /* (5) alternating map / reduce, as in a solver loop */
    t0 = now_us();
    for (int i = 0; i < reps; i++) {
      CHECK(nomp_run(id_map, a, b, &N));
      CHECK(nomp_run(id_red, a, b, &N, &s));
    }
    double mixed = (now_us() - t0) / reps;
    printf("{\"n\": %ld, \"map_issue_us\": %.2f, \"map_pipelined_us\": %.2f, \"map_run_sync_us\": %.2f, \"nvrtc_map_run_sync_us\": %.2f, "
           "\"reduce_run_us\": %.2f, \"map_plus_reduce_us\": %.2f, \"map_GBs_pipelined\": %.1f, \"dot_GBs\": %.1f}\n",
           n, issue, pipelined, map_sync, gen_sync, red, mixed, n * 24.0 / pipelined / 1e3, n * 16.0 / red / 1e3);
    fflush(stdout);
  }
  CHECK(nomp_finalize());
  return 0;
}
