/* TEST PROGRAM (tests/test_hostdev_cpu.py runs it on the CUDA test double with the sanitized runtime): the graph
 * extension of include/nomp-b200.h from C -- capture two launches (a native map and a generated reduction whose result
 * stays in device memory), replay them, check what a capturing stream refuses. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nomp-b200.h"
#include "nomp.h"

#define CHECK(x)                                                                                                       \
  do {                                                                                                                 \
    int err_ = (x);                                                                                                    \
    if (err_) {                                                                                                        \
      char *s_ = nomp_get_err_str(err_);                                                                               \
      fprintf(stderr, "%s failed: %s\n", #x, s_ ? s_ : "?");                                                           \
      return 1;                                                                                                        \
    }                                                                                                                  \
  } while (0)
#define EXPECT(cond)                                                                                                   \
  do {                                                                                                                 \
    if (!(cond)) {                                                                                                     \
      fprintf(stderr, "line %d: %s is false\n", __LINE__, #cond);                                                      \
      return 2;                                                                                                        \
    }                                                                                                                  \
  } while (0)

int main(int argc, const char **argv) {
  CHECK(nomp_init(argc, argv));
  enum { N = 1000 };
  double *a = calloc(N, 8), *b = calloc(N, 8), *s = calloc(1, 8);
  for (int i = 0; i < N; i++) a[i] = 1.0, b[i] = (double)(i % 7);
  CHECK(nomp_update(a, 0, N, 8, NOMP_TO));
  CHECK(nomp_update(b, 0, N, 8, NOMP_TO));
  CHECK(nomp_update(s, 0, 1, 8, NOMP_TO));
  const char *none[1] = {NULL}, *red[4] = {"reduce", "s", "+", NULL};
  int id_add = -1, id_sum = -1, n = N;
  CHECK(nomp_jit(&id_add, "void add(double *a, const double *b, int N) { for (int i = 0; i < N; i++) a[i] += b[i]; }", none, 3, "a",
                 sizeof(double), NOMP_PTR, "b", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT));
  CHECK(nomp_jit(&id_sum, "void sum(const double *a, int N, double *s) { for (int i = 0; i < N; i++) if (a[i] > 0) s[0] += a[i]; }", red,
                 3, "a", sizeof(double), NOMP_PTR, "N", sizeof(int), NOMP_INT, "s", sizeof(double), NOMP_FLOAT));
  nomp_b200_device_reductions(1);
  CHECK(nomp_run(id_add, a, b, &n)); /* once outside the capture: kernels load on their first launch */
  CHECK(nomp_run(id_sum, a, &n, s));

  int graph = -1;
  EXPECT(nomp_b200_graph_end(&graph) > 0); /* not capturing */
  CHECK(nomp_b200_graph_begin());
  EXPECT(nomp_b200_graph_begin() > 0);
  CHECK(nomp_run(id_add, a, b, &n));
  CHECK(nomp_run(id_sum, a, &n, s));
  EXPECT(nomp_sync() > 0);
  EXPECT(nomp_update(a, 0, N, 8, NOMP_FROM) > 0);
  double host_scalar = 0;
  EXPECT(nomp_run(id_sum, a, &n, &host_scalar) > 0); /* would have to wait for a stream that only records */
  CHECK(nomp_b200_graph_end(&graph));
  EXPECT(graph >= 0);
  for (int rep = 0; rep < 3; rep++) CHECK(nomp_b200_graph_launch(graph));
  CHECK(nomp_sync());
  nomp_b200_device_reductions(0);
  CHECK(nomp_update(a, 0, N, 8, NOMP_FROM));
  CHECK(nomp_update(s, 0, 1, 8, NOMP_FROM));
  double want = 0;
  for (int i = 0; i < N; i++) { /* one add outside the capture, none during it, three replays */
    EXPECT(a[i] == 1.0 + 4.0 * (double)(i % 7));
    want += a[i];
  }
  EXPECT(s[0] == want);
  CHECK(nomp_b200_graph_free(graph));
  EXPECT(nomp_b200_graph_launch(graph) > 0);
  EXPECT(nomp_b200_graph_free(graph) > 0);
  CHECK(nomp_b200_graph_begin()); /* a capture that is still open at finalize is simply dropped with the stream */
  CHECK(nomp_b200_graph_end(&graph));
  CHECK(nomp_update(a, 0, N, 8, NOMP_FREE));
  CHECK(nomp_update(b, 0, N, 8, NOMP_FREE));
  CHECK(nomp_update(s, 0, 1, 8, NOMP_FREE));
  CHECK(nomp_finalize());
  free(a), free(b), free(s);
  printf("graph smoke: ok\n");
  return 0;
}
