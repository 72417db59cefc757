/* nomp-mem.h -- checked allocation macros (header only).  The reference's tests use nomp_calloc() and nomp_free()
 * through this header (reference tests/nomp-test.h:15, :80; macros at reference include/nomp-mem.h:27-75):
 *   T *p = nomp_calloc(T, count);   p = nomp_realloc(p, T, count);   nomp_free(&p);   // frees and nulls p
 * Allocation failure is fatal, as in the reference. */
#ifndef LIBNOMP_B200_NOMP_MEM_H_
#define LIBNOMP_B200_NOMP_MEM_H_

#include <stdio.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

static inline void *nomp_mem_checked_(void *p, size_t bytes, const char *what, const char *file, unsigned line) {
  if (p == NULL && bytes != 0) {
    fprintf(stderr, "[Error] %s:%u: nomp_%s of %zu bytes failed\n", file, line, what, bytes);
    exit(EXIT_FAILURE);
  }
  return p;
}

static inline void nomp_free_(void **p) {
  if (p != NULL) {
    free(*p);
    *p = NULL;
  }
}

#define nomp_free(p) nomp_free_((void **)(p))
#define nomp_calloc(T, count)                                                                                    \
  ((T *)nomp_mem_checked_(calloc((size_t)(count) ? (size_t)(count) : 1, sizeof(T)), sizeof(T), "calloc",         \
                          __FILE__, __LINE__))
#define nomp_realloc(ptr, T, count)                                                                              \
  ((T *)nomp_mem_checked_(realloc((ptr), ((size_t)(count) ? (size_t)(count) : 1) * sizeof(T)), sizeof(T),        \
                          "realloc", __FILE__, __LINE__))

#ifdef __cplusplus
}
#endif

#endif /* LIBNOMP_B200_NOMP_MEM_H_ */
