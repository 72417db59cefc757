/* nomp-aux.h -- small string / environment helpers exported by libnomp.so.  The reference's tests include this
 * header and link nomp_copy_env (reference tests/nomp-test.h:14, tests/nomp-api-021.c:6; prototypes at reference
 * include/nomp-aux.h:14-26). */
#ifndef LIBNOMP_B200_NOMP_AUX_H_
#define LIBNOMP_B200_NOMP_AUX_H_

#include <stddef.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Concatenate n strings (each read up to max_len bytes) into a fresh heap string. */
char *nomp_str_cat(unsigned n, unsigned max_len, ...);
/* Parse a non-negative decimal integer; -1 if `str` (up to `size` bytes) is anything else. */
int nomp_str_toui(const char *str, size_t size);
/* Largest of n ints. */
int nomp_max(unsigned n, ...);
/* Heap copy of getenv(name) truncated to `size` bytes, or NULL if unset. */
char *nomp_copy_env(const char *name, size_t size);
/* Length of realpath(path) through *len; logs NOMP_USER_INPUT_IS_INVALID if the path does not exist. */
int nomp_path_len(size_t *len, const char *path);

#ifdef __cplusplus
}
#endif

#endif /* LIBNOMP_B200_NOMP_AUX_H_ */
