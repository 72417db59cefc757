/* String / environment helpers (behaviour of reference src/aux.c:26-146; exported because the reference tests link
 * nomp_copy_env, reference tests/nomp-api-021.c:6). */
#include <ctype.h>
#include <errno.h>

#include "nomp-aux.h"
#include "nomp-impl.h"

NOMP_EXPORT char *nomp_str_cat(unsigned n, unsigned max_len, ...) {
  va_list ap;
  size_t total = 0;
  va_start(ap, max_len);
  for (unsigned i = 0; i < n; i++) total += strnlen(va_arg(ap, const char *), max_len);
  va_end(ap);

  char *out = nomp_calloc(char, total + 1);
  size_t pos = 0;
  va_start(ap, max_len);
  for (unsigned i = 0; i < n; i++) {
    const char *s = va_arg(ap, const char *);
    size_t len = strnlen(s, max_len);
    memcpy(out + pos, s, len);
    pos += len;
  }
  va_end(ap);
  out[pos] = '\0';
  return out;
}

NOMP_EXPORT int nomp_str_toui(const char *str, size_t size) {
  if (str == NULL || size == 0) return -1;
  long value = 0;
  size_t i = 0;
  for (; i < size && str[i] != '\0'; i++) {
    if (!isdigit((unsigned char)str[i])) return -1;
    value = value * 10 + (str[i] - '0');
    if (value > INT_MAX) return -1;
  }
  return i == 0 ? -1 : (int)value;
}

NOMP_EXPORT int nomp_max(unsigned n, ...) {
  va_list ap;
  va_start(ap, n);
  int best = INT_MIN;
  for (unsigned i = 0; i < n; i++) {
    int v = va_arg(ap, int);
    if (v > best) best = v;
  }
  va_end(ap);
  return best;
}

NOMP_EXPORT char *nomp_copy_env(const char *name, size_t size) {
  const char *v = getenv(name);
  return v ? strndup(v, size) : NULL;
}

NOMP_EXPORT int nomp_path_len(size_t *len, const char *path) {
  if (len) *len = 0;
  char *abs = realpath(path, NULL);
  if (abs == NULL) {
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Unable to find path: \"%s\". Error: %s.", path,
                    strerror(errno));
  }
  if (len) *len = strnlen(abs, PATH_MAX);
  free(abs);
  return 0;
}
