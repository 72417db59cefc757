/* On-disk JIT cache.
 *
 * The reference pays the whole cold path on every process start: libclang parse, loopy transforms and code
 * generation, NVRTC, driver JIT (reference src/nomp.c:529-570, backends/unified-cuda-hip-impl.h:96-141); the only
 * thing it hashes is the source, to name libclang's temporary file (reference python/loopy_api.py:771).  Here two
 * kinds of entries are kept under one directory, both named by a SHA-256 of everything that can change them:
 *
 *   <hex>.knl    what the transform bridge produced for one nomp_jit() call: kernel name, launch-size expressions and
 *                the generated source with its "//!nomp" descriptor.  Key: kernel string, clauses, argument list,
 *                NOMP_JIT values, backend context (arch, max threads), the text of every transform / annotation script
 *                named by the clauses, and the text of the bridge itself (nomp.c:build_program).
 *   <hex>.cubin  NVRTC output for one generated source.  Key: source, options, NVRTC version (backends/cuda.c).
 *
 * Directory: $NOMP_JIT_CACHE_DIR, else $XDG_CACHE_HOME/libnomp_b200, else $HOME/.cache/libnomp_b200.  NOMP_JIT_CACHE=0
 * turns the cache off.  Entries are written to a temporary name and rename()d, so concurrent processes (one per GPU)
 * never see a partial file, and carry a SHA-256 trailer of their content: a corrupt or truncated entry is a miss and
 * is overwritten.
 */
#include <dirent.h>
#include <errno.h>
#include <sys/stat.h>
#include <unistd.h>

#include "nomp-impl.h"

/* ---- SHA-256 (FIPS 180-4) ------------------------------------------------------------------------------------- */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
    0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
    0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
    0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
    0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
    0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};

static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void sha_block(nomp_sha256_t *c, const unsigned char *p) {
  uint32_t w[64], s[8];
  for (int i = 0; i < 16; i++)
    w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
    const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  memcpy(s, c->h, sizeof(s));
  for (int i = 0; i < 64; i++) {
    const uint32_t S1 = rotr(s[4], 6) ^ rotr(s[4], 11) ^ rotr(s[4], 25), ch = (s[4] & s[5]) ^ (~s[4] & s[6]);
    const uint32_t t1 = s[7] + S1 + ch + K256[i] + w[i];
    const uint32_t S0 = rotr(s[0], 2) ^ rotr(s[0], 13) ^ rotr(s[0], 22);
    const uint32_t maj = (s[0] & s[1]) ^ (s[0] & s[2]) ^ (s[1] & s[2]);
    s[7] = s[6], s[6] = s[5], s[5] = s[4], s[4] = s[3] + t1, s[3] = s[2], s[2] = s[1], s[1] = s[0], s[0] = t1 + S0 + maj;
  }
  for (int i = 0; i < 8; i++) c->h[i] += s[i];
}

void nomp_sha256_init(nomp_sha256_t *c) {
  static const uint32_t h0[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  memcpy(c->h, h0, sizeof(h0));
  c->len = 0, c->fill = 0;
}

void nomp_sha256_update(nomp_sha256_t *c, const void *data, size_t n) {
  const unsigned char *p = (const unsigned char *)data;
  c->len += n;
  while (n > 0) {
    size_t take = 64 - c->fill < n ? 64 - c->fill : n;
    memcpy(c->buf + c->fill, p, take);
    c->fill += (unsigned)take, p += take, n -= take;
    if (c->fill == 64) sha_block(c, c->buf), c->fill = 0;
  }
}

/* One field of a key: length-prefixed, so ("ab","c") and ("a","bc") hash differently. */
void nomp_sha256_field(nomp_sha256_t *c, const char *s) {
  const uint64_t n = s ? strlen(s) : (uint64_t)-1;
  nomp_sha256_update(c, &n, sizeof(n));
  if (s) nomp_sha256_update(c, s, (size_t)n);
}

void nomp_sha256_hex(nomp_sha256_t *c, char hex[65]) {
  const uint64_t bits = c->len * 8;
  unsigned char pad[72] = {0x80};
  const size_t padlen = (c->fill < 56 ? 56 : 120) - c->fill;
  for (int i = 0; i < 8; i++) pad[padlen + i] = (unsigned char)(bits >> (56 - 8 * i));
  nomp_sha256_update(c, pad, padlen + 8);
  for (int i = 0; i < 8; i++) snprintf(hex + 8 * i, 9, "%08x", c->h[i]);
}

NOMP_EXPORT void nomp_b200_sha256_hex(const void *data, size_t n, char hex[65]) {
  nomp_sha256_t c;
  nomp_sha256_init(&c);
  nomp_sha256_update(&c, data, n);
  nomp_sha256_hex(&c, hex);
}

/* ---- directory ------------------------------------------------------------------------------------------------ */
static char cache_dir[PATH_MAX + 1];
static int cache_state = 0; /* 0 = not looked at yet, 1 = usable, -1 = off */
static unsigned long long stats[4];

static int mkdir_p(const char *path) {
  char tmp[PATH_MAX + 1];
  snprintf(tmp, sizeof(tmp), "%s", path);
  for (char *p = tmp + 1; *p; p++) {
    if (*p != '/') continue;
    *p = '\0';
    if (mkdir(tmp, 0777) && errno != EEXIST) return 1; /* parents: the user's umask decides, as for ~/.cache itself */
    *p = '/';
  }
  return mkdir(tmp, 0700) && errno != EEXIST; /* the cache directory itself: private */
}

/* Entries are loaded through cuModuleLoadData and run: their checksum protects against damage, not against somebody
 * else writing a valid entry.  So the directory must belong to this user and be writable by nobody else. */
static int private_dir(const char *path) {
  struct stat st;
  if (stat(path, &st) || !S_ISDIR(st.st_mode)) return 0;
  return st.st_uid == geteuid() && (st.st_mode & (S_IWGRP | S_IWOTH)) == 0;
}

const char *nomp_jit_cache_dir(void) {
  if (cache_state == 0) {
    cache_state = -1;
    const char *off = getenv("NOMP_JIT_CACHE"), *dir = getenv("NOMP_JIT_CACHE_DIR");
    const char *xdg = getenv("XDG_CACHE_HOME"), *home = getenv("HOME");
    if (off && !strcmp(off, "0")) return NULL;
    if (dir && dir[0]) snprintf(cache_dir, sizeof(cache_dir), "%s", dir);
    else if (xdg && xdg[0]) snprintf(cache_dir, sizeof(cache_dir), "%s/libnomp_b200", xdg);
    else if (home && home[0]) snprintf(cache_dir, sizeof(cache_dir), "%s/.cache/libnomp_b200", home);
    else return NULL;
    if (mkdir_p(cache_dir) || access(cache_dir, W_OK)) return NULL;
    if (!private_dir(cache_dir)) {
      nomp_log(0, NOMP_WARNING, "JIT cache directory \"%s\" is not owned by this user or is writable by others: cache off.", cache_dir);
      return NULL;
    }
    cache_state = 1;
  }
  return cache_state == 1 ? cache_dir : NULL;
}

/* Re-read the environment on the next use (nomp_init calls this: tests switch directories between runs). */
void nomp_jit_cache_reset(void) { cache_state = 0; }

void nomp_jit_cache_count(int which) { stats[which & 3]++; }

NOMP_EXPORT void nomp_b200_jit_cache_stats(unsigned long long out[4]) { memcpy(out, stats, sizeof(stats)); }

/* ---- entries -------------------------------------------------------------------------------------------------- */
/* Every entry ends with the SHA-256 (64 hex characters) of what precedes it.  A truncated or damaged file -- a crashed
 * writer, a full disk -- must never reach the CUDA driver: cuModuleLoadData trusts the ELF headers of the image it is
 * given and reads wherever they point. */
#define TRAILER 64

static void digest_of(const void *data, size_t n, char hex[65]) {
  nomp_sha256_t c;
  nomp_sha256_init(&c);
  nomp_sha256_update(&c, data, n);
  nomp_sha256_hex(&c, hex);
}

int nomp_jit_cache_load(const char *hex, const char *ext, char **data, size_t *size) {
  const char *dir = nomp_jit_cache_dir();
  if (!dir) return 1;
  char path[PATH_MAX + 128];
  snprintf(path, sizeof(path), "%s/%s.%s", dir, hex, ext);
  FILE *fp = fopen(path, "rb");
  if (!fp) return 1;
  int err = 1;
  long n = -1;
  if (!fseek(fp, 0, SEEK_END) && (n = ftell(fp)) >= TRAILER && !fseek(fp, 0, SEEK_SET)) {
    char *buf = nomp_calloc(char, (size_t)n + 1);
    char want[65];
    if (fread(buf, 1, (size_t)n, fp) == (size_t)n) {
      digest_of(buf, (size_t)n - TRAILER, want);
      if (!memcmp(want, buf + n - TRAILER, TRAILER)) {
        buf[n - TRAILER] = '\0';
        *data = buf, *size = (size_t)n - TRAILER, err = 0;
      }
    }
    if (err) free(buf);
  }
  fclose(fp);
  return err;
}

int nomp_jit_cache_store(const char *hex, const char *ext, const void *data, size_t size) {
  const char *dir = nomp_jit_cache_dir();
  if (!dir) return 1;
  char path[PATH_MAX + 128], tmp[PATH_MAX + 160];
  snprintf(path, sizeof(path), "%s/%s.%s", dir, hex, ext);
  snprintf(tmp, sizeof(tmp), "%s.%ld.tmp", path, (long)getpid());
  FILE *fp = fopen(tmp, "wb");
  if (!fp) return 1;
  char trailer[65];
  digest_of(data, size, trailer);
  const int ok = fwrite(data, 1, size, fp) == size && fwrite(trailer, 1, TRAILER, fp) == TRAILER;
  if (fclose(fp) || !ok || rename(tmp, path)) {
    unlink(tmp);
    return 1;
  }
  return 0;
}

/* Add the bytes of one file to a key; non-zero if it cannot be read. */
int nomp_sha256_file(nomp_sha256_t *c, const char *path) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return 1;
  char buf[4096];
  size_t n, total = 0;
  while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) nomp_sha256_update(c, buf, n), total += n;
  const int err = ferror(fp);
  fclose(fp);
  nomp_sha256_update(c, &total, sizeof(total));
  return err;
}

static int by_name(const void *a, const void *b) { return strcmp(*(char *const *)a, *(char *const *)b); }

/* Add every "*<suffix>" file of a directory, in name order. */
int nomp_sha256_dir(nomp_sha256_t *c, const char *dir, const char *suffix) {
  DIR *d = opendir(dir);
  if (!d) return 1;
  char **names = NULL;
  size_t n = 0, cap = 0;
  const size_t ls = strlen(suffix);
  for (struct dirent *e; (e = readdir(d));) {
    const size_t l = strlen(e->d_name);
    if (l < ls || strcmp(e->d_name + l - ls, suffix)) continue;
    if (n == cap) cap = cap ? 2 * cap : 16, names = nomp_realloc(names, char *, cap);
    names[n++] = strdup(e->d_name);
  }
  closedir(d);
  if (n) qsort(names, n, sizeof(*names), by_name);
  int err = 0;
  for (size_t i = 0; i < n; i++) {
    char path[PATH_MAX + 300];
    snprintf(path, sizeof(path), "%s/%s", dir, names[i]);
    nomp_sha256_field(c, names[i]);
    err |= nomp_sha256_file(c, path);
    free(names[i]);
  }
  free(names);
  return err || n == 0;
}

/* ---- diagnostics (include/nomp-b200.h): the entry store as seen from outside, used by the CPU tests ----------------- */
NOMP_EXPORT const char *nomp_b200_jit_cache_dir(void) {
  nomp_jit_cache_reset(); /* re-read the environment, as nomp_init does */
  return nomp_jit_cache_dir();
}

NOMP_EXPORT int nomp_b200_jit_cache_put(const char *hex, const char *ext, const void *data, size_t size) {
  return nomp_jit_cache_store(hex, ext, data, size);
}

/* 0 and *size = length of the entry (copied to buf if it fits in cap); 1 if there is no intact entry */
NOMP_EXPORT int nomp_b200_jit_cache_get(const char *hex, const char *ext, void *buf, size_t cap, size_t *size) {
  char *data = NULL;
  size_t n = 0;
  if (nomp_jit_cache_load(hex, ext, &data, &n)) return 1;
  if (buf && n <= cap) memcpy(buf, data, n);
  if (size) *size = n;
  free(data);
  return 0;
}
