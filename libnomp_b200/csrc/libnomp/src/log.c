/* Error registry and logging (behaviour of reference src/log.c:1-134; the tick-tock profiler of :136-264 has no
 * call sites in the reference and is kept as inert stubs). */
#include "nomp-impl.h"

const char *ERR_STR_USER_MAP_PTR_IS_INVALID = "Map pointer %p was not found on device.";
const char *ERR_STR_USER_DEVICE_IS_INVALID = "Device id %d passed into libnomp is not valid.";

typedef struct {
  char *text;
  int errorno;
} log_entry_t;

static log_entry_t *entries = NULL;
static unsigned n_entries = 0, cap_entries = 0;
static unsigned verbose_level = 0;

int nomp_log_set_verbose(unsigned verbose) {
  verbose_level = verbose;
  return 0;
}

unsigned nomp_log_get_verbose(void) { return verbose_level; }

int nomp_log_(const char *file, unsigned line, int errorno, nomp_log_type_t type, const char *fmt, ...) {
  static const char *kind[] = {"Error", "Warning", "Info"};
  char msg[BUFSIZ];
  int off = snprintf(msg, sizeof(msg), "[%s] %s:%u ", kind[type - 1], file, line);
  if (off < 0) off = 0;
  if ((size_t)off < sizeof(msg)) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg + off, sizeof(msg) - (size_t)off, fmt, ap);
    va_end(ap);
  }

  if (verbose_level >= (unsigned)type) {
    fprintf(stderr, "%s\n", msg);
    fflush(stderr);
  }
  if (type != NOMP_ERROR) return 0;

  if (n_entries == cap_entries) {
    cap_entries = cap_entries ? 2 * cap_entries : 16;
    entries = nomp_realloc(entries, log_entry_t, cap_entries);
  }
  entries[n_entries].text = strndup(msg, sizeof(msg));
  entries[n_entries].errorno = errorno;
  return (int)++n_entries; /* ids are 1-based */
}

NOMP_EXPORT char *nomp_get_err_str(unsigned id) {
  if (id == 0 || id > n_entries) return NULL;
  return strndup(entries[id - 1].text, BUFSIZ);
}

NOMP_EXPORT int nomp_get_err_no(unsigned id) {
  if (id == 0 || id > n_entries) return NOMP_USER_LOG_ID_IS_INVALID;
  return entries[id - 1].errorno;
}

void nomp_log_finalize(void) {
  for (unsigned i = 0; i < n_entries; i++) free(entries[i].text);
  free(entries);
  entries = NULL;
  n_entries = cap_entries = 0;
}

static int profile_level = 0;
int nomp_profile_set_level(int level) {
  profile_level = level;
  return 0;
}
void nomp_profile(const char *name, int toggle, int sync) { (void)name, (void)toggle, (void)sync; }
void nomp_profile_result(void) {}
void nomp_profile_finalize(void) {}
