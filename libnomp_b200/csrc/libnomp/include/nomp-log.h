/* Error/log registry (reference include/nomp-log.h:28-64, src/log.c:1-134). */
#ifndef LIBNOMP_B200_LOG_H_
#define LIBNOMP_B200_LOG_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { NOMP_ERROR = 1, NOMP_WARNING = 2, NOMP_INFO = 3 } nomp_log_type_t;

int nomp_log_set_verbose(unsigned verbose);
unsigned nomp_log_get_verbose(void);
/* Formats "[<Type>] <file>:<line> <message>".  Errors are stored and their 1-based id is returned; warnings and
 * infos return 0.  Use through nomp_log(). */
int nomp_log_(const char *file, unsigned line, int errorno, nomp_log_type_t type, const char *fmt, ...)
    __attribute__((format(printf, 5, 6)));
void nomp_log_finalize(void);

#define nomp_log(errorno, type, ...) nomp_log_(__FILE__, __LINE__, (errorno), (type), __VA_ARGS__)

/* INFO messages are formatted only when they would be printed (the reference formats one per nomp_check even at
 * verbose 0, reference include/nomp-impl.h:301-306 -- that cost is on the nomp_run path). */
#define nomp_info(...)                                                                                           \
  do {                                                                                                           \
    if (nomp_log_get_verbose() >= NOMP_INFO) nomp_log(0, NOMP_INFO, __VA_ARGS__);                                \
  } while (0)

/* Propagate a failing call: positive ids are errors. */
#define nomp_check(call)                                                                                         \
  do {                                                                                                           \
    nomp_info("Calling %s ...", #call);                                                                          \
    int nomp_check_err_ = (call);                                                                                \
    if (nomp_check_err_ > 0) return nomp_check_err_;                                                             \
  } while (0)

/* profiler stubs kept for API parity (reference src/log.c:136-264 has no call sites either) */
int nomp_profile_set_level(int level);
void nomp_profile(const char *name, int toggle, int sync);
void nomp_profile_result(void);
void nomp_profile_finalize(void);

#ifdef __cplusplus
}
#endif

#endif
