/*
 * nomp.h -- public C API of libnomp, B200-native implementation (libnomp_b200).
 *
 * This header is the drop-in boundary towards user programs and nompcc-generated code.  Every constant and
 * signature below has the value/shape of the reference's public header (reference include/nomp.h:23-48 argument
 * and map-direction enums, :65-151 error codes, :159-176 functions), so programs and the reference's own
 * tests/nomp-api-*.c compile and link against this implementation unmodified.
 *
 * Error convention (reference src/log.c:88, include/nomp-impl.h:301-306): a function returns 0 on success and a
 * positive, 1-based log id on failure; nomp_get_err_no(id) yields one of the negative NOMP_* codes and
 * nomp_get_err_str(id) a heap copy of "[Error] <file>:<line> <text>" that the caller frees.  The one exception
 * is nomp_finalize() before nomp_init(), which returns NOMP_FINALIZE_FAILURE itself (reference src/nomp.c:671).
 */
#ifndef LIBNOMP_B200_NOMP_H_
#define LIBNOMP_B200_NOMP_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Kernel argument kinds for nomp_jit().  NOMP_JIT may be OR-ed in: the value is then fixed at jit time, passed as
 * a fourth vararg (pointer to the value) and dropped from nomp_run()'s argument list. */
typedef enum { NOMP_INT = 2048, NOMP_UINT = 4096, NOMP_FLOAT = 8192, NOMP_PTR = 16384 } nomp_arg_type_t;

/* nomp_update() operations (bit flags; NOMP_TO on an unmapped range implies NOMP_ALLOC). */
typedef enum { NOMP_ALLOC = 1, NOMP_TO = 2, NOMP_FROM = 4, NOMP_FREE = 8 } nomp_map_direction_t;

typedef enum { NOMP_JIT = 1 } nomp_arg_properties_t;

/* Error numbers returned by nomp_get_err_no(). */
#define NOMP_SUCCESS 0
#define NOMP_USER_INPUT_IS_INVALID (-128)
#define NOMP_USER_MAP_PTR_IS_INVALID (-130)
#define NOMP_USER_MAP_OP_IS_INVALID (-132)
#define NOMP_USER_LOG_ID_IS_INVALID (-134)
#define NOMP_INITIALIZE_FAILURE (-256)
#define NOMP_FINALIZE_FAILURE (-258)
#define NOMP_PY_CALL_FAILURE (-384)
#define NOMP_LOOPY_CONVERSION_FAILURE (-386)
#define NOMP_LOOPY_KNL_NAME_NOT_FOUND (-388)
#define NOMP_LOOPY_CODEGEN_FAILURE (-390)
#define NOMP_LOOPY_GRIDSIZE_FAILURE (-392)
#define NOMP_CUDA_FAILURE (-512)
#define NOMP_HIP_FAILURE (-514)
#define NOMP_OPENCL_FAILURE (-516)

/* Initialise the runtime.  Recognised arguments ("--nomp-<key> <value>", other tokens are skipped):
 * install-dir, backend, platform, device, verbose, profile, scripts-dir, annotations-script.  The environment
 * variables NOMP_INSTALL_DIR, NOMP_BACKEND, NOMP_PLATFORM, NOMP_DEVICE, NOMP_VERBOSE, NOMP_PROFILE and
 * NOMP_SCRIPTS_DIR override the command line.  The only backend of this implementation is "cuda". */
int nomp_init(int argc, const char **argv);

/* Allocate / copy / free the device image of host elements [start_index, end_index) of `ptr`. */
int nomp_update(void *ptr, size_t start_index, size_t end_index, size_t unit_size, nomp_map_direction_t op);

/* Build a kernel from a C loop nest.  `clauses` is a NULL-terminated array of triples:
 *   {"transform", <python module>, <function>}   user schedule, called as function(kernel, context)
 *   {"annotate", <key>, <value>}                 forwarded to the annotations script
 *   {"reduce", <variable>, "+" | "*" | "min" | "max"}
 * followed by nargs argument descriptions (const char *name, size_t size, int type [, void *value if NOMP_JIT]).
 * *id < 0 requests a build; a non-negative *id is a cache hit and returns immediately. */
int nomp_jit(int *id, const char *src, const char **clauses, int nargs, ...);

/* Launch kernel `id`; one void* per runtime argument, in nomp_jit() order (pointers to scalars, host pointers of
 * mapped arrays, and for a reduce clause the host address that receives the result). */
int nomp_run(int id, ...);

/* Wait for all previously issued device work of this runtime. */
int nomp_sync(void);

char *nomp_get_err_str(unsigned id);
int nomp_get_err_no(unsigned id);

int nomp_finalize(void);
int nomp_finalize_excluding_interpreter(void);

#ifdef __cplusplus
}
#endif

#endif /* LIBNOMP_B200_NOMP_H_ */
