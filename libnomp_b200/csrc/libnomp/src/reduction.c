/* Finish of a reduce clause (replaces nomp_host_side_reduction, reference src/reduction.c:33-88).
 *
 * In the reference the device kernel leaves one partial per 512-element work group in the scratch buffer; this function
 * synchronises, copies ALL partials to the host and folds them in a serial loop, selecting the C type by
 * (domain, size == 4) (reference src/reduction.c:44-85).  Here the kernel (libnompk reduce.cu, or the generated
 * reduce skeleton) finishes the reduction on the device in the same launch, so what is left is bookkeeping:
 * map (domain, size) to the element type, let the backend all-reduce over ranks and wait for the published scalar,
 * and store it through the caller's pointer.  The result OVERWRITES *reduction_ptr (the incoming value is ignored,
 * reference python/reduction.py:68) and is valid when nomp_run returns (reference tests/nomp-api-500-impl.h:29-34). */
#include "nomp-impl.h"
#include "nompk.h"

int nomp_device_side_reduction(nomp_backend_t *backend, nomp_prog_t *prg) {
  const size_t size = (size_t)prg->reduction_size;
  if (size != 4 && size != 8)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Reduction variable must be 4 or 8 bytes wide.");
  if (prg->reduction_ptr == NULL)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Reduction variable was not passed to nomp_run.");
  int dtype;
  switch (prg->reduction_type) { /* same selection as the reference: 4 bytes -> int/unsigned/float, else the 8-byte type */
  case NOMP_INT: dtype = size == 4 ? NOMPK_I32 : NOMPK_I64; break;
  case NOMP_UINT: dtype = size == 4 ? NOMPK_U32 : NOMPK_U64; break;
  case NOMP_FLOAT: dtype = size == 4 ? NOMPK_F32 : NOMPK_F64; break;
  default:
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Reduction variable must be an integer or floating point scalar.");
  }
  return nomp_cuda_reduction_finish(backend, prg, dtype, size);
}
