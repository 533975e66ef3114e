/*
 * spmv.h -- the drop-in C API of spmv-b200 (B200 / sm_100a).
 *
 * The four entry points below are exactly what a client of the reference binds
 * (reference include/spmv.h:19,26,41-52,65-71); signatures are unchanged.  Semantics on the GPU:
 *
 *   spmv_create_handle_all_in_one  builds DEVICE-RESIDENT state: the CSR arrays are uploaded once
 *       (or adopted in place when they already are device pointers) and the layout of the requested
 *       SPMV_METHODS value is built on the device.  `size` selects fp64 (== sizeof(double)) or fp32
 *       (anything else, as in reference serial_spmv.c:48-54).  `nthreads` and `vectorizedWay` are
 *       stored in the handle and ignored for launch geometry.  `MtxToken` is accepted and ignored.
 *   spmv  launches the kernel family of handle->spmvMethod.  The RowPtr/ColIdx/Matrix_Val arguments
 *       are ignored (the handle owns the device copies).  Vector_Val_X / Vector_Val_Y may be HOST
 *       pointers (staged through the handle; y is complete on return, like the reference) or DEVICE
 *       pointers (launch is asynchronous on the handle's stream; see spmv_b200.h).  A NULL handle is a
 *       silent no-op (reference common.c:285).
 *   spmv_clear_handle / spmv_destory_handle (sic)  free the device state (and the handle).
 *
 * Deviations a client can observe (everything else follows the reference):
 *   - the matrix VALUES are snapshotted at create (the reference re-reads the caller's arrays on every call for
 *     Serial / Parallel / Balanced*): after changing Matrix_Val call spmv_b200_update_values() (spmv_b200.h) or clear
 *     and re-create the handle (spmv_b200_info(h, "values_snapshotted") == 1 says so);
 *   - fp32 Method_CSR5SPMV is a real CSR5 and handle->spmvMethod stays Method_CSR5SPMV (the reference runs SELL and
 *     stores Method_SellCSigma, common.c:177-180);
 *   - the handle owns staging buffers, partial-sum arrays and events: calls on the SAME handle from several threads are
 *     serialised by a per-handle mutex (their launches reach the handle's stream call after call); two streams on one
 *     handle are not supported -- use one handle per stream (different handles are independent);
 *   - a pageable HOST x or y of at least 1 MiB that is passed twice in a row is page-locked in place so that its
 *     copies run at PCIe speed, and released when the caller switches buffers or clears / destroys the handle:
 *     do not free such a buffer while the handle is alive (SPMV_B200_PIN_HOST=0 switches this off).
 *
 * All functions return void like the reference; failures (CUDA errors, no device) are reported on
 * stderr, latched in spmv_b200_last_error() and leave the handle in a state where spmv() is a no-op.
 * There is no CPU fallback.
 */
#include "spmv_Defines.h"
#ifndef SPMV_B200_SPMV_H
#define SPMV_B200_SPMV_H
#ifndef GEMV_GEMV_H_
#define GEMV_GEMV_H_

/* The reference header pulls these in (spmv.h:9-10) and its sample driver relies on that
 * (omp_set_num_threads at src/samples/test_spmv.c:88).  Kept for source compatibility of clients;
 * the library itself is built with SPMV_B200_NO_HOST_HEADERS. */
#if !defined(SPMV_B200_NO_HOST_HEADERS) && !defined(__CUDACC__)
#  if defined(_OPENMP)
#    include <omp.h>
#  endif
#  if defined(__AVX__) || defined(__AVX2__)
#    include <immintrin.h>
#  endif
#endif

#define ALIGENED_SIZE 64 /* sic, reference spmv.h:12 */

#if defined(__cplusplus)
extern "C" {
#endif

/* replaces reference include/spmv.h:19 (common.c:54-61) */
SPMV_B200_API void spmv_destory_handle(spmv_Handle_t this_handle);

/* replaces reference include/spmv.h:26 (common.c:69-71) */
SPMV_B200_API void spmv_clear_handle(spmv_Handle_t this_handle);

/* replaces reference include/spmv.h:41-52 (common.c:123-190) */
SPMV_B200_API void spmv_create_handle_all_in_one(spmv_Handle_t *Handle,
                                   BASIC_INT_TYPE m,
                                   BASIC_INT_TYPE n,
                                   BASIC_INT_TYPE *RowPtr,
                                   BASIC_INT_TYPE *ColIdx,
                                   void *Matrix_Val,
                                   BASIC_SIZE_TYPE nthreads,
                                   SPMV_METHODS Function,
                                   BASIC_SIZE_TYPE size,
                                   VECTORIZED_WAY vectorizedWay,
                                   const char *MtxToken);

/* replaces reference include/spmv.h:65-71 (common.c:278-304) */
SPMV_B200_API void spmv(const spmv_Handle_t handle,
          BASIC_INT_TYPE m,
          const BASIC_INT_TYPE *RowPtr,
          const BASIC_INT_TYPE *ColIdx,
          const void *Matrix_Val,
          const void *Vector_Val_X,
          void *Vector_Val_Y);

#if defined(__cplusplus)
}
#endif
#endif /* GEMV_GEMV_H_ */
#endif /* SPMV_B200_SPMV_H */
