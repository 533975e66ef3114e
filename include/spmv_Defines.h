/*
 * spmv_Defines.h -- types of the drop-in C boundary of spmv-b200.
 *
 * Declaration-compatible with the reference's include/spmv_Defines.h:10-70 (same macro names, enum
 * values, struct field order and types) so that a client compiled against the reference header links
 * and runs against libspmv_b200.so unchanged.  What differs is behind `extraHandle`: it points at the
 * library's device-resident state instead of a per-method host struct.
 */
#ifndef SPMV_B200_SPMV_DEFINES_H
#define SPMV_B200_SPMV_DEFINES_H
/* the reference's own include guards, so that mixing both headers in one TU cannot redefine things */
#ifndef GEMV_SPMV_DEFINES_H
#define GEMV_SPMV_DEFINES_H

#if defined(__cplusplus)
extern "C" {
#endif

/* symbols exported by libspmv_b200.so (the library is built with -fvisibility=hidden) */
#ifndef SPMV_B200_API
#  if defined(SPMV_B200_BUILD) && defined(__GNUC__)
#    define SPMV_B200_API __attribute__((visibility("default")))
#  else
#    define SPMV_B200_API
#  endif
#endif

/* reference spmv_Defines.h:10-16 -- index and size types, overridable */
#ifndef BASIC_INT_TYPE
#define BASIC_INT_TYPE int
#endif
#ifndef BASIC_SIZE_TYPE
#define BASIC_SIZE_TYPE unsigned long
#endif

/* reference spmv_Defines.h:18-24.  Accepted, stored in the handle, ignored by every kernel -- exactly
 * as in the reference, whose kernels hard-code AVX2. */
typedef enum VECTORIZED_WAY {
    VECTOR_NONE,
    VECTOR_AVX2,
    VECTOR_AVX512,
    VECTOR_TOTAL_SIZE
} VECTORIZED_WAY;

/* reference spmv_Defines.h:26-37.  Each value selects one device layout + kernel family:
 *   Method_Serial        reference-order CSR kernel (bit-identical to the reference's Method_Serial)
 *   Method_Parallel      CSR-vector, sub-warp per row sized from the mean row length
 *   Method_Balanced      nnz-balanced row blocks (the reference's csrSplitter, one block per warp)
 *   Method_Balanced2     merge-path tiles (rows + nnz diagonal), carry fix-up pass
 *   Method_Balanced_Yid  equal-nnz tiles, partial first/last rows fixed up in tile order
 *   Method_SellCSigma    SELL-32-sigma slices, rows sorted inside sigma windows
 *   Method_CSR5SPMV      CSR5 tiles (omega = 32) with bit-flag segmented sums
 * Values outside [0, Method_Total_Size) fall back to Method_Serial (reference common.c:136). */
typedef enum SPMV_METHODS {
    Method_Serial,
    Method_Parallel,
    Method_Balanced,
    Method_Balanced2,
    Method_Balanced_Yid,
    Method_SellCSigma,
    Method_CSR5SPMV,
    Method_Total_Size,
    Method_Numa
} SPMV_METHODS;

/* exported name tables, reference common.c:306-339 */
extern SPMV_B200_API const char *Vectorized_names[];
extern SPMV_B200_API const char *Methods_names[];
extern SPMV_B200_API const char *funcNames[];

/* reference spmv_Defines.h:44-70.  PUBLIC: clients read fields (e.g. handle->index in
 * src/samples/test_spmv.c:95).  RowPtr/ColIdx/Matrix_Val keep the caller's (borrowed) pointers;
 * index and Y_temp stay NULL and Level_3_opt_used stays 0 (the METIS path is compiled out upstream). */
typedef struct spmv_Handle {
    SPMV_METHODS spmvMethod;
    BASIC_SIZE_TYPE data_size;
    BASIC_SIZE_TYPE nthreads;
    VECTORIZED_WAY vectorizedWay;
    int Level_3_opt_used;
    BASIC_INT_TYPE *RowPtr;
    BASIC_INT_TYPE *ColIdx;
    BASIC_INT_TYPE *index;
    void *Matrix_Val;
    void *Y_temp;
    void *extraHandle; /* -> device-resident state owned by libspmv_b200 */
} spmv_Handle;

typedef spmv_Handle *spmv_Handle_t;

/* helper macros of reference spmv_Defines.h:73-82 (used by client code) */
#define CONVERT_FLOAT(pointer) *((float*)(pointer))
#define CONVERT_DOUBLE(pointer) *((double*)(pointer))
#define CONVERT_FLOAT_T(pointer) ((float*)(pointer))
#define CONVERT_DOUBLE_T(pointer) ((double*)(pointer))
#define CONVERT_EQU(pointer, size, other) \
    ((size)==sizeof(double))?(CONVERT_DOUBLE(pointer)=(other)):(CONVERT_FLOAT(pointer)=(other))
#define CONVERT_ADDEQU(pointer1, size, pointer2) \
    ((size)==sizeof(double))? \
    (CONVERT_DOUBLE(pointer1)+=CONVERT_DOUBLE(pointer2)):(CONVERT_FLOAT(pointer1)+=CONVERT_FLOAT(pointer2))

/* Dot_s_Products[] / Dot_d_Products[] are declared by the reference (spmv_Defines.h:84-91) but defined
 * nowhere in it; they are intentionally not declared here. */

#if defined(__cplusplus)
}
#endif
#endif /* GEMV_SPMV_DEFINES_H */
#endif /* SPMV_B200_SPMV_DEFINES_H */
