/* LAPACK entry points the reference's matrix package declares a use for
 * (packages/basics/matrix/c_src/matrix_ext_lapack.cu).  None of them is on the
 * training path; ../cblas_min.cc defines them to abort. */
#ifndef B200_ORACLE_MIN_CLAPACK_H
#define B200_ORACLE_MIN_CLAPACK_H
#include "cblas.h"
#ifdef __cplusplus
extern "C" {
#endif
int clapack_sgetrf(enum CBLAS_ORDER order, int m, int n, float* a, int lda, int* ipiv);
int clapack_sgetri(enum CBLAS_ORDER order, int n, float* a, int lda, const int* ipiv);
int clapack_spotrf(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, float* a, int lda);
#ifdef __cplusplus
}
#endif
#endif
