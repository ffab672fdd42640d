/* See atlas/clapack.h: declared for the build, not on the training path. */
#ifndef B200_ORACLE_MIN_LAPACKE_H
#define B200_ORACLE_MIN_LAPACKE_H
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
#ifdef __cplusplus
extern "C" {
#endif
int LAPACKE_sgesdd(int order, char jobz, int m, int n, float* a, int lda, float* s,
                   float* u, int ldu, float* vt, int ldvt);
int LAPACKE_sgetrf(int order, int m, int n, float* a, int lda, int* ipiv);
int LAPACKE_sgetri(int order, int n, float* a, int lda, const int* ipiv);
int LAPACKE_spotrf(int order, char uplo, int n, float* a, int lda);
#ifdef __cplusplus
}
#endif
#endif
