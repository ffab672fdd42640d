/* Minimal CBLAS declarations for building the reference's CPU path without an
 * external BLAS (test infrastructure; see oracle/ref_build/README.md).
 *
 * The reference's default profile includes <atlas/cblas.h>
 * (packages/basics/mathcore/c_src/cblas_headers.h:60-66) and calls the
 * routines below; their semantics are the published BLAS ones and are
 * implemented with plain loops in ../cblas_min.cc. */
#ifndef B200_ORACLE_MIN_CBLAS_H
#define B200_ORACLE_MIN_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif

enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 };
enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 };

#define B200_L1(P, T)                                                            \
  T cblas_##P##dot(int n, const T* x, int incx, const T* y, int incy);           \
  T cblas_##P##nrm2(int n, const T* x, int incx);                                \
  void cblas_##P##copy(int n, const T* x, int incx, T* y, int incy);             \
  void cblas_##P##axpy(int n, T alpha, const T* x, int incx, T* y, int incy);    \
  void cblas_##P##scal(int n, T alpha, T* x, int incx);                          \
  void cblas_##P##gemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, int m,   \
                       int n, T alpha, const T* a, int lda, const T* x, int incx,\
                       T beta, T* y, int incy);                                  \
  void cblas_##P##ger(enum CBLAS_ORDER order, int m, int n, T alpha, const T* x, \
                      int incx, const T* y, int incy, T* a, int lda);            \
  void cblas_##P##gemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta,          \
                       enum CBLAS_TRANSPOSE tb, int m, int n, int k, T alpha,    \
                       const T* a, int lda, const T* b, int ldb, T beta, T* c,   \
                       int ldc);                                                 \
  void catlas_##P##set(int n, T alpha, T* x, int incx);

B200_L1(s, float)
B200_L1(d, double)
#undef B200_L1

/* single-precision complex: interleaved (re, im) pairs behind void pointers */
void cblas_ccopy(int n, const void* x, int incx, void* y, int incy);
void cblas_caxpy(int n, const void* alpha, const void* x, int incx, void* y, int incy);
void cblas_cscal(int n, const void* alpha, void* x, int incx);
float cblas_scnrm2(int n, const void* x, int incx);
void cblas_zdotu_sub(int n, const void* x, int incx, const void* y, int incy, void* dotu);
void cblas_cgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, int m, int n,
                 const void* alpha, const void* a, int lda, const void* x, int incx,
                 const void* beta, void* y, int incy);
void cblas_cgeru(enum CBLAS_ORDER order, int m, int n, const void* alpha,
                 const void* x, int incx, const void* y, int incy, void* a, int lda);
void cblas_cgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                 int m, int n, int k, const void* alpha, const void* a, int lda,
                 const void* b, int ldb, const void* beta, void* c, int ldc);

#ifdef __cplusplus
}
#endif
#endif
