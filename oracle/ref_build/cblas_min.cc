// TEST INFRASTRUCTURE (oracle/): a minimal CBLAS so the reference's CPU path
// links without an external BLAS.
//
// The reference calls these routines from packages/basics/mathcore/c_src/
// {gemm,gemv,ger,axpy,copy,scal,dot,nrm2}.cu and ships no BLAS of its own
// (its NO_BLAS profile leaves sgemm/sger unimplemented,
// mathcore/c_src/cblas_headers.cc:100-119).  They are restated here from the
// published BLAS definitions as plain loops, float accumulation in double
// where the netlib reference accumulates in the working precision -- the
// result differs from an optimised BLAS only in summation order.  Nothing in
// the product links this file.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "atlas/cblas.h"
#include "atlas/clapack.h"
#include "lapacke.h"

namespace {

template <typename T>
inline const T* at(const T* x, int n, int inc, int i) {
  // BLAS convention: a negative increment walks the vector backwards.
  return inc >= 0 ? x + (long)i * inc : x + (long)(n - 1 - i) * (-inc);
}
template <typename T>
inline T* at(T* x, int n, int inc, int i) {
  return inc >= 0 ? x + (long)i * inc : x + (long)(n - 1 - i) * (-inc);
}

template <typename T, typename ACC>
T dot(int n, const T* x, int incx, const T* y, int incy) {
  ACC s = 0;
  for (int i = 0; i < n; ++i) s += (ACC)*at(x, n, incx, i) * (ACC)*at(y, n, incy, i);
  return (T)s;
}

template <typename T>
void axpy(int n, T a, const T* x, int incx, T* y, int incy) {
  for (int i = 0; i < n; ++i) *at(y, n, incy, i) += a * *at(x, n, incx, i);
}

// Element (r, c) of op(A) where A is stored with leading dimension ld in the
// given order.
template <typename T>
inline T elem(CBLAS_ORDER order, bool trans, const T* a, int ld, int r, int c) {
  if (trans) { int t = r; r = c; c = t; }
  return order == CblasRowMajor ? a[(long)r * ld + c] : a[(long)c * ld + r];
}
template <typename T>
inline T& celem(CBLAS_ORDER order, T* a, int ld, int r, int c) {
  return order == CblasRowMajor ? a[(long)r * ld + c] : a[(long)c * ld + r];
}

// C = alpha op(A) op(B) + beta C.  A column-major product is the row-major product of the swapped operands
// (C^T = op(B)^T op(A)^T), so only the row-major case is written out.  Per output row: a dot product over
// contiguous k when op(B)'s columns are contiguous in k (B transposed), otherwise an axpy of op(B)'s rows into
// an ACC-typed accumulator row, so that the innermost loop is unit-stride either way.
template <typename T, typename ACC>
void gemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k,
          T alpha, const T* a, int lda, const T* b, int ldb, T beta, T* c, int ldc) {
  if (order != CblasRowMajor) {
    gemm<T, ACC>(CblasRowMajor, tb, ta, n, m, k, alpha, b, ldb, a, lda, beta, c, ldc);
    return;
  }
  const bool at_ = ta != CblasNoTrans, bt_ = tb != CblasNoTrans;
#pragma omp parallel if ((long)m * n * k > 1 << 16)
  {
    std::vector<ACC> acc(bt_ ? 0 : n);
#pragma omp for schedule(static)
    for (int i = 0; i < m; ++i) {
      T* crow = c + (long)i * ldc;
      if (bt_) {
        for (int j = 0; j < n; ++j) {
          const T* brow = b + (long)j * ldb;   // op(B)[l, j] = B[j, l]
          ACC s = 0;
          if (!at_) {
            const T* arow = a + (long)i * lda;
            for (int l = 0; l < k; ++l) s += (ACC)arow[l] * (ACC)brow[l];
          } else {
            for (int l = 0; l < k; ++l) s += (ACC)a[(long)l * lda + i] * (ACC)brow[l];
          }
          crow[j] = (beta == T(0)) ? alpha * (T)s : alpha * (T)s + beta * crow[j];
        }
      } else {
        for (int j = 0; j < n; ++j) acc[j] = 0;
        for (int l = 0; l < k; ++l) {
          const ACC ail = (ACC)(at_ ? a[(long)l * lda + i] : a[(long)i * lda + l]);
          const T* brow = b + (long)l * ldb;   // op(B)[l, j] = B[l, j]
          for (int j = 0; j < n; ++j) acc[j] += ail * (ACC)brow[j];
        }
        for (int j = 0; j < n; ++j)
          crow[j] = (beta == T(0)) ? alpha * (T)acc[j] : alpha * (T)acc[j] + beta * crow[j];
      }
    }
  }
}

template <typename T, typename ACC>
void gemv(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, int m, int n, T alpha, const T* a, int lda,
          const T* x, int incx, T beta, T* y, int incy) {
  const bool t = ta != CblasNoTrans;
  const int rows = t ? n : m, cols = t ? m : n;
  for (int i = 0; i < rows; ++i) {
    ACC s = 0;
    for (int j = 0; j < cols; ++j) s += (ACC)elem(order, t, a, lda, i, j) * (ACC)*at(x, cols, incx, j);
    T& out = *at(y, rows, incy, i);
    out = (beta == T(0)) ? alpha * (T)s : alpha * (T)s + beta * out;
  }
}

template <typename T>
void ger(CBLAS_ORDER order, int m, int n, T alpha, const T* x, int incx, const T* y, int incy,
         T* a, int lda) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j)
      celem(order, a, lda, i, j) += alpha * *at(x, m, incx, i) * *at(y, n, incy, j);
}

typedef std::complex<float> cf;

[[noreturn]] void off_path(const char* what) {
  std::fprintf(stderr, "oracle/ref_build: %s is not on the training path and is not provided\n", what);
  std::abort();
}

}  // namespace

extern "C" {

#define DEFINE_REAL(P, T, ACC)                                                               \
  T cblas_##P##dot(int n, const T* x, int incx, const T* y, int incy) {                      \
    return dot<T, ACC>(n, x, incx, y, incy);                                                 \
  }                                                                                          \
  T cblas_##P##nrm2(int n, const T* x, int incx) {                                           \
    ACC s = 0;                                                                               \
    for (int i = 0; i < n; ++i) { ACC v = *at(x, n, incx, i); s += v * v; }                  \
    return (T)std::sqrt(s);                                                                  \
  }                                                                                          \
  void cblas_##P##copy(int n, const T* x, int incx, T* y, int incy) {                        \
    for (int i = 0; i < n; ++i) *at(y, n, incy, i) = *at(x, n, incx, i);                     \
  }                                                                                          \
  void cblas_##P##axpy(int n, T a, const T* x, int incx, T* y, int incy) {                   \
    axpy<T>(n, a, x, incx, y, incy);                                                         \
  }                                                                                          \
  void cblas_##P##scal(int n, T a, T* x, int incx) {                                         \
    for (int i = 0; i < n; ++i) *at(x, n, incx, i) *= a;                                     \
  }                                                                                          \
  void cblas_##P##gemv(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, int m, int n, T alpha, const T* a, \
                       int lda, const T* x, int incx, T beta, T* y, int incy) {              \
    gemv<T, ACC>(o, ta, m, n, alpha, a, lda, x, incx, beta, y, incy);                        \
  }                                                                                          \
  void cblas_##P##ger(CBLAS_ORDER o, int m, int n, T alpha, const T* x, int incx,            \
                      const T* y, int incy, T* a, int lda) {                                 \
    ger<T>(o, m, n, alpha, x, incx, y, incy, a, lda);                                        \
  }                                                                                          \
  void cblas_##P##gemm(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n,  \
                       int k, T alpha, const T* a, int lda, const T* b, int ldb, T beta,     \
                       T* c, int ldc) {                                                      \
    gemm<T, ACC>(o, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);                   \
  }                                                                                          \
  void catlas_##P##set(int n, T a, T* x, int incx) {                                         \
    for (int i = 0; i < n; ++i) *at(x, n, incx, i) = a;                                      \
  }

DEFINE_REAL(s, float, double)
DEFINE_REAL(d, double, double)

void cblas_ccopy(int n, const void* x, int incx, void* y, int incy) {
  const cf* xs = (const cf*)x; cf* ys = (cf*)y;
  for (int i = 0; i < n; ++i) *at(ys, n, incy, i) = *at(xs, n, incx, i);
}
void cblas_caxpy(int n, const void* alpha, const void* x, int incx, void* y, int incy) {
  axpy<cf>(n, *(const cf*)alpha, (const cf*)x, incx, (cf*)y, incy);
}
void cblas_cscal(int n, const void* alpha, void* x, int incx) {
  cf* xs = (cf*)x;
  for (int i = 0; i < n; ++i) *at(xs, n, incx, i) *= *(const cf*)alpha;
}
float cblas_scnrm2(int n, const void* x, int incx) {
  const cf* xs = (const cf*)x; double s = 0;
  for (int i = 0; i < n; ++i) s += std::norm(std::complex<double>(*at(xs, n, incx, i)));
  return (float)std::sqrt(s);
}
void cblas_zdotu_sub(int n, const void* x, int incx, const void* y, int incy, void* out) {
  typedef std::complex<double> cd;
  const cd* xs = (const cd*)x; const cd* ys = (const cd*)y; cd s = 0;
  for (int i = 0; i < n; ++i) s += *at(xs, n, incx, i) * *at(ys, n, incy, i);
  *(cd*)out = s;
}
void cblas_cgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int, int, const void*, const void*, int,
                 const void*, int, const void*, void*, int) { off_path("cblas_cgemv"); }
void cblas_cgeru(CBLAS_ORDER, int, int, const void*, const void*, int, const void*, int,
                 void*, int) { off_path("cblas_cgeru"); }
void cblas_cgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, const void*,
                 const void*, int, const void*, int, const void*, void*, int) { off_path("cblas_cgemm"); }

int clapack_sgetrf(CBLAS_ORDER, int, int, float*, int, int*) { off_path("clapack_sgetrf"); }
int clapack_sgetri(CBLAS_ORDER, int, float*, int, const int*) { off_path("clapack_sgetri"); }
int clapack_spotrf(CBLAS_ORDER, CBLAS_UPLO, int, float*, int) { off_path("clapack_spotrf"); }
int LAPACKE_sgesdd(int, char, int, int, float*, int, float*, float*, int, float*, int) { off_path("LAPACKE_sgesdd"); }
int LAPACKE_sgetrf(int, int, int, float*, int, int*) { off_path("LAPACKE_sgetrf"); }
int LAPACKE_sgetri(int, int, float*, int, const int*) { off_path("LAPACKE_sgetri"); }
int LAPACKE_spotrf(int, char, int, float*, int) { off_path("LAPACKE_spotrf"); }

}  // extern "C"
