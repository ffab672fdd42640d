// Replacement for packages/basics/mathcore/c_src/gemv.cu (see b200_bridge.h).
//
// doGemv<T>       : y = alpha op(A) x + beta y          (cblas_headers.h:456-463)
// doSparseGemv<T> : the same with a CSR / CSC matrix    (cblas_headers.h:490-500)
//
// DotProductANNComponent takes this path when the bunch is one pattern
// (ann/ann/c_src/dot_product_component.cc:82,145).
#include "b200_bridge.h"

namespace AprilMath {

  namespace {

    inline void hostGemv(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, int m, int n, float alpha, const float *a,
                         unsigned int lda, const float *x, unsigned int incx, float beta, float *y,
                         unsigned int incy) {
      cblas_sgemv(order, ta, m, n, alpha, a, lda, x, incx, beta, y, incy);
    }
    inline void hostGemv(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, int m, int n, double alpha, const double *a,
                         unsigned int lda, const double *x, unsigned int incx, double beta, double *y,
                         unsigned int incy) {
      cblas_dgemv(order, ta, m, n, alpha, a, lda, x, incx, beta, y, incy);
    }
    inline void hostGemv(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, int m, int n, ComplexF alpha,
                         const ComplexF *a, unsigned int lda, const ComplexF *x, unsigned int incx,
                         ComplexF beta, ComplexF *y, unsigned int incy) {
      cblas_cgemv(order, ta, m, n, &alpha, a, lda, x, incx, &beta, y, incy);
    }

#ifdef USE_B200
    inline void deviceGemv(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, int m, int n, float alpha,
                           const GPUMirroredMemoryBlock<float> *a, unsigned int lda,
                           const GPUMirroredMemoryBlock<float> *x, unsigned int incx, float beta,
                           GPUMirroredMemoryBlock<float> *y, unsigned int incy, unsigned int a_shift,
                           unsigned int x_shift, unsigned int y_shift) {
      // a column-major m x n matrix is the row-major n x m matrix, transposed
      bool trans = (ta != CblasNoTrans);
      if (order != CblasRowMajor) {
        trans = !trans;
        int t = m; m = n; n = t;
      }
      const float *ap = a->getGPUForRead() + a_shift;
      const float *xp = x->getGPUForRead() + x_shift;
      float *yp = (beta == 0.0f ? y->getGPUForWrite() : y->getGPUForReadAndWrite()) + y_shift;
      B200::StreamOrder order_guard;
      B200::check(b200_sgemv(B200::context(), trans, m, n, alpha, ap, (int)lda, xp, (int)incx, beta, yp,
                             (int)incy));
    }
    template <typename T>
    inline void deviceGemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int, int, T, const GPUMirroredMemoryBlock<T> *,
                           unsigned int, const GPUMirroredMemoryBlock<T> *, unsigned int, T,
                           GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int, unsigned int,
                           unsigned int) {
      B200::unsupported("gemv on double / complex matrices");
    }
#endif

  } // namespace

  template <typename T>
  void doGemv(CBLAS_ORDER major_order, CBLAS_TRANSPOSE a_transpose, int m, int n, T alpha,
              const GPUMirroredMemoryBlock<T> *a, unsigned int a_inc,
              const GPUMirroredMemoryBlock<T> *x, unsigned int x_inc, T beta,
              GPUMirroredMemoryBlock<T> *y, unsigned int y_inc,
              unsigned int a_shift, unsigned int x_shift, unsigned int y_shift, bool use_gpu) {
#ifdef USE_B200
    if (use_gpu) {
      deviceGemv(major_order, a_transpose, m, n, alpha, a, a_inc, x, x_inc, beta, y, y_inc, a_shift, x_shift,
                 y_shift);
      return;
    }
#else
    UNUSED_VARIABLE(use_gpu);
#endif
    hostGemv(major_order, a_transpose, m, n, alpha, a->getPPALForRead() + a_shift, a_inc,
             x->getPPALForRead() + x_shift, x_inc, beta, y->getPPALForReadAndWrite() + y_shift, y_inc);
  }

  template <typename T>
  void doSparseGemv(SPARSE_FORMAT sparse_format, CBLAS_TRANSPOSE a_transpose, int m, int n, T alpha,
                    const GPUMirroredMemoryBlock<T> *a_values,
                    const Int32GPUMirroredMemoryBlock *a_indices,
                    const Int32GPUMirroredMemoryBlock *a_first_index,
                    const GPUMirroredMemoryBlock<T> *x, unsigned int x_inc, T beta,
                    GPUMirroredMemoryBlock<T> *y, unsigned int y_inc,
                    unsigned int x_shift, unsigned int y_shift, bool use_gpu) {
#ifdef USE_B200
    if (use_gpu) B200::unsupported("sparse matrix-vector products");
#else
    UNUSED_VARIABLE(use_gpu);
#endif
    cblas_sparse_mv(sparse_format, a_transpose, m, n, alpha, a_values->getPPALForRead(),
                    a_indices->getPPALForRead(), a_first_index->getPPALForRead(),
                    x->getPPALForRead() + x_shift, (int)x_inc, beta,
                    y->getPPALForReadAndWrite() + y_shift, (int)y_inc);
  }

#define B200_INSTANTIATE_GEMV(T)                                                                       \
  template void doGemv<T>(CBLAS_ORDER, CBLAS_TRANSPOSE, int, int, T, const GPUMirroredMemoryBlock<T> *, \
                          unsigned int, const GPUMirroredMemoryBlock<T> *, unsigned int, T,             \
                          GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int, unsigned int,         \
                          unsigned int, bool);                                                          \
  template void doSparseGemv<T>(SPARSE_FORMAT, CBLAS_TRANSPOSE, int, int, T,                            \
                                const GPUMirroredMemoryBlock<T> *, const Int32GPUMirroredMemoryBlock *, \
                                const Int32GPUMirroredMemoryBlock *, const GPUMirroredMemoryBlock<T> *, \
                                unsigned int, T, GPUMirroredMemoryBlock<T> *, unsigned int,             \
                                unsigned int, unsigned int, bool);
  B200_INSTANTIATE_GEMV(float)
  B200_INSTANTIATE_GEMV(double)
  B200_INSTANTIATE_GEMV(ComplexF)
#undef B200_INSTANTIATE_GEMV

} // namespace AprilMath
