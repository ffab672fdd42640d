// Replacement for packages/basics/mathcore/c_src/gemm.cu (see b200_bridge.h).
//
// doGemm<T>   : C = alpha op(A) op(B) + beta C       (cblas_headers.h:381-388)
// doSparseMM  : the same with a CSR / CSC left operand (cblas_headers.h:412-428)
//
// use_gpu && float  -> b200_sgemm (include/b200ann.h) on the device halves of
//                      the three memory blocks
// otherwise         -> the CBLAS the reference is linked with, on the host halves
#include "b200_bridge.h"

namespace AprilMath {

  namespace {

    inline void hostGemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k,
                         float alpha, const float *a, unsigned int lda, const float *b, unsigned int ldb,
                         float beta, float *c, unsigned int ldc) {
      cblas_sgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
    }
    inline void hostGemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k,
                         double alpha, const double *a, unsigned int lda, const double *b, unsigned int ldb,
                         double beta, double *c, unsigned int ldc) {
      cblas_dgemm(order, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
    }
    inline void hostGemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k,
                         ComplexF alpha, const ComplexF *a, unsigned int lda, const ComplexF *b,
                         unsigned int ldb, ComplexF beta, ComplexF *c, unsigned int ldc) {
      cblas_cgemm(order, ta, tb, m, n, k, &alpha, a, lda, b, ldb, &beta, c, ldc);
    }

#ifdef USE_B200
    // Only fp32 has a device path.  A column-major product is the row-major
    // product of the swapped operands: C^T = op(B)^T op(A)^T.
    inline bool deviceGemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k,
                           float alpha, const GPUMirroredMemoryBlock<float> *a, unsigned int lda,
                           const GPUMirroredMemoryBlock<float> *b, unsigned int ldb, float beta,
                           GPUMirroredMemoryBlock<float> *c, unsigned int ldc, unsigned int a_shift,
                           unsigned int b_shift, unsigned int c_shift) {
      const float *ap = a->getGPUForRead() + a_shift;
      const float *bp = b->getGPUForRead() + b_shift;
      float *cp = (beta == 0.0f ? c->getGPUForWrite() : c->getGPUForReadAndWrite()) + c_shift;
      B200::StreamOrder order_guard;
      if (order == CblasRowMajor)
        B200::check(b200_sgemm(B200::context(), ta != CblasNoTrans, tb != CblasNoTrans, m, n, k, alpha,
                               ap, (int)lda, bp, (int)ldb, beta, cp, (int)ldc));
      else
        B200::check(b200_sgemm(B200::context(), tb != CblasNoTrans, ta != CblasNoTrans, n, m, k, alpha,
                               bp, (int)ldb, ap, (int)lda, beta, cp, (int)ldc));
      return true;
    }
    template <typename T>
    inline bool deviceGemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, T,
                           const GPUMirroredMemoryBlock<T> *, unsigned int, const GPUMirroredMemoryBlock<T> *,
                           unsigned int, T, GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int,
                           unsigned int, unsigned int) {
      B200::unsupported("gemm on double / complex matrices");
      return false;
    }
#endif

  } // namespace

  template <typename T>
  void doGemm(CBLAS_ORDER major_order, CBLAS_TRANSPOSE a_transpose, CBLAS_TRANSPOSE b_transpose,
              int m, int n, int k, T alpha,
              const GPUMirroredMemoryBlock<T> *a, unsigned int a_inc,
              const GPUMirroredMemoryBlock<T> *b, unsigned int b_inc, T beta,
              GPUMirroredMemoryBlock<T> *c, unsigned int c_inc,
              unsigned int a_shift, unsigned int b_shift, unsigned int c_shift, bool use_gpu) {
#ifdef USE_B200
    if (use_gpu) {
      deviceGemm(major_order, a_transpose, b_transpose, m, n, k, alpha, a, a_inc, b, b_inc, beta, c, c_inc,
                 a_shift, b_shift, c_shift);
      return;
    }
#else
    UNUSED_VARIABLE(use_gpu);
#endif
    hostGemm(major_order, a_transpose, b_transpose, m, n, k, alpha,
             a->getPPALForRead() + a_shift, a_inc, b->getPPALForRead() + b_shift, b_inc, beta,
             c->getPPALForReadAndWrite() + c_shift, c_inc);
  }

  template <typename T>
  void doSparseMM(CBLAS_ORDER major_order, SPARSE_FORMAT sparse_format,
                  CBLAS_TRANSPOSE a_transpose, CBLAS_TRANSPOSE b_transpose,
                  int m, int n, int k, T alpha,
                  const GPUMirroredMemoryBlock<T> *a_values,
                  const Int32GPUMirroredMemoryBlock *a_indices,
                  const Int32GPUMirroredMemoryBlock *a_first_index,
                  const GPUMirroredMemoryBlock<T> *b, int b_inc, T beta,
                  GPUMirroredMemoryBlock<T> *c, int c_inc, int b_shift, int c_shift, bool use_gpu) {
#ifdef USE_B200
    if (use_gpu) B200::unsupported("sparse matrix products");
#else
    UNUSED_VARIABLE(use_gpu);
#endif
    // the reference's own CSR/CSC product (mathcore/c_src/cblas_headers.cc)
    cblas_sparse_mm(major_order, sparse_format, a_transpose, b_transpose, CblasNoTrans, m, n, k, alpha,
                    a_values->getPPALForRead(), a_indices->getPPALForRead(),
                    a_first_index->getPPALForRead(), b->getPPALForRead() + b_shift, b_inc, beta,
                    c->getPPALForReadAndWrite() + c_shift, c_inc);
  }

#define B200_INSTANTIATE_GEMM(T)                                                                      \
  template void doGemm<T>(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, T,            \
                          const GPUMirroredMemoryBlock<T> *, unsigned int,                             \
                          const GPUMirroredMemoryBlock<T> *, unsigned int, T,                          \
                          GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int, unsigned int,        \
                          unsigned int, bool);                                                         \
  template void doSparseMM<T>(CBLAS_ORDER, SPARSE_FORMAT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int,  \
                              int, T, const GPUMirroredMemoryBlock<T> *,                               \
                              const Int32GPUMirroredMemoryBlock *, const Int32GPUMirroredMemoryBlock *, \
                              const GPUMirroredMemoryBlock<T> *, int, T, GPUMirroredMemoryBlock<T> *,   \
                              int, int, int, bool);
  B200_INSTANTIATE_GEMM(float)
  B200_INSTANTIATE_GEMM(double)
  B200_INSTANTIATE_GEMM(ComplexF)
#undef B200_INSTANTIATE_GEMM

} // namespace AprilMath
