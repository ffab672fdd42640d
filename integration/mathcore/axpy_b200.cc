// Replacement for packages/basics/mathcore/c_src/axpy.cu (see b200_bridge.h).
//
// doAxpy<T>       : y += alpha x                                  (cblas_headers.h:239-248)
// doAxpyLoop<T>   : `times` axpys walking x and y by a stride     (cblas_headers.h:250-262)
//                   -- the bias add (x_stride 0: one bias row onto every pattern,
//                   ann/ann/c_src/bias_component.cc:46-73) and the bias gradient
//                   (y_stride 0: every pattern's error row onto one vector, :87-122)
// doSparseAxpy<T> : y[idx[i]] += alpha x[i]                       (cblas_headers.h:289-298)
#include "b200_bridge.h"

namespace AprilMath {

  namespace {

    inline void hostAxpy(int n, float alpha, const float *x, unsigned int incx, float *y, unsigned int incy) {
      cblas_saxpy(n, alpha, x, incx, y, incy);
    }
    inline void hostAxpy(int n, double alpha, const double *x, unsigned int incx, double *y,
                         unsigned int incy) {
      cblas_daxpy(n, alpha, x, incx, y, incy);
    }
    inline void hostAxpy(int n, ComplexF alpha, const ComplexF *x, unsigned int incx, ComplexF *y,
                         unsigned int incy) {
      cblas_caxpy(n, &alpha, x, incx, y, incy);
    }

#ifdef USE_B200
    // y[j*incy] += alpha x[j*incx], j < n.  Contiguous vectors are one saxpy;
    // strided ones are the rank-1 update A[n,1] += alpha x 1^T with lda = incy.
    inline void deviceAxpyRaw(int n, float alpha, const float *x, unsigned int incx, float *y,
                              unsigned int incy) {
      if (incx == 1 && incy == 1) {
        B200::check(b200_saxpy(B200::context(), (size_t)n, alpha, x, y));
        return;
      }
      static float *one_dev = 0;
      if (one_dev == 0) {
        const float one = 1.0f;
        B200::check(b200_malloc(B200::context(), (void **)&one_dev, sizeof(float)));
        B200::check(b200_memcpy_h2d(B200::context(), one_dev, &one, sizeof(float)));
        B200::check(b200_sync(B200::context()));
      }
      B200::check(b200_sger(B200::context(), n, 1, alpha, x, (int)incx, one_dev, 1, y, (int)incy));
    }

    inline void deviceAxpy(int n, float alpha, const GPUMirroredMemoryBlock<float> *x, unsigned int incx,
                           unsigned int x_shift, GPUMirroredMemoryBlock<float> *y, unsigned int incy,
                           unsigned int y_shift) {
      const float *xp = x->getGPUForRead() + x_shift;
      float *yp = y->getGPUForReadAndWrite() + y_shift;
      B200::StreamOrder order_guard;
      deviceAxpyRaw(n, alpha, xp, incx, yp, incy);
    }

    inline void deviceAxpyLoop(int n, float alpha, GPUMirroredMemoryBlock<float> *x, unsigned int incx,
                               unsigned int x_shift, GPUMirroredMemoryBlock<float> *y, unsigned int incy,
                               unsigned int y_shift, unsigned int times, unsigned int x_stride,
                               unsigned int y_stride) {
      const float *xp = x->getGPUForRead() + x_shift;
      float *yp = y->getGPUForReadAndWrite() + y_shift;
      B200::StreamOrder order_guard;
      if (incx == 1 && incy == 1 && x_stride == 0 && alpha == 1.0f && y_stride == (unsigned int)n) {
        // one row broadcast onto `times` contiguous rows: the bias add, in place
        B200::check(b200_bias_fwd(B200::context(), (int)times, n, yp, xp, yp));
      }
      else if (incx == 1 && incy == 1 && y_stride == 0) {
        // `times` rows summed into one: the bias gradient  y = alpha * sum_rows(x) + 1 * y
        B200::check(b200_bias_grad(B200::context(), (int)times, n, xp, (int)x_stride, alpha, 1.0f, yp));
      }
      else {
        for (unsigned int i = 0; i < times; ++i)
          deviceAxpyRaw(n, alpha, xp + (size_t)i * x_stride, incx, yp + (size_t)i * y_stride, incy);
      }
    }

    template <typename T>
    inline void deviceAxpy(int, T, const GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int,
                           GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int) {
      B200::unsupported("axpy on double / complex matrices");
    }
    template <typename T>
    inline void deviceAxpyLoop(int, T, GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int,
                               GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int, unsigned int,
                               unsigned int, unsigned int) {
      B200::unsupported("axpy on double / complex matrices");
    }
#endif

  } // namespace

  template <typename T>
  void doAxpy(int N, T alpha, const GPUMirroredMemoryBlock<T> *x, unsigned int x_inc, unsigned int x_shift,
              GPUMirroredMemoryBlock<T> *y, unsigned int y_inc, unsigned int y_shift, bool use_gpu) {
#ifdef USE_B200
    if (use_gpu) {
      deviceAxpy(N, alpha, x, x_inc, x_shift, y, y_inc, y_shift);
      return;
    }
#else
    UNUSED_VARIABLE(use_gpu);
#endif
    hostAxpy(N, alpha, x->getPPALForRead() + x_shift, x_inc, y->getPPALForReadAndWrite() + y_shift, y_inc);
  }

  template <typename T>
  void doAxpyLoop(int N, T alpha, GPUMirroredMemoryBlock<T> *x, unsigned int x_inc, unsigned int x_shift,
                  GPUMirroredMemoryBlock<T> *y, unsigned int y_inc, unsigned int y_shift,
                  unsigned int times, const unsigned int x_stride, const unsigned int y_stride,
                  bool use_gpu) {
#ifdef USE_B200
    if (use_gpu) {
      deviceAxpyLoop(N, alpha, x, x_inc, x_shift, y, y_inc, y_shift, times, x_stride, y_stride);
      return;
    }
#else
    UNUSED_VARIABLE(use_gpu);
#endif
    const T *xp = x->getPPALForRead() + x_shift;
    T *yp = y->getPPALForReadAndWrite() + y_shift;
    for (unsigned int i = 0; i < times; ++i)
      hostAxpy(N, alpha, xp + (size_t)i * x_stride, x_inc, yp + (size_t)i * y_stride, y_inc);
  }

  template <typename T>
  void doSparseAxpy(int NNZ, T alpha, const GPUMirroredMemoryBlock<T> *x_values,
                    const Int32GPUMirroredMemoryBlock *x_indices, GPUMirroredMemoryBlock<T> *y,
                    unsigned int x_shift, unsigned int y_shift, unsigned int y_inc, bool use_gpu) {
#ifdef USE_B200
    if (use_gpu) B200::unsupported("sparse axpy");
#else
    UNUSED_VARIABLE(use_gpu);
#endif
    const T *xv = x_values->getPPALForRead() + x_shift;
    const int *xi = x_indices->getPPALForRead() + x_shift;
    T *yp = y->getPPALForReadAndWrite() + y_shift;
    for (int i = 0; i < NNZ; ++i) {
      T &dst = yp[(size_t)xi[i] * y_inc];
      dst = dst + alpha * xv[i];
    }
  }

#define B200_INSTANTIATE_AXPY(T)                                                                        \
  template void doAxpy<T>(int, T, const GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int,         \
                          GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int, bool);                \
  template void doAxpyLoop<T>(int, T, GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int,           \
                              GPUMirroredMemoryBlock<T> *, unsigned int, unsigned int, unsigned int,     \
                              const unsigned int, const unsigned int, bool);                             \
  template void doSparseAxpy<T>(int, T, const GPUMirroredMemoryBlock<T> *,                               \
                                const Int32GPUMirroredMemoryBlock *, GPUMirroredMemoryBlock<T> *,        \
                                unsigned int, unsigned int, unsigned int, bool);
  B200_INSTANTIATE_AXPY(float)
  B200_INSTANTIATE_AXPY(double)
  B200_INSTANTIATE_AXPY(ComplexF)
#undef B200_INSTANTIATE_AXPY

} // namespace AprilMath
