// See b200_bridge.h.  Only built with -DUSE_CUDA -DUSE_B200.
#include "b200_bridge.h"

#ifdef USE_B200
namespace AprilMath {
  namespace B200 {

    b200_ctx *context() {
      static b200_ctx *ctx = 0;
      if (ctx == 0) {
        CUDA::GPUHelper::initHelper();  // the reference's cuBLAS handle and device 0 first
        check(b200_create(0, &ctx));
        // APRIL_B200_MATH=tf32 selects the tensor-core contraction (tolerance 2e-3
        // per GEMM output); the default keeps the reference's fp32 results.
        const char *m = getenv("APRIL_B200_MATH");
        if (m && m[0] == 't') check(b200_set_math_mode(ctx, B200_MATH_TF32));
      }
      return ctx;
    }

    cudaEvent_t StreamOrder::event() {
      static cudaEvent_t ev = 0;
      if (ev == 0) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      return ev;
    }

  } // namespace B200
} // namespace AprilMath
#endif
