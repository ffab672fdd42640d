// Reference-side glue between APRIL-ANN's mathcore and libb200ann.so.
//
// These files are what a maintainer adds to packages/basics/mathcore/c_src/ --
// they are written against the reference's headers (gpu_mirrored_memory_block.h,
// cblas_headers.h, gpu_helper.h) and the C ABI in include/b200ann.h, and are
// compiled against both in this repository (integration/Makefile), so the
// signatures below are proven, not prose.
//
//   gemm_b200.cc   replaces mathcore/c_src/gemm.cu   (doGemm, doSparseMM)
//   gemv_b200.cc   replaces mathcore/c_src/gemv.cu   (doGemv, doSparseGemv)
//   axpy_b200.cc   replaces mathcore/c_src/axpy.cu   (doAxpy, doAxpyLoop, doSparseAxpy)
//
// Those three are exactly the translation units of the reference's USE_CUDA
// build that no longer compile with a current toolkit (cusparse<t>csrmm,
// cusparse<t>csrmv and cusparse<t>axpyi were removed from cuSPARSE); every
// other mathcore / matrix / ann .cu file still builds with nvcc 12.9 for
// sm_100a.  With the three replaced, a USE_CUDA build links again and its
// dense fp32 BLAS runs on the tcgen05 path.
//
// Build flags: -DUSE_CUDA -DUSE_B200 for the GPU build; with neither, the files
// reduce to the CPU wrappers and are drop-in for a CPU build too (that is how
// their CPU branch is tested here, see integration/README.md).
#ifndef B200_BRIDGE_H
#define B200_BRIDGE_H

#include "cblas_headers.h"
#include "complex_number.h"
#include "error_print.h"
#include "gpu_mirrored_memory_block.h"
#include "unused_variable.h"

#if defined(USE_B200) && !defined(USE_CUDA)
#error "USE_B200 needs the reference's USE_CUDA build (device half of GPUMirroredMemoryBlock)"
#endif

#ifdef USE_B200
#include <cuda_runtime_api.h>

#include "b200ann.h"
#include "gpu_helper.h"

namespace AprilMath {
  namespace B200 {

    /// The process-wide context, created on first use on the device the
    /// reference's GPUHelper initialised (gpu_helper.h:54-89 takes device 0).
    b200_ctx *context();

    /// Turns a non-zero C-ABI status into the reference's fatal error
    /// (util/c_src/error_print.h:55-77), keeping the library's message.
    inline void check(int status) {
      if (status != B200_OK) ERROR_EXIT1(status, "b200: %s\n", b200_last_error_string());
    }

    /// Orders one library call after what the reference's current stream holds
    /// and the reference's stream after the call: GPUHelper hands out stream 0
    /// or its own streams (gpu_helper.h:115-148), the library launches on its
    /// own non-blocking stream.
    class StreamOrder {
      cudaStream_t ref_stream, lib_stream;
      static cudaEvent_t event();
    public:
      StreamOrder() :
        ref_stream((cudaStream_t)CUDA::GPUHelper::getCurrentStream()),
        lib_stream((cudaStream_t)b200_stream(context())) {
        cudaEventRecord(event(), ref_stream);
        cudaStreamWaitEvent(lib_stream, event(), 0);
      }
      ~StreamOrder() {
        cudaEventRecord(event(), lib_stream);
        cudaStreamWaitEvent(ref_stream, event(), 0);
      }
    };

    /// double / complex / sparse operands have no tcgen05 path in the library:
    /// the caller has asked for the GPU, the data may live only there, so this
    /// is an error rather than a silent host fallback.
    inline void unsupported(const char *what) {
      ERROR_EXIT1(128, "b200: %s is not available on the GPU path; call set_use_cuda(false) on this matrix\n", what);
    }

  } // namespace B200
} // namespace AprilMath
#endif // USE_B200

#endif // B200_BRIDGE_H
