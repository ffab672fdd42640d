// Shared declarations for the sm_100a kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200ann.h"

#define NEAR_ZERO_F 1e-6f  // packages/basics/mathcore/c_src/cmath_overloads.h:38

struct b200_ctx {
  int device = 0;
  int sm_count = 148;
  int math_mode = B200_MATH_FP32;
  cudaStream_t stream = nullptr;          // the stream kernels are launched on (main, or a side branch)
  cudaStream_t main_stream = nullptr;
  // side branches: independent work of a step (weight gradients, per-tensor SGD, loss statistics)
  // runs beside the critical path; captured into the step's CUDA graph as parallel branches
  static constexpr int kBranches = 3;
  cudaStream_t side_stream[kBranches] = {};
  cudaEvent_t ev_fork[kBranches] = {}, ev_side[kBranches] = {};
  bool side_open[kBranches] = {};
  int cur_branch = -1;                    // -1: main
  int sm_budget = 0;                      // SMs a persistent contraction may plan for (0: all)
  std::vector<std::pair<size_t, void *>> deferred_free;   // blocks released while branches are open
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_compute = nullptr, ev_comm = nullptr;
  cudaEvent_t ev_bucket[16] = {};        // completion of the async gradient buckets
  cudaEvent_t ev_fence[8] = {};          // b200_fence_record / b200_fence_wait
  uint64_t launches = 0;
  // caching pool: size -> free blocks ; ptr -> size for live blocks
  std::multimap<size_t, void *> free_blocks;
  std::map<void *, size_t> live_blocks;
  std::mutex mu;
  // scratch for reductions / split-K etc.
  void *scratch[kBranches + 1] = {};      // one per branch (+ main): concurrent branches must not share it
  size_t scratch_bytes[kBranches + 1] = {};
  std::vector<void *> retired_scratch;    // outgrown scratch blocks that captured graphs may still reference
  // NCCL (dlopen'ed lazily)
  void *nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  // tcgen05 GEMM state (tensor-map cache lives in gemm_tc.cu)
  void *tc_state = nullptr;
};

void b200_set_error(const char *fmt, ...);
int b200_check_cuda(cudaError_t e, const char *what, const char *file, int line);
void *b200_scratch(b200_ctx *ctx, size_t bytes);

#define CUDA_TRY(expr)                                                        \
  do {                                                                        \
    int _st = b200_check_cuda((expr), #expr, __FILE__, __LINE__);             \
    if (_st) return _st;                                                      \
  } while (0)

#define LAUNCH_CHECK(ctx)                                                     \
  do {                                                                        \
    (ctx)->launches++;                                                        \
    int _st = b200_check_cuda(cudaGetLastError(), "kernel launch", __FILE__, __LINE__); \
    if (_st) return _st;                                                      \
  } while (0)

// Kernels that run beside a contraction CTA inside a step (side branches) ask for the same shared-memory
// carve-out as the contraction (maximum shared memory): CTAs of kernels with different carve-outs are not
// co-scheduled on one SM, and with the default preference these kernels only got the SMs a contraction
// left idle.  Call once per kernel, before its launch.
// (function attributes belong to the device's context: the "already done" flags are kept per device, so
// that one process may drive several devices -- one host thread over N contexts, or a thread per device)
#define B200_MAX_DEVICES 64
#define ONCE_PER_DEVICE(ctx)                                          \
  ([&]() -> bool {                                                    \
    static bool _done[B200_MAX_DEVICES];                              \
    const int _d = (ctx)->device & (B200_MAX_DEVICES - 1);            \
    const bool _first = !_done[_d];                                   \
    _done[_d] = true;                                                 \
    return _first;                                                    \
  }())
#define PREFER_MAX_SMEM_ONCE(kernel)                                                                          \
  do {                                                                                                        \
    if (ONCE_PER_DEVICE(ctx))                                                                                 \
      cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
  } while (0)
// every C-ABI entry point that launches or allocates makes the context's device current first
int b200_make_current(b200_ctx *ctx);
#define B200_ENTER(ctx)                                \
  do {                                                 \
    if (ctx) {                                         \
      int _st = b200_make_current(ctx);                \
      if (_st) return _st;                             \
    }                                                  \
  } while (0)

#define ARG_CHECK(cond, msg)                                                  \
  do {                                                                        \
    if (!(cond)) {                                                            \
      b200_set_error("%s: bad argument: %s", __func__, msg);                  \
      return B200_ERR_BAD_ARG;                                                \
    }                                                                         \
  } while (0)

// ---------------------------------------------------------------- scalar functors
// Device restatement of the reference's scalar definitions (cmath_overloads.h:717-733,
// 969-989, 1053-1080, 1115-1124).  expf (not __expf): parity mode needs <=2 ulp.
__device__ __forceinline__ float act_apply(int act, float x) {
  switch (act) {
    case B200_ACT_LOGISTIC: return 1.0f / (expf(-x) + 1.0f);
    case B200_ACT_TANH: return 2.0f / (expf(-x) + 1.0f) - 1.0f;
    case B200_ACT_RELU: return x > 0.0f ? x : 0.0f;
    default: return x;
  }
}
// derivative evaluated from the activation OUTPUT y (relu: y>0 <=> x>0)
__device__ __forceinline__ float act_deriv_from_output(int act, float y) {
  switch (act) {
    case B200_ACT_LOGISTIC: {
      float v = fminf(fmaxf(y, NEAR_ZERO_F), 1.0f - NEAR_ZERO_F);
      return v * (1.0f - v);
    }
    case B200_ACT_TANH: {
      float v = fminf(fmaxf(y, -1.0f + NEAR_ZERO_F), 1.0f - NEAR_ZERO_F);
      return 0.5f * (1.0f - v * v);
    }
    case B200_ACT_RELU: return y > 0.0f ? 1.0f : 0.0f;
    default: return 1.0f;
  }
}

// Epilogue shared by the FFMA and the tcgen05 contractions:
//   v = alpha*acc ; v += bias[n] ; v = act(v) ; v *= act'(dsrc[m,n]) ; v += beta*C[m,n]
struct GemmEpilogue {
  float alpha = 1.0f, beta = 0.0f;
  const float *bias = nullptr;
  int act = B200_ACT_NONE;
  int dact = B200_ACT_NONE;
  const float *dsrc = nullptr;
  int ld_dsrc = 0;
};

__device__ __forceinline__ float epilogue_apply(const GemmEpilogue &e, float acc, int m, int n,
                                                float cold) {
  float v = e.alpha * acc;
  if (e.bias) v += __ldg(e.bias + n);
  if (e.act != B200_ACT_NONE) v = act_apply(e.act, v);
  if (e.dact != B200_ACT_NONE) v *= act_deriv_from_output(e.dact, __ldg(e.dsrc + (size_t)m * e.ld_dsrc + n));
  if (e.beta != 0.0f) v += e.beta * cold;
  return v;
}

// same, with the bias / derivative-source / old-C values already loaded (vectorised callers)
__device__ __forceinline__ float epilogue_apply_v(const GemmEpilogue &e, float acc, float bias, float dsrc,
                                                  float cold) {
  float v = e.alpha * acc;
  if (e.bias) v += bias;
  if (e.act != B200_ACT_NONE) v = act_apply(e.act, v);
  if (e.dact != B200_ACT_NONE) v *= act_deriv_from_output(e.dact, dsrc);
  if (e.beta != 0.0f) v += e.beta * cold;
  return v;
}

// internal entry points shared between translation units
int gemm_simt(b200_ctx *ctx, int transA, int transB, int M, int N, int K, const float *A, int lda,
              const float *B, int ldb, float *C, int ldc, const GemmEpilogue &ep);
// returns B200_ERR_UNSUPPORTED when the shape/alignment cannot use the tensor path
int gemm_tc(b200_ctx *ctx, int transA, int transB, int M, int N, int K, const float *A, int lda,
            const float *B, int ldb, float *C, int ldc, const GemmEpilogue &ep);
int gemm_dispatch(b200_ctx *ctx, int transA, int transB, int M, int N, int K, const float *A,
                  int lda, const float *B, int ldb, float *C, int ldc, const GemmEpilogue &ep);
void gemm_tc_destroy(b200_ctx *ctx);
// conv_tc.cu: the convolution contractions on the tensor cores (im2col + gemm_tc)
bool conv_tc_applicable(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw);
int conv_tc_fwd(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw, const float *x, const float *w,
                const float *bias, int act, float *y);
int conv_tc_bwd_data(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw, const float *dy,
                     const float *w, float *dx);
int conv_tc_bwd_weight(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw, const float *dy,
                       const float *x, float scale, float beta, float *dw);
// skinny.cu: layers with N <= 16 output neurons and the bias gradient (HBM-bound, exact fp32)
bool skinny_applicable(int M, int N, int K);
int skinny_fwd(b200_ctx *ctx, int M, int N, int K, const float *X, int ldx, const float *W, int ldw, const float *bias,
               int act, float *Y, int ldy);
int skinny_bwd_data(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy, const float *W, int ldw, int dact,
                    const float *Yprev, int ldyp, float *dX, int lddx);
int skinny_bwd_weight(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy, const float *X, int ldx, float scale,
                      float beta, float *dW, int lddw, float *db);
int colsum_scaled(b200_ctx *ctx, int M, int N, const float *dy, int ld, float scale, float beta, float *out);

// Programmatic dependent launch: a kernel launched with launch_pdl may start while its predecessor on the stream is
// still draining; it must call pdl_wait() before it touches anything the predecessor wrote, and may call
// pdl_launch_dependents() once the next kernel's prologue can no longer get in its way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
