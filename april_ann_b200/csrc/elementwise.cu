// HBM-bound element-wise and small reduction kernels: float4 grid-stride maps with
// warp-shuffle reductions.  Replace the generic Map1/Map2 functor launches and the
// two-pass shared-memory reductions of the reference
// (mathcore/c_src/cuda_kernel_templates.h:87-302,651-847; activation_function_kernels.cu:60-182;
//  axpy.cu:111 axpyLoopKernel for the bias add / bias gradient).
#include "common.cuh"

namespace {

constexpr int TPB = 256;

inline int grid_for(size_t n_vec, int sm_count) {
  size_t blocks = (n_vec + TPB - 1) / TPB;
  size_t cap = (size_t)sm_count * 16;  // multiple of the SM count, 16 resident 256-thread CTAs/SM at most 8
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

// ---- generic map kernels ----------------------------------------------------
template <class F>
__global__ void __launch_bounds__(TPB) map1_kernel(size_t n, const float *x, float *y, F f,
                                                   bool vec) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const size_t n4 = n >> 2;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    float4 *y4 = reinterpret_cast<float4 *>(y);
    for (size_t i = tid; i < n4; i += nth) {
      float4 v = x4[i];
      v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
      y4[i] = v;
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += nth) y[i] = f(x[i]);
  } else {
    for (size_t i = tid; i < n; i += nth) y[i] = f(x[i]);
  }
}

template <class F>
__global__ void __launch_bounds__(TPB) map2_kernel(size_t n, const float *a, const float *b,
                                                   float *y, F f, bool vec) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const size_t n4 = n >> 2;
    const float4 *a4 = reinterpret_cast<const float4 *>(a);
    const float4 *b4 = reinterpret_cast<const float4 *>(b);
    float4 *y4 = reinterpret_cast<float4 *>(y);
    for (size_t i = tid; i < n4; i += nth) {
      float4 p = a4[i], q = b4[i], v;
      v.x = f(p.x, q.x); v.y = f(p.y, q.y); v.z = f(p.z, q.z); v.w = f(p.w, q.w);
      y4[i] = v;
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += nth) y[i] = f(a[i], b[i]);
  } else {
    for (size_t i = tid; i < n; i += nth) y[i] = f(a[i], b[i]);
  }
}

struct ActFwd { int act; __device__ float operator()(float x) const { return act_apply(act, x); } };
struct ActBwd { int act; __device__ float operator()(float y, float dy) const { return act_deriv_from_output(act, y) * dy; } };
struct Axpy { float a; __device__ float operator()(float x, float y) const { return fmaf(a, x, y); } };
struct Scal { float a; __device__ float operator()(float x) const { return a * x; } };
struct Ident { __device__ float operator()(float x) const { return x; } };
struct Mul { __device__ float operator()(float x, float y) const { return x * y; } };

template <class F>
int launch_map1(b200_ctx *ctx, size_t n, const float *x, float *y, F f) {
  if (n == 0) return B200_OK;
  const bool vec = aligned16(x) && aligned16(y);
  map1_kernel<F><<<grid_for(vec ? (n >> 2) + 1 : n, ctx->sm_count), TPB, 0, ctx->stream>>>(n, x, y, f, vec);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
template <class F>
int launch_map2(b200_ctx *ctx, size_t n, const float *a, const float *b, float *y, F f) {
  if (n == 0) return B200_OK;
  const bool vec = aligned16(a) && aligned16(b) && aligned16(y);
  map2_kernel<F><<<grid_for(vec ? (n >> 2) + 1 : n, ctx->sm_count), TPB, 0, ctx->stream>>>(n, a, b, y, f, vec);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

// ---- bias --------------------------------------------------------------------
// y[m,n] = x[m,n] + b[n]   (bias_component.cc:46-73)
__global__ void __launch_bounds__(TPB) bias_fwd_kernel(int M, int N, const float *__restrict__ x,
                                                       const float *__restrict__ b, float *__restrict__ y) {
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    y[i] = x[i] + __ldg(b + (i % N));
}

// ---- reductions ----------------------------------------------------------------
template <bool SQ, bool ACCUM>
__global__ void __launch_bounds__(TPB) reduce_partial_kernel(size_t n, const float *__restrict__ x, float *__restrict__ part) {
  __shared__ float sm[TPB / 32];
  float s = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = x[i];
    s += SQ ? v * v : v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < TPB / 32 ? sm[threadIdx.x] : 0.0f;
    t = warp_sum(t);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
  }
}
template <bool ACCUM>
__global__ void reduce_final_kernel(int nparts, const float *__restrict__ part, float *__restrict__ out) {
  float s = 0.0f;
  for (int i = threadIdx.x; i < nparts; i += 32) s += part[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) *out = ACCUM ? *out + s : s;
}

template <bool SQ, bool ACCUM>
int reduce_impl(b200_ctx *ctx, size_t n, const float *x, float *out) {
  int blocks = grid_for(n, ctx->sm_count);
  if (blocks > 1024) blocks = 1024;
  float *part = (float *)b200_scratch(ctx, 1024 * sizeof(float));
  if (!part) { b200_set_error("scratch allocation failed"); return B200_ERR_ALLOC; }
  reduce_partial_kernel<SQ, ACCUM><<<blocks, TPB, 0, ctx->stream>>>(n, x, part);
  LAUNCH_CHECK(ctx);
  reduce_final_kernel<ACCUM><<<1, 32, 0, ctx->stream>>>(blocks, part, out);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

// stats[0] += sum, stats[1] += sum sq, stats[2] += M, stats[3] = sum of this bunch, in double (loss_function.h:100-107 keeps a
// double running mean/variance on the host; here the sufficient statistics stay on the device)
__global__ void loss_accumulate_kernel(int M, const float *__restrict__ rows, double *__restrict__ stats) {
  __shared__ double s1[8], s2[8];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    double v = (double)rows[i];
    a += v;
    b += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb = 0;
    for (int i = 0; i < 8; ++i) { ta += s1[i]; tb += s2[i]; }
    stats[0] += ta;
    stats[1] += tb;
    stats[2] += (double)M;
    stats[3] = ta;  // sum over this bunch only
  }
}

__global__ void counter_increment_kernel(int64_t *c) { *c += 1; }

// out[i,:] = data[idx[i],:]   (datasetToken.h:182-209 gathers row by row on the host)
__global__ void __launch_bounds__(TPB) gather_rows_kernel(int nrows, int cols, const float *__restrict__ data,
                                                          const int32_t *__restrict__ idx, float *__restrict__ out,
                                                          bool vec) {
  const int row = blockIdx.x;
  if (row >= nrows) return;
  const float *src = data + (size_t)__ldg(idx + row) * cols;
  float *dst = out + (size_t)row * cols;
  if (vec) {
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int i = threadIdx.x; i < (cols >> 2); i += blockDim.x) d4[i] = s4[i];
  } else {
    for (int i = threadIdx.x; i < cols; i += blockDim.x) dst[i] = src[i];
  }
}

}  // namespace

extern "C" int b200_actf_fwd(b200_ctx *ctx, int act, size_t n, const float *x, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && y, "NULL pointer");
  ARG_CHECK(act == B200_ACT_LOGISTIC || act == B200_ACT_TANH || act == B200_ACT_RELU ||
                act == B200_ACT_LINEAR || act == B200_ACT_NONE,
            "row-wise activations use b200_softmax_fwd / b200_log_softmax_fwd");
  return launch_map1(ctx, n, x, y, ActFwd{act == B200_ACT_LINEAR ? B200_ACT_NONE : act});
}
extern "C" int b200_actf_bwd(b200_ctx *ctx, int act, size_t n, const float *y, const float *dy, float *dx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && y && dy && dx, "NULL pointer");
  return launch_map2(ctx, n, y, dy, dx, ActBwd{act == B200_ACT_LINEAR ? B200_ACT_NONE : act});
}
extern "C" int b200_saxpy(b200_ctx *ctx, size_t n, float alpha, const float *x, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && y, "NULL pointer");
  return launch_map2(ctx, n, x, y, y, Axpy{alpha});
}
extern "C" int b200_sscal(b200_ctx *ctx, size_t n, float alpha, float *x) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x, "NULL pointer");
  return launch_map1(ctx, n, x, x, Scal{alpha});
}
extern "C" int b200_scopy(b200_ctx *ctx, size_t n, const float *x, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && y, "NULL pointer");
  return launch_map1(ctx, n, x, y, Ident{});
}
extern "C" int b200_cmul(b200_ctx *ctx, size_t n, const float *x, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && y, "NULL pointer");
  return launch_map2(ctx, n, x, y, y, Mul{});
}
extern "C" int b200_sum(b200_ctx *ctx, size_t n, const float *x, float *out) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && out, "NULL pointer");
  return reduce_impl<false, false>(ctx, n, x, out);
}
extern "C" int b200_nrm2sq(b200_ctx *ctx, size_t n, const float *x, float *out) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && out, "NULL pointer");
  return reduce_impl<true, true>(ctx, n, x, out);
}
extern "C" int b200_bias_fwd(b200_ctx *ctx, int M, int N, const float *x, const float *b, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && b && y, "NULL pointer");
  if (M <= 0 || N <= 0) return B200_OK;
  bias_fwd_kernel<<<grid_for((size_t)M * N, ctx->sm_count), TPB, 0, ctx->stream>>>(M, N, x, b, y);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
extern "C" int b200_bias_grad(b200_ctx *ctx, int M, int N, const float *dy, int lddy, float scale,
                              float beta, float *db) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dy && db, "NULL pointer");
  if (N <= 0) return B200_OK;
  // db[n] = beta*db[n] + scale * sum_m dy[m,n]  (bias_component.cc:87-122): two-stage column sum, skinny.cu
  return colsum_scaled(ctx, M, N, dy, lddy, scale, beta, db);
}
extern "C" int b200_loss_accumulate(b200_ctx *ctx, int M, const float *loss_rows, double *stats) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && loss_rows && stats, "NULL pointer");
  PREFER_MAX_SMEM_ONCE(loss_accumulate_kernel);
  loss_accumulate_kernel<<<1, 256, 0, ctx->stream>>>(M, loss_rows, stats);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
extern "C" int b200_counter_increment(b200_ctx *ctx, int64_t *count_dev) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && count_dev, "NULL pointer");
  counter_increment_kernel<<<1, 1, 0, ctx->stream>>>(count_dev);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
extern "C" int b200_gather_rows(b200_ctx *ctx, int nrows, int cols, const float *data, const int32_t *idx,
                                float *out) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && data && idx && out, "NULL pointer");
  if (nrows <= 0 || cols <= 0) return B200_OK;
  const bool vec = (cols % 4 == 0) && aligned16(data) && aligned16(out);
  gather_rows_kernel<<<nrows, TPB, 0, ctx->stream>>>(nrows, cols, data, idx, out, vec);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
