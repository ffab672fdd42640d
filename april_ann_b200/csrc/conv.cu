// Convolution (valid, strided, NCHW), convolution bias and max-pooling kernels.
// The reference loops over output pixels on the host and issues one tiny GEMM plus two
// strided copies per pixel and per pass (ann/ann/c_src/convolution_component.cc:135-354),
// adds the bias with one axpy per pixel (convolution_bias_component.cc:120-222) and runs the
// max-pooling backward on the host (maxpooling_component.cc:208-260).  Here every pass is
// one launch with the image / error planes staged in shared memory.
//   W[n, C*kh*kw], window flattened in (plane,row,col) order (convolution_component.cc:108-118,175).
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int OT = 8;           // output planes per CTA in the forward / weight-gradient kernels
constexpr int CT = 4;           // input planes per CTA in the data-gradient kernel
constexpr size_t SMEM_CAP = 160 * 1024;

// ---------------------------------------------------------------- forward
// grid (B, ceil(n/OT)); smem: image [C*H*W] + weights [OT][CK]
__global__ void __launch_bounds__(TPB) conv_fwd_kernel(int C, int H, int W, int n, int kh, int kw, int sh, int sw,
                                                       int oH, int oW, const float *__restrict__ x,
                                                       const float *__restrict__ w, const float *__restrict__ bias,
                                                       int act, float *__restrict__ y) {
  extern __shared__ float smem[];
  const int b = blockIdx.x, o0 = blockIdx.y * OT;
  const int CK = C * kh * kw, CHW = C * H * W, P = oH * oW;
  float *img = smem, *ws = smem + CHW;
  const float *xb = x + (size_t)b * CHW;
  for (int i = threadIdx.x; i < CHW; i += TPB) img[i] = __ldg(xb + i);
  for (int i = threadIdx.x; i < OT * CK; i += TPB) {
    const int o = o0 + i / CK;
    ws[i] = (o < n) ? __ldg(w + (size_t)o * CK + (i % CK)) : 0.0f;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += TPB) {
    const int oy = p / oW, ox = p % oW;
    float acc[OT];
#pragma unroll
    for (int o = 0; o < OT; ++o) acc[o] = 0.0f;
    for (int c = 0; c < C; ++c)
      for (int i = 0; i < kh; ++i) {
        const float *row = img + (c * H + oy * sh + i) * W + ox * sw;
        const int kbase = (c * kh + i) * kw;
        for (int j = 0; j < kw; ++j) {
          const float v = row[j];
#pragma unroll
          for (int o = 0; o < OT; ++o) acc[o] = fmaf(v, ws[o * CK + kbase + j], acc[o]);
        }
      }
#pragma unroll
    for (int o = 0; o < OT; ++o) {
      if (o0 + o < n) {
        float v = acc[o];
        if (bias) v += __ldg(bias + o0 + o);
        if (act != B200_ACT_NONE) v = act_apply(act, v);
        y[((size_t)b * n + o0 + o) * P + p] = v;
      }
    }
  }
}

// ---------------------------------------------------------------- data gradient (gather form, deterministic)
// dx[b,c,y,x] = sum_{o,i,j} dy[b,o,(y-i)/sh,(x-j)/sw] * w[o,c,i,j]  for exact, in-range divisions
// grid (B, ceil(C/CT)); smem: dy planes [n*P] + weights [n][CT][kh*kw]
__global__ void __launch_bounds__(TPB) conv_dgrad_kernel(int C, int H, int W, int n, int kh, int kw, int sh, int sw,
                                                         int oH, int oW, const float *__restrict__ dy,
                                                         const float *__restrict__ w, float *__restrict__ dx) {
  extern __shared__ float smem[];
  const int b = blockIdx.x, c0 = blockIdx.y * CT;
  const int K2 = kh * kw, CK = C * K2, P = oH * oW;
  float *dys = smem, *ws = smem + n * P;
  const float *dyb = dy + (size_t)b * n * P;
  for (int i = threadIdx.x; i < n * P; i += TPB) dys[i] = __ldg(dyb + i);
  for (int i = threadIdx.x; i < n * CT * K2; i += TPB) {
    const int o = i / (CT * K2), r = i % (CT * K2), c = c0 + r / K2;
    ws[i] = (c < C) ? __ldg(w + (size_t)o * CK + c * K2 + (r % K2)) : 0.0f;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < H * W; q += TPB) {
    const int yy = q / W, xx = q % W;
    float acc[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[c] = 0.0f;
    for (int i = 0; i < kh; ++i) {
      const int ty = yy - i;
      if (ty < 0 || ty % sh) continue;
      const int oy = ty / sh;
      if (oy >= oH) continue;
      for (int j = 0; j < kw; ++j) {
        const int tx = xx - j;
        if (tx < 0 || tx % sw) continue;
        const int ox = tx / sw;
        if (ox >= oW) continue;
        const int p = oy * oW + ox, k = i * kw + j;
        for (int o = 0; o < n; ++o) {
          const float g = dys[o * P + p];
#pragma unroll
          for (int c = 0; c < CT; ++c) acc[c] = fmaf(g, ws[(o * CT + c) * K2 + k], acc[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CT; ++c)
      if (c0 + c < C) dx[((size_t)b * C + c0 + c) * H * W + q] = acc[c];
  }
}

// ---------------------------------------------------------------- weight gradient
// stage 1: grid (nchunks, ceil(n/OT)): each CTA walks its images, threads own weight elements
//          (c,i,j) and accumulate OT output planes in registers -> partial[chunk][n][CK]
// stage 2: dw = beta*dw + scale * sum_chunk partial      (fixed order: deterministic)
__global__ void __launch_bounds__(TPB) conv_wgrad_partial_kernel(int B, int C, int H, int W, int n, int kh, int kw,
                                                                 int sh, int sw, int oH, int oW, int imgs_per_chunk,
                                                                 const float *__restrict__ dy,
                                                                 const float *__restrict__ x,
                                                                 float *__restrict__ partial) {
  extern __shared__ float smem[];
  const int chunk = blockIdx.x, o0 = blockIdx.y * OT;
  const int K2 = kh * kw, CK = C * K2, CHW = C * H * W, P = oH * oW;
  float *img = smem, *dys = smem + CHW;  // [OT][P]
  const int b0 = chunk * imgs_per_chunk, b1 = min(B, b0 + imgs_per_chunk);
  for (int e0 = 0; e0 < CK; e0 += TPB) {
    const int e = e0 + threadIdx.x;
    const bool live = e < CK;
    const int c = live ? e / K2 : 0, i = live ? (e % K2) / kw : 0, j = live ? e % kw : 0;
    float acc[OT];
#pragma unroll
    for (int o = 0; o < OT; ++o) acc[o] = 0.0f;
    for (int b = b0; b < b1; ++b) {
      __syncthreads();
      const float *xb = x + (size_t)b * CHW;
      for (int t = threadIdx.x; t < CHW; t += TPB) img[t] = __ldg(xb + t);
      for (int t = threadIdx.x; t < OT * P; t += TPB) {
        const int o = o0 + t / P;
        dys[t] = (o < n) ? __ldg(dy + ((size_t)b * n + o) * P + (t % P)) : 0.0f;
      }
      __syncthreads();
      if (live) {
        const float *xc = img + (c * H + i) * W + j;
        for (int oy = 0; oy < oH; ++oy)
          for (int ox = 0; ox < oW; ++ox) {
            const float v = xc[oy * sh * W + ox * sw];
            const int p = oy * oW + ox;
#pragma unroll
            for (int o = 0; o < OT; ++o) acc[o] = fmaf(v, dys[o * P + p], acc[o]);
          }
      }
    }
    if (live) {
#pragma unroll
      for (int o = 0; o < OT; ++o)
        if (o0 + o < n) partial[((size_t)chunk * n + o0 + o) * CK + e] = acc[o];
    }
  }
}
__global__ void __launch_bounds__(TPB) chunk_reduce_kernel(int nchunks, size_t count, const float *__restrict__ partial,
                                                           float scale, float beta, float *__restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int k = 0; k < nchunks; ++k) s += partial[(size_t)k * count + i];
    out[i] = (beta != 0.0f ? beta * out[i] : 0.0f) + scale * s;
  }
}

// ---------------------------------------------------------------- convolution bias
__global__ void __launch_bounds__(TPB) conv_bias_fwd_kernel(size_t total, int n, int HW, const float *__restrict__ x,
                                                            const float *__restrict__ bias, float *__restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    y[i] = x[i] + __ldg(bias + (i / HW) % n);
}
// db[p] = beta*db[p] + scale * sum_{b,pixels} dy[b,p,:]: two stages with a fixed summation order.  Stage 1:
// CTA (plane, chunk of images) -> part[chunk][plane]; stage 2: one thread per plane sums the chunks.
// (One CTA per plane left 16 CTAs summing 4.7 M values each: 126 us on the critical tail of the C4 step.)
__global__ void __launch_bounds__(TPB) conv_bias_grad_partial_kernel(int B, int n, int HW, int per_chunk,
                                                                     const float *__restrict__ dy, float *__restrict__ part) {
  __shared__ float sm[TPB / 32];
  const int p = blockIdx.x, b0 = blockIdx.y * per_chunk, b1 = min(B, b0 + per_chunk);
  float s = 0.0f;
  for (int b = b0; b < b1; ++b) {
    const float *src = dy + ((size_t)b * n + p) * HW;
    for (int q = threadIdx.x; q < HW; q += TPB) s += __ldg(src + q);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < TPB / 32; ++i) t += sm[i];
    part[(size_t)blockIdx.y * n + p] = t;
  }
}
__global__ void conv_bias_grad_final_kernel(int n, int chunks, const float *__restrict__ part, float scale, float beta,
                                            float *__restrict__ db) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float t = 0.0f;
  for (int c = 0; c < chunks; ++c) t += part[(size_t)c * n + p];
  db[p] = (beta != 0.0f ? beta * db[p] : 0.0f) + scale * t;
}

// ---------------------------------------------------------------- max pooling
// one thread per output element; first maximum wins (matrix_ext_reductions.cu:307-320 uses '>');
// argmax holds the raw position inside the input tensor (maxpooling_component.cc:166-189)
__global__ void __launch_bounds__(TPB) maxpool_fwd_kernel(size_t total, int H, int W, int kh, int kw, int sh, int sw,
                                                          int oH, int oW, const float *__restrict__ x,
                                                          float *__restrict__ y, int32_t *__restrict__ argmax) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(t % oW), oy = (int)((t / oW) % oH);
    const size_t plane = t / ((size_t)oW * oH);
    const size_t base = plane * H * W + (size_t)(oy * sh) * W + ox * sw;
    float best = x[base];
    size_t pos = base;
    for (int i = 0; i < kh; ++i)
      for (int j = 0; j < kw; ++j) {
        const size_t q = base + (size_t)i * W + j;
        const float v = x[q];
        if (v > best) { best = v; pos = q; }
      }
    y[t] = best;
    if (argmax) argmax[t] = (int32_t)pos;
  }
}
// gather form of the scatter-add of maxpooling_component.cc:232-236: every input element sums the
// errors of the windows that selected it (deterministic, no atomics)
__global__ void __launch_bounds__(TPB) maxpool_bwd_kernel(size_t total_in, int H, int W, int kh, int kw, int sh, int sw,
                                                          int oH, int oW, const float *__restrict__ dy,
                                                          const int32_t *__restrict__ argmax, float *__restrict__ dx) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total_in; t += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(t % W), yy = (int)((t / W) % H);
    const size_t plane = t / ((size_t)W * H);
    float s = 0.0f;
    // windows (oy,ox) with oy*sh <= yy < oy*sh+kh
    const int oy_lo = max(0, (yy - kh + sh) / sh), oy_hi = min(oH - 1, yy / sh);
    const int ox_lo = max(0, (xx - kw + sw) / sw), ox_hi = min(oW - 1, xx / sw);
    for (int oy = oy_lo; oy <= oy_hi; ++oy)
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const size_t o = (plane * oH + oy) * oW + ox;
        if ((size_t)argmax[o] == t) s += dy[o];
      }
    dx[t] = s;
  }
}

inline int blocks_for(size_t n, int sm_count) {
  size_t b = (n + TPB - 1) / TPB, cap = (size_t)sm_count * 8;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

template <class K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > SMEM_CAP) {
    b200_set_error("convolution: image/weights tile needs %zu bytes of shared memory (limit %zu)", bytes, SMEM_CAP);
    return B200_ERR_UNSUPPORTED;
  }
  if (bytes > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return B200_OK;
}

}  // namespace

extern "C" int b200_conv2d_fwd(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw,
                               const float *x, const float *w, const float *bias, int act, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && w && y, "NULL pointer");
  ARG_CHECK(B > 0 && C > 0 && n > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && H >= kh && W >= kw, "bad geometry");
  const int oH = (H - kh) / sh + 1, oW = (W - kw) / sw + 1;
  const size_t smem = ((size_t)C * H * W + (size_t)OT * C * kh * kw) * sizeof(float);
  int st = set_smem(conv_fwd_kernel, smem);
  if (st) return st;
  dim3 grid(B, (n + OT - 1) / OT);
  conv_fwd_kernel<<<grid, TPB, smem, ctx->stream>>>(C, H, W, n, kh, kw, sh, sw, oH, oW, x, w, bias,
                                                    act == B200_ACT_LINEAR ? B200_ACT_NONE : act, y);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

extern "C" int b200_conv2d_bwd_data(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw,
                                    const float *dy, const float *w, float *dx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dy && w && dx, "NULL pointer");
  ARG_CHECK(B > 0 && C > 0 && n > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && H >= kh && W >= kw, "bad geometry");
  const int oH = (H - kh) / sh + 1, oW = (W - kw) / sw + 1;
  const size_t smem = ((size_t)n * oH * oW + (size_t)n * CT * kh * kw) * sizeof(float);
  int st = set_smem(conv_dgrad_kernel, smem);
  if (st) return st;
  dim3 grid(B, (C + CT - 1) / CT);
  conv_dgrad_kernel<<<grid, TPB, smem, ctx->stream>>>(C, H, W, n, kh, kw, sh, sw, oH, oW, dy, w, dx);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

extern "C" int b200_conv2d_bwd_weight(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh,
                                      int sw, const float *dy, const float *x, float scale, float beta, float *dw,
                                      float *db) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dy && x && dw, "NULL pointer");
  ARG_CHECK(B > 0 && C > 0 && n > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && H >= kh && W >= kw, "bad geometry");
  const int oH = (H - kh) / sh + 1, oW = (W - kw) / sw + 1;
  const int CK = C * kh * kw;
  // enough chunks to cover the machine ~2x with the (chunk, plane-tile) grid
  const int otiles = (n + OT - 1) / OT;
  int nchunks = (2 * ctx->sm_count + otiles - 1) / otiles;
  if (nchunks > B) nchunks = B;
  const int ipc = (B + nchunks - 1) / nchunks;
  nchunks = (B + ipc - 1) / ipc;
  float *partial = (float *)b200_scratch(ctx, (size_t)nchunks * n * CK * sizeof(float));
  if (!partial) { b200_set_error("scratch allocation failed"); return B200_ERR_ALLOC; }
  const size_t smem = ((size_t)C * H * W + (size_t)OT * oH * oW) * sizeof(float);
  int st = set_smem(conv_wgrad_partial_kernel, smem);
  if (st) return st;
  dim3 grid(nchunks, otiles);
  conv_wgrad_partial_kernel<<<grid, TPB, smem, ctx->stream>>>(B, C, H, W, n, kh, kw, sh, sw, oH, oW, ipc, dy, x,
                                                              partial);
  LAUNCH_CHECK(ctx);
  const size_t count = (size_t)n * CK;
  chunk_reduce_kernel<<<blocks_for(count, ctx->sm_count), TPB, 0, ctx->stream>>>(nchunks, count, partial, scale, beta,
                                                                              dw);
  LAUNCH_CHECK(ctx);
  if (db) return b200_conv_bias_grad(ctx, B, n, oH * oW, dy, scale, beta, db);
  return B200_OK;
}

extern "C" int b200_conv_bias_fwd(b200_ctx *ctx, int B, int n, int HW, const float *x, const float *bias, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && bias && y, "NULL pointer");
  const size_t total = (size_t)B * n * HW;
  if (!total) return B200_OK;
  conv_bias_fwd_kernel<<<blocks_for(total, ctx->sm_count), TPB, 0, ctx->stream>>>(total, n, HW, x, bias, y);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
extern "C" int b200_conv_bias_grad(b200_ctx *ctx, int B, int n, int HW, const float *dy, float scale, float beta,
                                   float *db) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dy && db, "NULL pointer");
  if (n <= 0) return B200_OK;
  int chunks = (2 * ctx->sm_count + n - 1) / n;   // ~2 CTAs per SM in all
  if (chunks > B) chunks = B;
  if (chunks < 1) chunks = 1;
  const int per_chunk = (B + chunks - 1) / chunks;
  chunks = (B + per_chunk - 1) / per_chunk;
  float *part = (float *)b200_scratch(ctx, ((size_t)chunks * n + 64) * sizeof(float));
  if (!part) { b200_set_error("scratch allocation failed"); return B200_ERR_ALLOC; }
  conv_bias_grad_partial_kernel<<<dim3((unsigned)n, (unsigned)chunks), TPB, 0, ctx->stream>>>(B, n, HW, per_chunk, dy, part);
  LAUNCH_CHECK(ctx);
  conv_bias_grad_final_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, chunks, part, scale, beta, db);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

extern "C" int b200_maxpool_fwd(b200_ctx *ctx, int B, int C, int H, int W, int kh, int kw, int sh, int sw,
                                const float *x, float *y, int32_t *argmax) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && y, "NULL pointer");
  ARG_CHECK(kh > 0 && kw > 0 && sh > 0 && sw > 0 && H >= kh && W >= kw, "bad geometry");
  ARG_CHECK((size_t)B * C * H * W < (size_t)INT32_MAX, "tensor too large for int32 positions");
  const int oH = (H - kh) / sh + 1, oW = (W - kw) / sw + 1;
  const size_t total = (size_t)B * C * oH * oW;
  maxpool_fwd_kernel<<<blocks_for(total, ctx->sm_count), TPB, 0, ctx->stream>>>(total, H, W, kh, kw, sh, sw, oH, oW, x,
                                                                             y, argmax);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
extern "C" int b200_maxpool_bwd(b200_ctx *ctx, int B, int C, int H, int W, int kh, int kw, int sh, int sw,
                                const float *dy, const int32_t *argmax, float *dx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dy && argmax && dx, "NULL pointer");
  const int oH = (H - kh) / sh + 1, oW = (W - kw) / sw + 1;
  const size_t total = (size_t)B * C * H * W;
  maxpool_bwd_kernel<<<blocks_for(total, ctx->sm_count), TPB, 0, ctx->stream>>>(total, H, W, kh, kw, sh, sw, oH, oW,
                                                                             dy, argmax, dx);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
