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
// Weight gradient of a convolution with a SHORT window (C*kh*kw <= 32 window elements, e.g. the 25 of a first
// layer on grey images) -- a long reduction over every pixel of every image into a tiny result.  Each warp owns
// four output planes and one slice of the pixels; lane = window element.  Per pixel a warp issues one broadcast
// 16-byte read of the four error values, one read of the image value of its window element, and four FFMA.
// The error planes are stored pixel-major in shared memory ([pixel][plane]) so that the four values are one word.
// (A kernel that gives a thread one window element and walks all planes kept 25 of 256 threads busy: 146 us for
// C4's first layer; one thread per weight element with scalar reads issued 7.6 instructions per FFMA: 70 us.)
constexpr int WS_SLICES = 8;
__global__ void __launch_bounds__(1024) conv_wgrad_small_kernel(int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw,
                                                               int oH, int oW, int imgs_per_chunk,
                                                               const float *__restrict__ dy, const float *__restrict__ x,
                                                               float *__restrict__ partial) {
  extern __shared__ float smem[];
  const int K2 = kh * kw, CK = C * K2, CHW = C * H * W, P = oH * oW;
  const int n4 = (n + 3) & ~3;                    // planes padded to a multiple of 4 (zero error there)
  float *img = smem;                              // [CHW]
  float *dyt = smem + ((CHW + 3) & ~3);           // [P][n4], 16-byte aligned rows
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int groups = n4 >> 2;                     // warps needed per slice
  const int e = lane;                             // window element
  const bool live_e = e < CK;
  const int c = live_e ? e / K2 : 0, i = live_e ? (e % K2) / kw : 0, j = live_e ? e % kw : 0;
  const int b0 = blockIdx.x * imgs_per_chunk, b1 = min(B, b0 + imgs_per_chunk);
  const int nwarps = blockDim.x >> 5;
  // (plane group, pixel slice) pairs are dealt to the warps round-robin
  float acc[4][4];                                // up to 4 pairs per warp
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[a][o] = 0.0f;
  const int pairs = groups * WS_SLICES;
  for (int b = b0; b < b1; ++b) {
    __syncthreads();
    const float *xb = x + (size_t)b * CHW;
    for (int q = t; q < CHW; q += blockDim.x) img[q] = __ldg(xb + q);
    const float *dyb = dy + (size_t)b * n * P;
    // four planes of one pixel per thread: coalesced reads along the pixels of each plane, one 16-byte write
    for (int q = t; q < groups * P; q += blockDim.x) {
      const int g = q / P, p = q - g * P;
      float4 v;
      v.x = 4 * g + 0 < n ? __ldg(dyb + (size_t)(4 * g + 0) * P + p) : 0.0f;
      v.y = 4 * g + 1 < n ? __ldg(dyb + (size_t)(4 * g + 1) * P + p) : 0.0f;
      v.z = 4 * g + 2 < n ? __ldg(dyb + (size_t)(4 * g + 2) * P + p) : 0.0f;
      v.w = 4 * g + 3 < n ? __ldg(dyb + (size_t)(4 * g + 3) * P + p) : 0.0f;
      reinterpret_cast<float4 *>(dyt)[p * groups + g] = v;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int pr = warp + a * nwarps;
      if (pr < pairs) {
        const int g = pr % groups, sl = pr / groups;
        const float *xc = img + (c * H + i) * W + j;
        // slice sl takes the output rows oy = sl, sl + WS_SLICES, ...
        for (int oy = sl; oy < oH; oy += WS_SLICES) {
          const float *xr = xc + oy * sh * W;
          const float4 *dr = reinterpret_cast<const float4 *>(dyt + (size_t)(oy * oW) * n4) + g;
#pragma unroll 4
          for (int ox = 0; ox < oW; ++ox) {
            const float4 d = dr[ox * groups];
            const float v = live_e ? xr[ox * sw] : 0.0f;
            acc[a][0] = fmaf(v, d.x, acc[a][0]);
            acc[a][1] = fmaf(v, d.y, acc[a][1]);
            acc[a][2] = fmaf(v, d.z, acc[a][2]);
            acc[a][3] = fmaf(v, d.w, acc[a][3]);
          }
        }
      }
    }
  }
  // partial[chunk][slice][o][e]: the slices are summed by the chunk reduction
  if (live_e) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int pr = warp + a * nwarps;
      if (pr < pairs) {
        const int g = pr % groups, sl = pr / groups;
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (4 * g + o < n) partial[(((size_t)blockIdx.x * WS_SLICES + sl) * n + 4 * g + o) * CK + e] = acc[a][o];
      }
    }
  }
}
// out[i] = beta*out[i] + scale * sum_chunk partial[chunk][i]: one warp per output element, lanes stride over the
// chunks, fixed shuffle tree (deterministic).  (One thread per element walked hundreds of chunks one dependent
// load at a time: 24 us for the 400 weights of C4's first layer.)
__global__ void __launch_bounds__(TPB) chunk_reduce_kernel(int nchunks, size_t count, const float *__restrict__ partial,
                                                           float scale, float beta, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t i = warp; i < count; i += nwarps) {
    float s = 0.0f;
    for (int k = lane; k < nchunks; k += 32) s += __ldg(partial + (size_t)k * count + i);
    s = warp_sum(s);
    if (lane == 0) out[i] = (beta != 0.0f ? beta * out[i] : 0.0f) + scale * s;
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
  if ((HW & 3) == 0 && ((((uintptr_t)dy) & 15) == 0)) {
    // the planes of the chunk's images as one index space of 16-byte loads: every thread stays busy
    const int hw4 = HW >> 2, total = (b1 - b0) * hw4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = threadIdx.x; t < total; t += TPB) {
      const int b = b0 + t / hw4, q4 = t - (t / hw4) * hw4;
      const float4 v = __ldg(reinterpret_cast<const float4 *>(dy + ((size_t)b * n + p) * HW) + q4);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    s = (a.x + a.y) + (a.z + a.w);
  } else {
    for (int b = b0; b < b1; ++b) {
      const float *src = dy + ((size_t)b * n + p) * HW;
      for (int q = threadIdx.x; q < HW; q += TPB) s += __ldg(src + q);
    }
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
  // (the host checks that the input has fewer than 2^31 elements: 32-bit index arithmetic)
  const unsigned tot = (unsigned)total, oHW = (unsigned)(oW * oH);
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += gridDim.x * blockDim.x) {
    const unsigned plane = t / oHW, r = t - plane * oHW;
    const unsigned oy = r / (unsigned)oW, ox = r - oy * (unsigned)oW;
    const unsigned base = plane * (unsigned)(H * W) + (oy * sh) * W + ox * sw;
    float best = __ldg(x + base);
    unsigned pos = base;
    for (int i = 0; i < kh; ++i)
      for (int j = 0; j < kw; ++j) {
        const unsigned q = base + i * W + j;
        const float v = __ldg(x + q);
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
  const unsigned tot = (unsigned)total_in, HWu = (unsigned)(H * W);
  if (kh == sh && kw == sw) {
    // windows do not overlap: every input element belongs to at most one window
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += gridDim.x * blockDim.x) {
      const unsigned plane = t / HWu, r = t - plane * HWu;
      const unsigned yy = r / (unsigned)W, xx = r - yy * (unsigned)W;
      const unsigned oy = yy / (unsigned)kh, ox = xx / (unsigned)kw;
      float s = 0.0f;
      if (oy < (unsigned)oH && ox < (unsigned)oW) {
        const unsigned o = (plane * oH + oy) * oW + ox;
        if ((unsigned)__ldg(argmax + o) == t) s = __ldg(dy + o);
      }
      dx[t] = s;
    }
    return;
  }
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
  if (conv_tc_applicable(ctx, B, C, H, W, n, kh, kw, sh, sw)) return conv_tc_fwd(ctx, B, C, H, W, n, kh, kw, sh, sw, x, w, bias, act, y);
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
  if (conv_tc_applicable(ctx, B, C, H, W, n, kh, kw, sh, sw)) return conv_tc_bwd_data(ctx, B, C, H, W, n, kh, kw, sh, sw, dy, w, dx);
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
  if (conv_tc_applicable(ctx, B, C, H, W, n, kh, kw, sh, sw)) {
    int st = conv_tc_bwd_weight(ctx, B, C, H, W, n, kh, kw, sh, sw, dy, x, scale, beta, dw);
    if (st) return st;
    if (db) return b200_conv_bias_grad(ctx, B, n, oH * oW, dy, scale, beta, db);
    return B200_OK;
  }
  const int CK = C * kh * kw;
  if (CK <= 32 && ((n + 3) / 4) * WS_SLICES <= 4 * 32) {
    // short window: warps own four planes and a slice of the pixels (conv_wgrad_small_kernel)
    const int n4 = (n + 3) & ~3, P = oH * oW;
    const size_t smem_s = ((size_t)((C * H * W + 3) & ~3) + (size_t)P * n4) * sizeof(float);
    if (smem_s <= SMEM_CAP) {
      int nchunks = 2 * ctx->sm_count;
      if (nchunks > B) nchunks = B;
      const int ipc = (B + nchunks - 1) / nchunks;
      nchunks = (B + ipc - 1) / ipc;
      float *partial = (float *)b200_scratch(ctx, (size_t)nchunks * WS_SLICES * n * CK * sizeof(float));
      if (!partial) { b200_set_error("scratch allocation failed"); return B200_ERR_ALLOC; }
      int st = set_smem(conv_wgrad_small_kernel, smem_s);
      if (st) return st;
      // one warp per (four planes, pixel slice) pair, up to 32 warps
      int threads = 32 * ((n4 >> 2) * WS_SLICES);
      if (threads > 1024) threads = 1024;
      conv_wgrad_small_kernel<<<nchunks, threads, smem_s, ctx->stream>>>(B, C, H, W, n, kh, kw, sh, sw, oH, oW, ipc, dy, x, partial);
      LAUNCH_CHECK(ctx);
      const size_t count = (size_t)n * CK;
      chunk_reduce_kernel<<<blocks_for(count * 32, ctx->sm_count), TPB, 0, ctx->stream>>>(nchunks * WS_SLICES, count, partial, scale, beta, dw);
      LAUNCH_CHECK(ctx);
      if (db) return b200_conv_bias_grad(ctx, B, n, oH * oW, dy, scale, beta, db);
      return B200_OK;
    }
  }
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
  chunk_reduce_kernel<<<blocks_for(count * 32, ctx->sm_count), TPB, 0, ctx->stream>>>(nchunks, count, partial, scale, beta,
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
  ARG_CHECK(total < (size_t)INT32_MAX, "tensor too large for int32 positions");
  maxpool_bwd_kernel<<<blocks_for(total, ctx->sm_count), TPB, 0, ctx->stream>>>(total, H, W, kh, kw, sh, sw, oH, oW,
                                                                             dy, argmax, dx);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
