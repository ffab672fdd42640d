// Convolution on the tensor cores: the three contractions of ConvolutionANNComponent
// (ann/ann/c_src/convolution_component.cc:135-218 forward, :221-297 data gradient, :300-354 weight gradient) as
// ONE tcgen05 contraction each instead of the reference's GEMM per output pixel.
//
// The reference flattens every sliding window into a row [planes*kh*kw] (getRewrappedMatrix, :175) and multiplies
// the [bunch, C*kh*kw] window matrix by W^T once per pixel.  Here the window matrix of ALL pixels is built once
// (im2col: rows = (image, pixel) in raster order, columns in the reference's (plane, row, col) order), so that
//   forward        Yr[M, n]    = col[M, K] . W[n, K]^T                 M = B*oH*oW, K = C*kh*kw
//   data gradient  dcol[M, K]  = dYr[M, n] . W[n, K]       then col2im (gather form: every input element sums the
//                                                          windows that cover it -- the reference's beta = 1 loop, :265)
//   weight grad.   dW[n, K]   += scale * dYr[M, n]^T . col[M, K]       a 2-tile problem with a contraction over all
//                                                          pixels: split into as many k slices as SMs (reduce-add)
// run through the same gemm_tc kernels as the dense layers.  Yr / dYr are the pixel-major ("NHWC") forms of the
// NCHW tensors of the API; two small transposition kernels convert, the forward one fused with bias + activation.
// Temporaries live in the per-branch scratch of the context (never freed while captured graphs may reference them).
//
// Used in TF32 mode for K >= 32 (C4's second convolution: K = 400, a 32768 x 32 x 400 contraction per pass); small
// windows (C4's first convolution, K = 25) and the fp32 parity mode stay on the direct FFMA kernels of conv.cu.
#include "common.cuh"

namespace {

constexpr int TPB = 256;

inline int blocks_for(size_t n, int sm_count, int per_sm = 8) {
  size_t b = (n + TPB - 1) / TPB, cap = (size_t)sm_count * per_sm;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

// col[m, k] for k < K, 0 for K <= k < Kp ; m = (b*oH + oy)*oW + ox ; k = (c*kh + i)*kw + j
// One CTA per image (grid-stride): the image and the table of window offsets sit in shared memory, every warp
// builds one row at a time and writes it with 16-byte streaming stores (512 contiguous bytes per warp pass).
// (Gathering straight from global memory fetched one 32-byte sector per 4-byte element: 380 MB of L2 traffic for
// a 4.7 MB input.)
__global__ void __launch_bounds__(TPB) im2col_kernel(int B, int C, int H, int W, int kh, int kw, int sh, int sw, int oH,
                                                     int oW, int K, int Kp, const float *__restrict__ x, float *__restrict__ col) {
  extern __shared__ int smem_i[];
  int *koff = smem_i;                                         // [Kp]: offset of window element k (-1: padding)
  float *img = reinterpret_cast<float *>(smem_i + Kp);        // [C*H*W]
  for (int k = threadIdx.x; k < Kp; k += TPB) {
    const int j = k % kw, i = (k / kw) % kh, c = k / (kw * kh);
    koff[k] = k < K ? (c * H + i) * W + j : -1;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kp4 = Kp >> 2, P = oH * oW, CHW = C * H * W;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float *xb = x + (size_t)b * CHW;
    for (int q = threadIdx.x; q < CHW; q += TPB) img[q] = __ldg(xb + q);
    __syncthreads();
    for (int p = warp; p < P; p += TPB / 32) {
      const int oy = p / oW, ox = p - oy * oW;
      const float *xw = img + (oy * sh) * W + ox * sw;
      float4 *dst = reinterpret_cast<float4 *>(col + ((size_t)b * P + p) * Kp);
      for (int f = lane; f < kp4; f += 32) {
        const int4 o = reinterpret_cast<const int4 *>(koff)[f];
        float4 v;
        v.x = o.x >= 0 ? xw[o.x] : 0.0f;
        v.y = o.y >= 0 ? xw[o.y] : 0.0f;
        v.z = o.z >= 0 ? xw[o.z] : 0.0f;
        v.w = o.w >= 0 ? xw[o.w] : 0.0f;
        __stcs(dst + f, v);
      }
    }
  }
}

// dx[b, c, y, x] = sum over the windows (oy, ox) and offsets (i, j) with oy*sh + i == y, ox*sw + j == x of
// dcol[(b, oy, ox), (c, i, j)].  One CTA per image: the dcol rows of the image are staged in shared memory with
// coalesced loads (a direct gather would touch one 4-byte word per 32-byte sector).  UNIT: stride 1 in both
// directions (no divisibility tests in the inner loops).
template <bool UNIT>
__global__ void __launch_bounds__(512) col2im_kernel(int C, int H, int W, int kh, int kw, int sh, int sw, int oH, int oW, int K,
                                                     int Kp, const float *__restrict__ dcol, float *__restrict__ dx,
                                                     int pix_per_pass) {
  extern __shared__ float sm[];     // [pix_per_pass][Kp + 1]: the odd row pitch spreads consecutive pixels over the banks
  const int b = blockIdx.x, P = oH * oW;
  const int CHW = C * H * W, HW = H * W, LDS = Kp + 1;
  float *dxb = dx + (size_t)b * CHW;
  // when the whole image does not fit, windows are processed in passes of pix_per_pass pixels and dx accumulates
  for (int p0 = 0; p0 < P; p0 += pix_per_pass) {
    const int np = min(pix_per_pass, P - p0);
    __syncthreads();
    const float4 *src = reinterpret_cast<const float4 *>(dcol + ((size_t)b * P + p0) * Kp);
    for (int t = threadIdx.x; t < np * (Kp >> 2); t += 512) {
      const float4 v = __ldcs(src + t);
      const int row = (t << 2) / Kp, cc = (t << 2) - row * Kp;
      float *d = sm + row * LDS + cc;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < CHW; q += 512) {
      const int c = q / HW, r = q - c * HW;
      const int yy = r / W, xx = r - yy * W;
      float s = (p0 == 0) ? 0.0f : dxb[q];
      if (UNIT) {
        // windows (oy, ox) = (yy - i, xx - j) that exist.  The address of tap (i, j) is
        //   ((yy - i)*oW + (xx - j) - p0)*LDS + (c*kh + i)*kw + j  =  a0 - i*di - j*dj
        const int a0 = ((yy * oW + xx) - p0) * LDS + c * kh * kw;
        const int di = oW * LDS - kw, dj = LDS - 1;
        const int i_lo = max(0, yy - (oH - 1)), i_hi = min(kh - 1, yy);
        const int j_lo = max(0, xx - (oW - 1)), j_hi = min(kw - 1, xx);
        if (np == P) {
          // the whole image is resident: no per-tap range test on the pixel
          for (int i = i_lo; i <= i_hi; ++i) {
            const float *row = sm + a0 - i * di;
#pragma unroll 5
            for (int j = j_lo; j <= j_hi; ++j) s += row[-j * dj];
          }
        } else {
          for (int i = i_lo; i <= i_hi; ++i)
            for (int j = j_lo; j <= j_hi; ++j) {
              const int p = (yy - i) * oW + (xx - j) - p0;
              if (p >= 0 && p < np) s += sm[a0 - i * di - j * dj];
            }
        }
      } else {
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
            const int p = oy * oW + ox - p0;
            if (p < 0 || p >= np) continue;
            s += sm[p * LDS + (c * kh + i) * kw + j];
          }
        }
      }
      dxb[q] = s;
    }
  }
}

// y[b, o, p] = act(yr[b*P + p, o] + bias[o])     (pixel-major rows -> NCHW, 32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(TPB) rows_to_nchw_kernel(int P, int n, int ldr, const float *__restrict__ yr,
                                                           const float *__restrict__ bias, int act, float *__restrict__ y) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows of 32
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, o = o0 + tx;
    tile[r][tx] = (p < P && o < n) ? __ldg(yr + ((size_t)b * P + p) * ldr + o) : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int o = o0 + r, p = p0 + tx;
    if (o < n && p < P) {
      float v = tile[tx][r];
      if (bias) v += __ldg(bias + o);
      if (act != B200_ACT_NONE) v = act_apply(act, v);
      y[((size_t)b * n + o) * P + p] = v;
    }
  }
}
// dyr[b*P + p, o] = dy[b, o, p] ; columns n <= o < ldr are zero-filled
__global__ void __launch_bounds__(TPB) nchw_to_rows_kernel(int P, int n, int ldr, const float *__restrict__ dy,
                                                           float *__restrict__ dyr) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int o = o0 + r, p = p0 + tx;
    tile[r][tx] = (o < n && p < P) ? __ldg(dy + ((size_t)b * n + o) * P + p) : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, o = o0 + tx;
    if (p < P && o < ldr) dyr[((size_t)b * P + p) * ldr + o] = tile[tx][r];
  }
}
// wp[o, k] = w[o, k] for o < n, k < K, 0 elsewhere up to [n4, Kp] (weights with a row pitch TMA accepts and with
// the rows the zero-padded planes of dyr multiply)
__global__ void __launch_bounds__(TPB) pad_rows_kernel(int n, int n4, int K, int Kp, const float *__restrict__ w,
                                                       float *__restrict__ wp) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n4 * Kp; t += gridDim.x * blockDim.x) {
    const int o = t / Kp, k = t % Kp;
    wp[t] = (o < n && k < K) ? __ldg(w + (size_t)o * K + k) : 0.0f;
  }
}
// dw[o, k] = beta*dw[o, k] + acc[o, k]   (acc has row pitch Kp and already carries the scale)
__global__ void __launch_bounds__(TPB) unpad_axpby_kernel(int n, int K, int Kp, const float *__restrict__ acc, float beta,
                                                          float *__restrict__ dw) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * K; t += gridDim.x * blockDim.x) {
    const int o = t / K, k = t % K;
    dw[t] = (beta != 0.0f ? beta * dw[t] : 0.0f) + acc[(size_t)o * Kp + k];
  }
}

struct Geo {
  int B, C, H, W, n, kh, kw, sh, sw, oH, oW, P, K, Kp, n4;
  size_t M;
};
inline Geo geo(int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw) {
  Geo g;
  g.B = B; g.C = C; g.H = H; g.W = W; g.n = n; g.kh = kh; g.kw = kw; g.sh = sh; g.sw = sw;
  g.oH = (H - kh) / sh + 1;
  g.oW = (W - kw) / sw + 1;
  g.P = g.oH * g.oW;
  g.K = C * kh * kw;
  g.Kp = (g.K + 3) & ~3;
  g.n4 = (n + 3) & ~3;
  g.M = (size_t)B * g.P;
  return g;
}
inline size_t al(size_t floats) { return (floats + 127) & ~size_t(127); }   // 512-byte aligned carve-outs

int launch_im2col(b200_ctx *ctx, const Geo &g, const float *x, float *col) {
  const size_t smem = ((size_t)g.Kp + (size_t)g.C * g.H * g.W) * sizeof(float);
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) { b200_set_error("conv_tc: image of %d values does not fit in shared memory", g.C * g.H * g.W); return B200_ERR_UNSUPPORTED; }
    CUDA_TRY(cudaFuncSetAttribute(im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int blocks = g.B < ctx->sm_count * 8 ? g.B : ctx->sm_count * 8;
  im2col_kernel<<<blocks, TPB, smem, ctx->stream>>>(g.B, g.C, g.H, g.W, g.kh, g.kw, g.sh, g.sw, g.oH, g.oW, g.K, g.Kp, x, col);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
int padded_weights(b200_ctx *ctx, const Geo &g, const float *w, float *wp_buf, const float **wp) {
  if (g.Kp == g.K && g.n4 == g.n && (((uintptr_t)w) & 15) == 0) { *wp = w; return B200_OK; }
  pad_rows_kernel<<<blocks_for((size_t)g.n4 * g.Kp, ctx->sm_count), TPB, 0, ctx->stream>>>(g.n, g.n4, g.K, g.Kp, w, wp_buf);
  LAUNCH_CHECK(ctx);
  *wp = wp_buf;
  return B200_OK;
}

}  // namespace

// the tensor-core path takes convolutions whose window is long enough for a contraction to pay
bool conv_tc_applicable(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw) {
  if (ctx->math_mode != B200_MATH_TF32) return false;
  const Geo g = geo(B, C, H, W, n, kh, kw, sh, sw);
  return g.K >= 32 && g.n >= 8 && g.M >= 1024 && g.M * (size_t)g.Kp < ((size_t)1 << 31) && (size_t)g.C * g.H * g.W < ((size_t)1 << 24);
}

int conv_tc_fwd(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw, const float *x, const float *w,
                const float *bias, int act, float *y) {
  const Geo g = geo(B, C, H, W, n, kh, kw, sh, sw);
  float *scr = (float *)b200_scratch(ctx, (al(g.M * g.Kp) + al((size_t)g.n4 * g.Kp) + al(g.M * g.n4)) * sizeof(float));
  if (!scr) { b200_set_error("conv_tc_fwd: scratch allocation failed"); return B200_ERR_ALLOC; }
  float *col = scr, *wp_buf = col + al(g.M * g.Kp), *yr = wp_buf + al((size_t)g.n4 * g.Kp);
  int st = launch_im2col(ctx, g, x, col);
  if (st) return st;
  const float *wp;
  st = padded_weights(ctx, g, w, wp_buf, &wp);
  if (st) return st;
  GemmEpilogue ep;
  st = gemm_dispatch(ctx, 0, 1, (int)g.M, g.n, g.K, col, g.Kp, wp, g.Kp, yr, g.n4, ep);
  if (st) return st;
  dim3 grid((g.P + 31) / 32, (g.n + 31) / 32, g.B);
  rows_to_nchw_kernel<<<grid, TPB, 0, ctx->stream>>>(g.P, g.n, g.n4, yr, bias, act == B200_ACT_LINEAR ? B200_ACT_NONE : act, y);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

int conv_tc_bwd_data(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw, const float *dy,
                     const float *w, float *dx) {
  const Geo g = geo(B, C, H, W, n, kh, kw, sh, sw);
  float *scr = (float *)b200_scratch(ctx, (al(g.M * g.Kp) + al((size_t)g.n4 * g.Kp) + al(g.M * g.n4)) * sizeof(float));
  if (!scr) { b200_set_error("conv_tc_bwd_data: scratch allocation failed"); return B200_ERR_ALLOC; }
  float *dcol = scr, *wp_buf = dcol + al(g.M * g.Kp), *dyr = wp_buf + al((size_t)g.n4 * g.Kp);
  dim3 grid((g.P + 31) / 32, (g.n4 + 31) / 32, g.B);
  nchw_to_rows_kernel<<<grid, TPB, 0, ctx->stream>>>(g.P, g.n, g.n4, dy, dyr);
  LAUNCH_CHECK(ctx);
  const float *wp;
  int st = padded_weights(ctx, g, w, wp_buf, &wp);
  if (st) return st;
  GemmEpilogue ep;
  // dcol[M, K] = dyr[M, n] . wp[n, K]   (contraction over the n output planes; the padded columns n..n4 are zero)
  st = gemm_dispatch(ctx, 0, 0, (int)g.M, g.Kp, g.n4, dyr, g.n4, wp, g.Kp, dcol, g.Kp, ep);
  if (st) return st;
  size_t smem = (size_t)g.P * (g.Kp + 1) * sizeof(float);
  int pix = g.P;
  const size_t cap = 104 * 1024;          // two CTAs per SM
  if (smem > cap) {
    pix = (int)(cap / ((size_t)(g.Kp + 1) * sizeof(float)));
    if (pix < 1) { b200_set_error("conv_tc_bwd_data: window of %d values does not fit in shared memory", g.Kp); return B200_ERR_UNSUPPORTED; }
    smem = (size_t)pix * (g.Kp + 1) * sizeof(float);
  }
  if (ONCE_PER_DEVICE(ctx)) {
    CUDA_TRY(cudaFuncSetAttribute(col2im_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
    CUDA_TRY(cudaFuncSetAttribute(col2im_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
  }
  if (g.sh == 1 && g.sw == 1)
    col2im_kernel<true><<<g.B, 512, smem, ctx->stream>>>(g.C, g.H, g.W, g.kh, g.kw, g.sh, g.sw, g.oH, g.oW, g.K, g.Kp, dcol, dx, pix);
  else
    col2im_kernel<false><<<g.B, 512, smem, ctx->stream>>>(g.C, g.H, g.W, g.kh, g.kw, g.sh, g.sw, g.oH, g.oW, g.K, g.Kp, dcol, dx, pix);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

int conv_tc_bwd_weight(b200_ctx *ctx, int B, int C, int H, int W, int n, int kh, int kw, int sh, int sw, const float *dy,
                       const float *x, float scale, float beta, float *dw) {
  const Geo g = geo(B, C, H, W, n, kh, kw, sh, sw);
  float *scr = (float *)b200_scratch(ctx, (al(g.M * g.Kp) + al((size_t)g.n4 * g.Kp) + al(g.M * g.n4)) * sizeof(float));
  if (!scr) { b200_set_error("conv_tc_bwd_weight: scratch allocation failed"); return B200_ERR_ALLOC; }
  float *col = scr, *acc = col + al(g.M * g.Kp), *dyr = acc + al((size_t)g.n4 * g.Kp);
  int st = launch_im2col(ctx, g, x, col);
  if (st) return st;
  dim3 grid((g.P + 31) / 32, (g.n4 + 31) / 32, g.B);
  nchw_to_rows_kernel<<<grid, TPB, 0, ctx->stream>>>(g.P, g.n, g.n4, dy, dyr);
  LAUNCH_CHECK(ctx);
  // acc[n4, Kp] = 0, then += scale * dyr^T . col with TMA reduce-add stores: the contraction over all M pixels is cut
  // into k slices that fill the device
  st = b200_memset_zero(ctx, acc, (size_t)g.n4 * g.Kp * sizeof(float));
  if (st) return st;
  GemmEpilogue ep;
  ep.alpha = scale;
  ep.beta = 1.0f;
  st = gemm_dispatch(ctx, 1, 0, g.n4, g.Kp, (int)g.M, dyr, g.n4, col, g.Kp, acc, g.Kp, ep);
  if (st) return st;
  unpad_axpby_kernel<<<blocks_for((size_t)g.n * g.K, ctx->sm_count), TPB, 0, ctx->stream>>>(g.n, g.K, g.Kp, acc, beta, dw);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
