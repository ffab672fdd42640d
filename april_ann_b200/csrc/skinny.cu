// Dense layers with a handful of output neurons (N <= 16: the 10-class output layer of the
// BASELINE MLPs) and the bias gradient of any layer.  These contractions are HBM-bound
// (2*N flops per activation element read), so they run on the CUDA cores with exact fp32 FMAs,
// streaming the activation matrix once with 16-byte accesses; a 128-row tensor-core tile would be
// >90 % padding and the generic FFMA tile kernel leaves most SMs idle.
//
//   forward        Y[M,N]  = act( X[M,K] . W[N,K]^T + b )          reads X once
//   data gradient  dX[M,K] = ( dY[M,N] . W[N,K] ) (.) act'(Yprev)  writes dX once
//   weight grad.   dW[N,K] = beta*dW + scale * dY^T . X            reads X once   (+ db)
//   bias gradient  db[N]   = beta*db + scale * sum_m dY[m,:]       reads dY once
// (dot_product_component.cc:63-98,123-152,194-216 ; bias_component.cc:87-122).
// Reductions over the bunch are two-stage and deterministic: per-row-chunk partials in the
// context scratch, then one fixed-order sum.
#include "common.cuh"

namespace {

constexpr int SK_MAXN = 16;

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float dot4(const float4 &a, const float4 &b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}

// ---------------------------------------------------------------- forward
// one warp per RPW rows; lanes stride over K (float4 when VEC), W comes from L1/L2
template <int NT, int RPW, bool VEC>
__global__ void __launch_bounds__(128) skinny_fwd_kernel(int M, int N, int K, const float *__restrict__ X, int ldx,
                                                         const float *__restrict__ W, int ldw,
                                                         const float *__restrict__ bias, int act, float *__restrict__ Y,
                                                         int ldy) {
  const int lane = threadIdx.x & 31;
  const int m0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  if (m0 >= M) return;
  float acc[RPW][NT];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[r][n] = 0.0f;
  if (VEC) {
    for (int k = lane * 4; k < K; k += 128) {
      float4 xv[RPW];
#pragma unroll
      for (int r = 0; r < RPW; ++r)
        xv[r] = (m0 + r < M) ? ldg4(X + (size_t)(m0 + r) * ldx + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (n < N) {
          const float4 wv = ldg4(W + (size_t)n * ldw + k);
#pragma unroll
          for (int r = 0; r < RPW; ++r) acc[r][n] = dot4(xv[r], wv, acc[r][n]);
        }
      }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      float xv[RPW];
#pragma unroll
      for (int r = 0; r < RPW; ++r) xv[r] = (m0 + r < M) ? __ldg(X + (size_t)(m0 + r) * ldx + k) : 0.0f;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (n < N) {
          const float wv = __ldg(W + (size_t)n * ldw + k);
#pragma unroll
          for (int r = 0; r < RPW; ++r) acc[r][n] = fmaf(xv[r], wv, acc[r][n]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    float mine = 0.0f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float s = warp_sum(acc[r][n]);
      if (lane == n) mine = s;
    }
    if (lane < N && m0 + r < M) {
      float v = mine;
      if (bias) v += __ldg(bias + lane);
      Y[(size_t)(m0 + r) * ldy + lane] = act_apply(act, v);
    }
  }
}

// ---------------------------------------------------------------- forward, lane = row
// CTA = 32 bunch rows x 16 warps; W is staged once per CTA in shared memory as [K/4][NT] float4 so
// that every read is a warp-wide broadcast; each warp owns 1/16 of K and each lane one row, so X is
// streamed with 8 independent 16-byte loads per lane in flight and no shuffle reduction is needed.
// Partial sums of the 16 K-slices meet in shared memory.
template <int NT>
__global__ void __launch_bounds__(512) skinny_fwd_rows_kernel(int M, int N, int K, const float *__restrict__ X, int ldx,
                                                              const float *__restrict__ W, int ldw,
                                                              const float *__restrict__ bias, int act,
                                                              float *__restrict__ Y, int ldy) {
  extern __shared__ float4 sk_smem[];
  const int K4 = K >> 2;
  float4 *sW = sk_smem;                                          // [K4][NT]
  float *red = reinterpret_cast<float *>(sk_smem + (size_t)K4 * NT);   // [16][32][NT]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int idx = tid; idx < NT * K4; idx += 512) {
    const int n = idx / K4, k4 = idx - n * K4;
    sW[(size_t)k4 * NT + n] = (n < N) ? ldg4(W + (size_t)n * ldw + 4 * k4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int r0 = blockIdx.x * 32;
  const int row = min(r0 + lane, M - 1);
  const float *xrow = X + (size_t)row * ldx;
  const int per = (K4 + 15) >> 4;
  const int kb = w * per, ke = min(K4, kb + per);
  float acc[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc[n] = 0.0f;
  int k4 = kb;
  for (; k4 + 8 <= ke; k4 += 8) {
    float4 xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) xv[u] = ldg4(xrow + 4 * (k4 + u));
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[n] = dot4(xv[u], sW[(size_t)(k4 + u) * NT + n], acc[n]);
  }
  for (; k4 < ke; ++k4) {
    const float4 xv = ldg4(xrow + 4 * k4);
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = dot4(xv, sW[(size_t)k4 * NT + n], acc[n]);
  }
#pragma unroll
  for (int n = 0; n < NT; ++n) red[((size_t)w * 32 + lane) * NT + n] = acc[n];
  __syncthreads();
  for (int idx = tid; idx < 32 * N; idx += 512) {
    const int r = idx / N, n = idx - r * N;
    if (r0 + r >= M) continue;
    float v = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) v += red[((size_t)i * 32 + r) * NT + n];
    if (bias) v += __ldg(bias + n);
    Y[(size_t)(r0 + r) * ldy + n] = act_apply(act, v);
  }
}

// ---------------------------------------------------------------- data gradient
// thread = 4 consecutive input features (1 when !VEC); blockIdx.y = chunk of ROWS bunch rows
template <int NT, int ROWS, bool VEC>
__global__ void __launch_bounds__(128) skinny_bwd_data_kernel(int M, int N, int K, const float *__restrict__ dY, int lddy,
                                                              const float *__restrict__ W, int ldw, int dact,
                                                              const float *__restrict__ Yprev, int ldyp,
                                                              float *__restrict__ dX, int lddx) {
  __shared__ float sdy[ROWS][NT];
  constexpr int VW = VEC ? 4 : 1;
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) * VW;
  const int m0 = blockIdx.y * ROWS;
  for (int i = threadIdx.x; i < ROWS * NT; i += blockDim.x) {
    const int r = i / NT, n = i % NT;
    sdy[r][n] = (m0 + r < M && n < N) ? __ldg(dY + (size_t)(m0 + r) * lddy + n) : 0.0f;
  }
  __syncthreads();
  if (k >= K) return;
  float w[NT][VW];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    if (VEC) {
      const float4 t = (n < N) ? ldg4(W + (size_t)n * ldw + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      w[n][0] = t.x; w[n][VW > 1 ? 1 : 0] = t.y; w[n][VW > 2 ? 2 : 0] = t.z; w[n][VW > 3 ? 3 : 0] = t.w;
    } else {
      w[n][0] = (n < N) ? __ldg(W + (size_t)n * ldw + k) : 0.0f;
    }
  }
  const int rows = min(ROWS, M - m0);
#pragma unroll 4
  for (int r = 0; r < rows; ++r) {
    float a[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) a[e] = 0.0f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float d = sdy[r][n];
#pragma unroll
      for (int e = 0; e < VW; ++e) a[e] = fmaf(d, w[n][e], a[e]);
    }
    const size_t m = (size_t)(m0 + r);
    if (dact != B200_ACT_NONE) {
      if (VEC) {
        const float4 y = ldg4(Yprev + m * ldyp + k);
        a[0] *= act_deriv_from_output(dact, y.x);
        a[VW > 1 ? 1 : 0] *= act_deriv_from_output(dact, y.y);
        a[VW > 2 ? 2 : 0] *= act_deriv_from_output(dact, y.z);
        a[VW > 3 ? 3 : 0] *= act_deriv_from_output(dact, y.w);
      } else {
        a[0] *= act_deriv_from_output(dact, __ldg(Yprev + m * ldyp + k));
      }
    }
    if (VEC) *reinterpret_cast<float4 *>(dX + m * lddx + k) = make_float4(a[0], a[VW > 1 ? 1 : 0], a[VW > 2 ? 2 : 0], a[VW > 3 ? 3 : 0]);
    else dX[m * lddx + k] = a[0];
  }
}

// ---------------------------------------------------------------- weight gradient, stage 1
// CTA = 8 warps over one slab of 32*VW input features and one chunk of ROWS bunch rows: warp w takes
// rows w, w+8, ...; lanes take VW consecutive features each (coalesced rows of X).  The 8 warps'
// sums meet in shared memory; the chunk's partial goes to part[chunk][N*K], the bias partial (from
// the first slab) to part_b[chunk][N].
template <int NT, int ROWS, bool VEC>
__global__ void __launch_bounds__(256) skinny_wgrad_partial_kernel(int M, int N, int K, const float *__restrict__ dY,
                                                                   int lddy, const float *__restrict__ X, int ldx,
                                                                   float *__restrict__ part, float *__restrict__ part_b) {
  constexpr int VW = VEC ? 4 : 1;
  constexpr int COLS = 32 * VW;
  __shared__ float sdy[ROWS][NT];
  __shared__ float red[4][NT][COLS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int k = blockIdx.x * COLS + lane * VW;
  const int m0 = blockIdx.y * ROWS;
  for (int i = threadIdx.x; i < ROWS * NT; i += 256) {
    const int r = i / NT, n = i % NT;
    sdy[r][n] = (m0 + r < M && n < N) ? __ldg(dY + (size_t)(m0 + r) * lddy + n) : 0.0f;
  }
  __syncthreads();
  if (blockIdx.x == 0 && part_b && threadIdx.x < N) {
    float s = 0.0f;
    for (int r = 0; r < ROWS; ++r) s += sdy[r][threadIdx.x];
    part_b[(size_t)blockIdx.y * N + threadIdx.x] = s;
  }
  float acc[NT][VW];
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int e = 0; e < VW; ++e) acc[n][e] = 0.0f;
  if (k < K) {
    const int rows = min(ROWS, M - m0);
    int r = w;
    for (; r + 56 < rows; r += 64) {      // 8 rows of this warp in flight
      float x[8][VW];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (VEC) {
          const float4 t = ldg4(X + (size_t)(m0 + r + 8 * u) * ldx + k);
          x[u][0] = t.x; x[u][VW > 1 ? 1 : 0] = t.y; x[u][VW > 2 ? 2 : 0] = t.z; x[u][VW > 3 ? 3 : 0] = t.w;
        } else {
          x[u][0] = __ldg(X + (size_t)(m0 + r + 8 * u) * ldx + k);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const float d = sdy[r + 8 * u][n];
#pragma unroll
          for (int e = 0; e < VW; ++e) acc[n][e] = fmaf(d, x[u][e], acc[n][e]);
        }
    }
    for (; r < rows; r += 8) {
      float x[VW];
      if (VEC) {
        const float4 t = ldg4(X + (size_t)(m0 + r) * ldx + k);
        x[0] = t.x; x[VW > 1 ? 1 : 0] = t.y; x[VW > 2 ? 2 : 0] = t.z; x[VW > 3 ? 3 : 0] = t.w;
      } else {
        x[0] = __ldg(X + (size_t)(m0 + r) * ldx + k);
      }
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const float d = sdy[r][n];
#pragma unroll
        for (int e = 0; e < VW; ++e) acc[n][e] = fmaf(d, x[e], acc[n][e]);
      }
    }
  }
  // fixed-order tree over the 8 warps (4+4 -> 2+2 -> 1+1): deterministic
#pragma unroll
  for (int h = 4; h >= 1; h >>= 1) {
    if (w >= h && w < 2 * h) {
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < VW; ++e) red[w - h][n][lane * VW + e] = acc[n][e];
    }
    __syncthreads();
    if (w < h) {
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < VW; ++e) acc[n][e] += red[w][n][lane * VW + e];
    }
    __syncthreads();
  }
  if (w == 0 && k < K) {
    float *dst = part + (size_t)blockIdx.y * N * K;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (n < N) {
        if (VEC) *reinterpret_cast<float4 *>(dst + (size_t)n * K + k) = make_float4(acc[n][0], acc[n][VW > 1 ? 1 : 0], acc[n][VW > 2 ? 2 : 0], acc[n][VW > 3 ? 3 : 0]);
        else dst[(size_t)n * K + k] = acc[n][0];
      }
    }
  }
}

// ---------------------------------------------------------------- column sums, stage 1
// part[chunk][n] = sum over the chunk's rows of dy[m,n]; a warp covers 128 (VEC) or 32 columns
template <int ROWS, bool VEC>
__global__ void __launch_bounds__(256) colsum_partial_kernel(int M, int N, const float *__restrict__ dy, int ld,
                                                             float *__restrict__ part) {
  constexpr int VW = VEC ? 4 : 1;
  __shared__ float sm[8][32 * VW + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 * VW + lane * VW;
  const int m0 = blockIdx.y * ROWS, m1 = min(M, m0 + ROWS);
  float s[VW];
#pragma unroll
  for (int e = 0; e < VW; ++e) s[e] = 0.0f;
  if (n < N) {
#pragma unroll 4
    for (int m = m0 + w; m < m1; m += 8) {
      if (VEC) {
        const float4 t = ldg4(dy + (size_t)m * ld + n);
        s[0] += t.x; s[VW > 1 ? 1 : 0] += t.y; s[VW > 2 ? 2 : 0] += t.z; s[VW > 3 ? 3 : 0] += t.w;
      } else {
        s[0] += __ldg(dy + (size_t)m * ld + n);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < VW; ++e) sm[w][lane * VW + e] = s[e];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VW; c += 256) {
    const int col = blockIdx.x * 32 * VW + c;
    if (col < N) {
      float t = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += sm[i][c];
      part[(size_t)blockIdx.y * N + col] = t;
    }
  }
}

// ---------------------------------------------------------------- stage 2 (shared)
// out[i] = beta*out[i] + scale * sum_c part[c][i]   for up to two jobs (blockIdx.y)
struct ReduceJob {
  const float *part;
  float *out;
  int n;
  int ld_out, cols;   // out index = (i / cols) * ld_out + i % cols
};
__global__ void __launch_bounds__(256) reduce_partials_kernel(int nchunks, float scale, float beta, ReduceJob j0, ReduceJob j1) {
  const ReduceJob j = blockIdx.y == 0 ? j0 : j1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.n) return;
  float s = 0.0f;
  int c = 0;
  for (; c + 8 <= nchunks; c += 8) {     // 8 independent loads in flight, summed in chunk order
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(j.part + (size_t)(c + u) * j.n + i);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; c < nchunks; ++c) s += __ldg(j.part + (size_t)c * j.n + i);
  float *o = j.out + (size_t)(i / j.cols) * j.ld_out + (i % j.cols);
  *o = (beta != 0.0f ? beta * *o : 0.0f) + scale * s;
}

inline bool al16(const void *p) { return (((uintptr_t)p) & 15) == 0; }
inline int pad4(int n) { return (n + 3) & ~3; }

}  // namespace

bool skinny_applicable(int M, int N, int K) { return N >= 1 && N <= SK_MAXN && M >= 1 && K >= 1; }

#define NT_DISPATCH(N, ...)                                  \
  switch (pad4(N)) {                                         \
    case 4: { constexpr int NT = 4; __VA_ARGS__; } break;    \
    case 8: { constexpr int NT = 8; __VA_ARGS__; } break;    \
    case 12: { constexpr int NT = 12; __VA_ARGS__; } break;  \
    default: { constexpr int NT = 16; __VA_ARGS__; } break;  \
  }

int skinny_fwd(b200_ctx *ctx, int M, int N, int K, const float *X, int ldx, const float *W, int ldw, const float *bias,
               int act, float *Y, int ldy) {
  const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && (ldw % 4 == 0) && al16(X) && al16(W);
  if (vec) {
    // W resident in shared memory: [K/4][NT] float4 + the 16x32xNT partials
    const int nt = pad4(N);
    const size_t smem = (size_t)(K / 4) * nt * 16 + (size_t)16 * 32 * nt * 4;
    if (smem <= 200 * 1024) {
      const int grid = (M + 31) / 32;
      NT_DISPATCH(N, {
        auto kern = skinny_fwd_rows_kernel<NT>;
        static bool attr = false;
        if (!attr) { CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
        kern<<<grid, 512, smem, ctx->stream>>>(M, N, K, X, ldx, W, ldw, bias, act, Y, ldy);
      });
      LAUNCH_CHECK(ctx);
      return B200_OK;
    }
  }
  constexpr int RPW = 2, WARPS = 4;
  const int grid = (M + RPW * WARPS - 1) / (RPW * WARPS);
  NT_DISPATCH(N, {
    if (vec) skinny_fwd_kernel<NT, RPW, true><<<grid, WARPS * 32, 0, ctx->stream>>>(M, N, K, X, ldx, W, ldw, bias, act, Y, ldy);
    else skinny_fwd_kernel<NT, RPW, false><<<grid, WARPS * 32, 0, ctx->stream>>>(M, N, K, X, ldx, W, ldw, bias, act, Y, ldy);
  });
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

int skinny_bwd_data(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy, const float *W, int ldw, int dact,
                    const float *Yprev, int ldyp, float *dX, int lddx) {
  const bool vec = (K % 4 == 0) && (ldw % 4 == 0) && (lddx % 4 == 0) && al16(W) && al16(dX) &&
                   (dact == B200_ACT_NONE || ((ldyp % 4 == 0) && al16(Yprev)));
  constexpr int ROWS = 16;
  const int kthreads = vec ? K / 4 : K;
  dim3 grid((kthreads + 127) / 128, (M + ROWS - 1) / ROWS);
  NT_DISPATCH(N, {
    if (vec) skinny_bwd_data_kernel<NT, ROWS, true><<<grid, 128, 0, ctx->stream>>>(M, N, K, dY, lddy, W, ldw, dact, Yprev, ldyp, dX, lddx);
    else skinny_bwd_data_kernel<NT, ROWS, false><<<grid, 128, 0, ctx->stream>>>(M, N, K, dY, lddy, W, ldw, dact, Yprev, ldyp, dX, lddx);
  });
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

int skinny_bwd_weight(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy, const float *X, int ldx, float scale,
                      float beta, float *dW, int lddw, float *db) {
  constexpr int ROWS = 128;
  const int chunks = (M + ROWS - 1) / ROWS;
  const size_t nk = (size_t)N * K;
  const size_t part_elems = ((size_t)chunks * nk + 3) & ~size_t(3);   // [chunk][N*K], then the bias partials
  float *part = (float *)b200_scratch(ctx, (part_elems + (size_t)chunks * N + 64) * sizeof(float));
  if (!part) { b200_set_error("skinny_bwd_weight: scratch allocation failed"); return B200_ERR_ALLOC; }
  float *part_b = part + part_elems;
  const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && al16(X);
  dim3 grid((K + (vec ? 127 : 31)) / (vec ? 128 : 32), chunks);
  NT_DISPATCH(N, {
    if (vec) skinny_wgrad_partial_kernel<NT, ROWS, true><<<grid, 256, 0, ctx->stream>>>(M, N, K, dY, lddy, X, ldx, part, db ? part_b : nullptr);
    else skinny_wgrad_partial_kernel<NT, ROWS, false><<<grid, 256, 0, ctx->stream>>>(M, N, K, dY, lddy, X, ldx, part, db ? part_b : nullptr);
  });
  LAUNCH_CHECK(ctx);
  ReduceJob j0{part, dW, (int)nk, lddw, K};
  ReduceJob j1{part_b, db, db ? N : 0, 1, 1};
  dim3 rgrid((unsigned)((nk + 255) / 256), db ? 2 : 1);
  reduce_partials_kernel<<<rgrid, 256, 0, ctx->stream>>>(chunks, scale, beta, j0, j1);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

int colsum_scaled(b200_ctx *ctx, int M, int N, const float *dy, int ld, float scale, float beta, float *out) {
  constexpr int ROWS = 64;
  const int chunks = (M + ROWS - 1) / ROWS;
  float *part = (float *)b200_scratch(ctx, ((size_t)chunks * N + 64) * sizeof(float));
  if (!part) { b200_set_error("colsum: scratch allocation failed"); return B200_ERR_ALLOC; }
  const bool vec = (N % 4 == 0) && (ld % 4 == 0) && al16(dy);
  if (vec) {
    dim3 grid((N + 127) / 128, chunks);
    colsum_partial_kernel<ROWS, true><<<grid, 256, 0, ctx->stream>>>(M, N, dy, ld, part);
  } else {
    dim3 grid((N + 31) / 32, chunks);
    colsum_partial_kernel<ROWS, false><<<grid, 256, 0, ctx->stream>>>(M, N, dy, ld, part);
  }
  LAUNCH_CHECK(ctx);
  ReduceJob j0{part, out, N, 1, 1};
  reduce_partials_kernel<<<dim3((N + 255) / 256, 1), 256, 0, ctx->stream>>>(chunks, scale, beta, j0, j0);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
