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

// ---------------------------------------------------------------- forward, K split over the warps
// CTA = 8 warps x R bunch rows.  Warp w takes the 128-float K chunks w, w+8, ...; a lane loads one
// float4 of each of the R rows (coalesced 512-byte row segments) and of each W row (L1/L2 hits: W is
// a few tens of KB shared by every CTA), so all loads of a chunk are independent and in flight
// together -- at these sizes the kernel is one memory latency long, not bandwidth bound.  The
// R*NT per-lane partial sums of a warp are reduced with a halving butterfly (31 shuffles per 32
// values, lane L ends up with value L), the 8 warps meet in shared memory in fixed order.
template <int V>
__device__ __forceinline__ void butterfly32(float (&v)[V], int base, int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float a = v[base + i], b = v[base + i + o];
      const float keep = up ? b : a, send = up ? a : b;
      v[base + i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

template <int NT, int R>
__global__ void __launch_bounds__(256) skinny_fwd_ksplit_kernel(int M, int N, int K, const float *__restrict__ X, int ldx,
                                                                const float *__restrict__ W, int ldw,
                                                                const float *__restrict__ bias, int act,
                                                                float *__restrict__ Y, int ldy) {
  constexpr int V = R * NT;
  static_assert(V % 32 == 0, "R*NT must be a multiple of 32");
  __shared__ float red[8][V];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int m0 = blockIdx.x * R;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.0f;
#pragma unroll 2
  for (int k = w * 128 + lane * 4; k < K; k += 8 * 128) {
    float4 xv[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
      xv[r] = ldg4(X + (size_t)min(m0 + r, M - 1) * ldx + k);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      // no branch on n < N (it would put every W load in its own basic block and serialise the L2 round
      // trips): padding columns re-read the last row, their sums are never stored
      const float4 wv = ldg4(W + (size_t)min(n, N - 1) * ldw + k);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r * NT + n] = dot4(xv[r], wv, acc[r * NT + n]);
    }
  }
#pragma unroll
  for (int g = 0; g < V / 32; ++g) {
    butterfly32<V>(acc, g * 32, lane);
    red[w][g * 32 + lane] = acc[g * 32];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V; i += 256) {
    const int r = i / NT, n = i - r * NT;
    if (m0 + r >= M || n >= N) continue;
    float v = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v += red[j][i];
    if (bias) v += __ldg(bias + n);
    Y[(size_t)(m0 + r) * ldy + n] = act_apply(act, v);
  }
}

// ---------------------------------------------------------------- classifier output layer, one launch
// logits = X.W^T + b (N <= 16), logp = log_softmax(logits), loss_rows = MCCE(logp, target),
// grad = d loss / d logits, and the data gradient of the layer below,
// dX = (grad . W) (.) act'(X)  (X is that layer's activation output) -- the three launches that sit
// between the last hidden contraction and the first big data-gradient contraction on the critical
// path of a step.  Same K-split structure as skinny_fwd_ksplit_kernel; a CTA owns R complete rows, so
// the row-wise loss (same arithmetic and reduction order as row_kernel<OP_LSM_MCCE, 8, 4>,
// activation_function_kernels.cu:289-325 + loss_kernels.cu:171-185 +
// multiclass_cross_entropy_loss_function.cc:61-71) and the data gradient of its rows need nothing from
// other CTAs.
template <int NT, int R>
__global__ void __launch_bounds__(256, (R <= 4) ? 2 : 1) output_layer_fused_kernel(int M, int N, int K, const float *__restrict__ X, int ldx,
                                                                 const float *__restrict__ W, int ldw,
                                                                 const float *__restrict__ bias,
                                                                 const float *__restrict__ target, float *__restrict__ logits,
                                                                 float *__restrict__ logp_out, float *__restrict__ loss_rows,
                                                                 float *__restrict__ grad, int dact, float *__restrict__ dX,
                                                                 int lddx) {
  constexpr int V = R * NT;
  static_assert(V % 32 == 0 && R <= 8, "R*NT must be a multiple of 32");
  __shared__ float red[8][V];
  __shared__ float sg[R][NT];          // logits, then the gradient rows
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int m0 = blockIdx.x * R;
  pdl_wait();                  // X is the previous kernel's output
  pdl_launch_dependents();     // the data-gradient contraction that follows may set itself up on the idle SMs
  {
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.0f;
#pragma unroll 2
    for (int k = w * 128 + lane * 4; k < K; k += 8 * 128) {
      float4 xv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) xv[r] = ldg4(X + (size_t)min(m0 + r, M - 1) * ldx + k);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const float4 wv = ldg4(W + (size_t)min(n, N - 1) * ldw + k);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r * NT + n] = dot4(xv[r], wv, acc[r * NT + n]);
      }
    }
#pragma unroll
    for (int g = 0; g < V / 32; ++g) {
      butterfly32<V>(acc, g * 32, lane);
      red[w][g * 32 + lane] = acc[g * 32];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V; i += 256) {
    const int r = i / NT, n = i - r * NT;
    float v = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v += red[j][i];
    if (bias && n < N) v += __ldg(bias + n);
    sg[r][n] = v;
    if (logits && m0 + r < M && n < N) logits[(size_t)(m0 + r) * N + n] = v;
  }
  __syncthreads();
  // log_softmax + MCCE + gradient: 8 threads per row, elements t and t+8 (N <= 16)
  if (threadIdx.x < 8 * R) {
    const int r = threadIdx.x >> 3, t = threadIdx.x & 7;
    const bool active = m0 + r < M;
    const size_t off = (size_t)min(m0 + r, M - 1) * N;
    float va[2], vb[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = t + 8 * j;
      va[j] = (c < N) ? sg[r][c] : 0.0f;
      vb[j] = (c < N) ? __ldg(target + off + c) : 0.0f;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (t + 8 * j < N) mx = fmaxf(mx, va[j]);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (t + 8 * j < N) { va[j] -= mx; sum += expf(va[j]); }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float lse = logf(sum);
    float loss = 0.0f;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = t + 8 * j;
      if (c < N) {
        const float logp = va[j] - lse;
        if (active && logp_out) logp_out[off + c] = logp;
        const float tc = fminf(fmaxf(vb[j], NEAR_ZERO_F), 1.0f - NEAR_ZERO_F);
        if (tc > NEAR_ZERO_F) loss += -tc * logp;
        const float cl = fminf(fmaxf(logp, logf(NEAR_ZERO_F)), logf(1.0f - NEAR_ZERO_F));
        const float g = expf(cl) - vb[j];
        if (active && grad) grad[off + c] = g;
        sg[r][c] = g;
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (active && t == 0 && loss_rows) loss_rows[m0 + r] = loss;
  }
  if (dX == nullptr) return;
  __syncthreads();
  // data gradient of the rows of this CTA: thread = the same 4 input features as in the forward loop.
  // No branches around the loads (they would serialise the L2 round trips): rows past M are clamped,
  // only the stores are predicated.
#pragma unroll 1
  for (int k = w * 128 + lane * 4; k < K; k += 8 * 128) {
    // every load of the iteration is issued before the first use (a load consumed right after its issue
    // serialises the L2 round trips: 25 us instead of 9 for the whole kernel)
    float4 wv[NT], y[R];
#pragma unroll
    for (int n = 0; n < NT; ++n) wv[n] = ldg4(W + (size_t)min(n, N - 1) * ldw + k);
    if (dact != B200_ACT_NONE) {
#pragma unroll
      for (int r = 0; r < R; ++r) y[r] = ldg4(X + (size_t)min(m0 + r, M - 1) * ldx + k);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const float g = (n < N) ? sg[r][n] : 0.0f;
        a.x = fmaf(g, wv[n].x, a.x);
        a.y = fmaf(g, wv[n].y, a.y);
        a.z = fmaf(g, wv[n].z, a.z);
        a.w = fmaf(g, wv[n].w, a.w);
      }
      if (dact != B200_ACT_NONE) {
        a.x *= act_deriv_from_output(dact, y[r].x);
        a.y *= act_deriv_from_output(dact, y[r].y);
        a.z *= act_deriv_from_output(dact, y[r].z);
        a.w *= act_deriv_from_output(dact, y[r].w);
      }
      if (m0 + r < M) *reinterpret_cast<float4 *>(dX + (size_t)(m0 + r) * lddx + k) = a;
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
  // all derivative-source loads of the chunk are issued before the first use (ROWS independent loads)
  float yv[ROWS][VW];
  if (dact != B200_ACT_NONE) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const size_t m = (size_t)min(m0 + r, M - 1);
      if (VEC) {
        const float4 y = ldg4(Yprev + m * ldyp + k);
        yv[r][0] = y.x; yv[r][VW > 1 ? 1 : 0] = y.y; yv[r][VW > 2 ? 2 : 0] = y.z; yv[r][VW > 3 ? 3 : 0] = y.w;
      } else {
        yv[r][0] = __ldg(Yprev + m * ldyp + k);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (m0 + r >= M) break;
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
#pragma unroll
      for (int e = 0; e < VW; ++e) a[e] *= act_deriv_from_output(dact, yv[r][e]);
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
  if (vec && K >= 512) {
    constexpr int R = 8;
    const int grid = (M + R - 1) / R;
    NT_DISPATCH(N, { skinny_fwd_ksplit_kernel<NT, R><<<grid, 256, 0, ctx->stream>>>(M, N, K, X, ldx, W, ldw, bias, act, Y, ldy); });
    LAUNCH_CHECK(ctx);
    return B200_OK;
  }
  if (vec) {
    // W resident in shared memory: [K/4][NT] float4 + the 16x32xNT partials
    const int nt = pad4(N);
    const size_t smem = (size_t)(K / 4) * nt * 16 + (size_t)16 * 32 * nt * 4;
    if (smem <= 200 * 1024) {
      const int grid = (M + 31) / 32;
      NT_DISPATCH(N, {
        auto kern = skinny_fwd_rows_kernel<NT>;
        if (ONCE_PER_DEVICE(ctx)) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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

extern "C" int b200_output_layer_fused(b200_ctx *ctx, int M, int N, int K, const float *X, int ldx, const float *W, int ldw,
                                       const float *bias, const float *target, float *logits, float *logp,
                                       float *loss_rows, float *grad, int dact, float *dX, int lddx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && X && W && target, "NULL pointer");
  ARG_CHECK(M >= 1 && N >= 3 && K >= 1, "bad sizes");
  const bool ok = N <= SK_MAXN && (K % 4 == 0) && (ldx % 4 == 0) && (ldw % 4 == 0) && al16(X) && al16(W) &&
                  (dX == nullptr || ((lddx % 4 == 0) && al16(dX)));
  if (!ok) {
    b200_set_error("b200_output_layer_fused: needs N <= %d, K %% 4 == 0 and 16-byte aligned rows", SK_MAXN);
    return B200_ERR_UNSUPPORTED;
  }
  // 8 rows per CTA, one CTA per SM (4 rows per CTA with two CTAs per SM measured 14 us against 9 us: the
  // class padding to 16 and the second read of W cost more than the extra warps hide)
  constexpr int R = 8;
  const int grid = (M + R - 1) / R;
  NT_DISPATCH(N, {
    CUDA_TRY(launch_pdl(output_layer_fused_kernel<NT, R>, dim3(grid), dim3(256), 0, ctx->stream, M, N, K, X, ldx, W, ldw, bias,
                        target, logits, logp, loss_rows, grad, dact, dX, lddx));
  });
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

int skinny_bwd_data(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy, const float *W, int ldw, int dact,
                    const float *Yprev, int ldyp, float *dX, int lddx) {
  const bool vec = (K % 4 == 0) && (ldw % 4 == 0) && (lddx % 4 == 0) && al16(W) && al16(dX) &&
                   (dact == B200_ACT_NONE || ((ldyp % 4 == 0) && al16(Yprev)));
  constexpr int ROWS = 8;
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
    if (vec) {
      PREFER_MAX_SMEM_ONCE((skinny_wgrad_partial_kernel<NT, ROWS, true>));
      skinny_wgrad_partial_kernel<NT, ROWS, true><<<grid, 256, 0, ctx->stream>>>(M, N, K, dY, lddy, X, ldx, part, db ? part_b : nullptr);
    } else {
      skinny_wgrad_partial_kernel<NT, ROWS, false><<<grid, 256, 0, ctx->stream>>>(M, N, K, dY, lddy, X, ldx, part, db ? part_b : nullptr);
    }
  });
  LAUNCH_CHECK(ctx);
  ReduceJob j0{part, dW, (int)nk, lddw, K};
  ReduceJob j1{part_b, db, db ? N : 0, 1, 1};
  dim3 rgrid((unsigned)((nk + 255) / 256), db ? 2 : 1);
  PREFER_MAX_SMEM_ONCE(reduce_partials_kernel);
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
    PREFER_MAX_SMEM_ONCE((colsum_partial_kernel<ROWS, true>));
    colsum_partial_kernel<ROWS, true><<<grid, 256, 0, ctx->stream>>>(M, N, dy, ld, part);
  } else {
    dim3 grid((N + 31) / 32, chunks);
    PREFER_MAX_SMEM_ONCE((colsum_partial_kernel<ROWS, false>));
    colsum_partial_kernel<ROWS, false><<<grid, 256, 0, ctx->stream>>>(M, N, dy, ld, part);
  }
  LAUNCH_CHECK(ctx);
  ReduceJob j0{part, out, N, 1, 1};
  PREFER_MAX_SMEM_ONCE(reduce_partials_kernel);
  reduce_partials_kernel<<<dim3((N + 255) / 256, 1), 256, 0, ctx->stream>>>(chunks, scale, beta, j0, j0);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
