// One-pass row kernels: softmax / log_softmax forward, softmax backward, the MSE /
// cross-entropy / multi-class cross-entropy losses with their gradients, and the fused
// log_softmax + MCCE + gradient pass.  HBM-bound: every input element is read from
// global memory once (rows are cached in registers), every output written once.
//
// Replace applySoftmax / applyLogSoftmax / applySoftmaxDerivative
// (ann/ann/c_src/activation_function_kernels.cu:184-355; the reference's GPU branch issues
// O(classes) axpy/cmul launches per call) and the loss maps + per-row reductions of
// ann/loss/c_src/loss_kernels.cu:38-266.
//
// A row is owned by TPR threads (8 or 32 lanes of a warp, or a whole 256-thread CTA),
// each holding VPT strided elements, so C <= TPR*VPT.  Reductions are warp shuffles
// (+ one shared-memory hop for the CTA-per-row shape).
#include <math.h>

#include "common.cuh"

namespace {

struct SumOp { __device__ static float id() { return 0.0f; } __device__ static float f(float a, float b) { return a + b; } };
struct MaxOp { __device__ static float id() { return -INFINITY; } __device__ static float f(float a, float b) { return fmaxf(a, b); } };
struct MinOp { __device__ static float id() { return INFINITY; } __device__ static float f(float a, float b) { return fminf(a, b); } };

template <int TPR, class R>
__device__ __forceinline__ float group_allreduce(float v, float *sm) {
  if (TPR <= 32) {
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) v = R::f(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  } else {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = R::f(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = R::id();
#pragma unroll
    for (int i = 0; i < TPR / 32; ++i) r = R::f(r, sm[i]);
    __syncthreads();
    return r;
  }
}

enum RowOpKind { OP_SOFTMAX, OP_LOG_SOFTMAX, OP_SOFTMAX_BWD, OP_MCCE, OP_MSE, OP_CE, OP_LSM_MCCE };

struct RowArgs {
  const float *a;   // primary input  (x / y / logp / out / logits)
  const float *b;   // second input   (dy / target)
  float *o1;        // primary output (y / dx / grad / logp)
  float *o2;        // second output  (grad for the fused op)
  float *rows;      // per-row scalar output (loss rows)
};

template <int KIND, int TPR, int VPT>
__global__ void __launch_bounds__(256) row_kernel(int M, int C, RowArgs p) {
  __shared__ float sm[8];
  constexpr int RPB = 256 / TPR;
  int row = blockIdx.x * RPB + threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  const bool active = row < M;
  if (!active) row = M - 1;  // keep the lanes in the shuffles; writes are masked
  const size_t off = (size_t)row * C;

  float va[VPT], vb[VPT];
  constexpr bool NEEDS_B = (KIND != OP_SOFTMAX && KIND != OP_LOG_SOFTMAX);
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int c = t + j * TPR;
    va[j] = (c < C) ? __ldg(p.a + off + c) : 0.0f;
    vb[j] = (NEEDS_B && c < C) ? __ldg(p.b + off + c) : 0.0f;
  }

  if (KIND == OP_SOFTMAX) {
    // activation_function_kernels.cu:209-248: subtract the row minimum, clamped to max-30
    float mx = -INFINITY, mn = INFINITY;
#pragma unroll
    for (int j = 0; j < VPT; ++j)
      if (t + j * TPR < C) { mx = fmaxf(mx, va[j]); mn = fminf(mn, va[j]); }
    mx = group_allreduce<TPR, MaxOp>(mx, sm);
    mn = group_allreduce<TPR, MinOp>(mn, sm);
    if (mx - mn > 30.0f) mn = mx - 30.0f;
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < VPT; ++j)
      if (t + j * TPR < C) { va[j] = expf(va[j] - mn); s += va[j]; }
    s = group_allreduce<TPR, SumOp>(s, sm);
    const float ratio = 1.0f / s;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = t + j * TPR;
      if (active && c < C) p.o1[off + c] = va[j] * ratio;
    }
  } else if (KIND == OP_LOG_SOFTMAX || KIND == OP_LSM_MCCE) {
    // activation_function_kernels.cu:289-325: x - max, then subtract log(sum exp)
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < VPT; ++j)
      if (t + j * TPR < C) mx = fmaxf(mx, va[j]);
    mx = group_allreduce<TPR, MaxOp>(mx, sm);
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < VPT; ++j)
      if (t + j * TPR < C) { va[j] -= mx; s += expf(va[j]); }
    s = group_allreduce<TPR, SumOp>(s, sm);
    const float lse = logf(s);
    float loss = 0.0f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = t + j * TPR;
      if (c < C) {
        const float logp = va[j] - lse;
        if (KIND == OP_LOG_SOFTMAX) {
          if (active) p.o1[off + c] = logp;
        } else {
          // fused MCCE (loss_kernels.cu:171-185) + gradient
          // (multiclass_cross_entropy_loss_function.cc:61-71)
          if (active && p.o1) p.o1[off + c] = logp;
          const float tc = fminf(fmaxf(vb[j], NEAR_ZERO_F), 1.0f - NEAR_ZERO_F);
          if (tc > NEAR_ZERO_F) loss += -tc * logp;
          const float cl = fminf(fmaxf(logp, logf(NEAR_ZERO_F)), logf(1.0f - NEAR_ZERO_F));
          if (active && p.o2) p.o2[off + c] = expf(cl) - vb[j];
        }
      }
    }
    if (KIND == OP_LSM_MCCE) {
      loss = group_allreduce<TPR, SumOp>(loss, sm);
      if (active && t == 0 && p.rows) p.rows[row] = loss;
    }
  } else if (KIND == OP_SOFTMAX_BWD) {
    // activation_function_kernels.cu:331-355: dx = y * (dy - sum_j y_j dy_j)
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < VPT; ++j)
      if (t + j * TPR < C) s += va[j] * vb[j];
    s = group_allreduce<TPR, SumOp>(s, sm);
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = t + j * TPR;
      if (active && c < C) p.o1[off + c] = (vb[j] - s) * va[j];
    }
  } else {
    float loss = 0.0f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = t + j * TPR;
      if (c >= C) continue;
      const float o = va[j], tg = vb[j];
      float g;
      if (KIND == OP_MSE) {
        // loss_kernels.cu:38-44 ; mse_loss_function.cc:87
        const float d = o - tg;
        loss += 0.5f * d * d;
        g = d;
      } else if (KIND == OP_MCCE) {
        const float tc = fminf(fmaxf(tg, NEAR_ZERO_F), 1.0f - NEAR_ZERO_F);
        if (tc > NEAR_ZERO_F) loss += -tc * o;
        const float cl = fminf(fmaxf(o, logf(NEAR_ZERO_F)), logf(1.0f - NEAR_ZERO_F));
        g = expf(cl) - tg;
      } else {  // OP_CE  loss_kernels.cu:48-83,121-133
        const float log_o = fminf(fmaxf(o, logf(NEAR_ZERO_F)), logf(1.0f - NEAR_ZERO_F));
        const float ev = expf(log_o);
        const float log_inv_o = (float)log(1.0 - (double)ev);
        const float tc = fminf(fmaxf(tg, NEAR_ZERO_F), 1.0f - NEAR_ZERO_F);
        const float inv_t = fminf(fmaxf(1.0f - tg, NEAR_ZERO_F), 1.0f - NEAR_ZERO_F);
        float s = (tc > NEAR_ZERO_F) ? -tc * log_o : 0.0f;
        if (inv_t > NEAR_ZERO_F) s -= inv_t * log_inv_o;
        loss += s;
        g = ev - tg;
      }
      if (active && p.o1) p.o1[off + c] = g;
    }
    loss = group_allreduce<TPR, SumOp>(loss, sm);
    if (active && t == 0 && p.rows) p.rows[row] = loss;
  }
}

// Fused log_softmax + MCCE + gradient for wide rows (the 10 000-class output of the NNLM configuration):
// one 256-thread CTA per row, 16-byte accesses, only the logits are cached in registers (the target is
// read once, in the second phase), so four CTAs are resident per SM.  Same arithmetic as
// row_kernel<OP_LSM_MCCE> (activation_function_kernels.cu:289-325, loss_kernels.cu:171-185,
// multiclass_cross_entropy_loss_function.cc:61-71); the generic kernel held logits and target in 80
// registers per thread with 4-byte accesses and reached 2.7 TB/s at 4096 x 10000.
template <int V4>   // float4 per thread: C <= 256 * 4 * V4
__global__ void __launch_bounds__(256, 4) lsm_mcce_wide_kernel(int M, int C, RowArgs p) {
  __shared__ float sm[8];
  const int row = blockIdx.x, t = threadIdx.x;
  const size_t off = (size_t)row * C;
  const int C4 = C >> 2;
  const float4 *a4 = reinterpret_cast<const float4 *>(p.a + off);
  const float4 *b4 = reinterpret_cast<const float4 *>(p.b + off);
  float4 va[V4];
#pragma unroll
  for (int j = 0; j < V4; ++j) {
    const int c = t + j * 256;
    va[j] = (c < C4) ? __ldcs(a4 + c) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  }
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < V4; ++j) mx = fmaxf(mx, fmaxf(fmaxf(va[j].x, va[j].y), fmaxf(va[j].z, va[j].w)));
  mx = group_allreduce<256, MaxOp>(mx, sm);
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < V4; ++j) {
    if (t + j * 256 < C4) {
      va[j].x -= mx; va[j].y -= mx; va[j].z -= mx; va[j].w -= mx;
      s += expf(va[j].x) + expf(va[j].y) + expf(va[j].z) + expf(va[j].w);
    }
  }
  s = group_allreduce<256, SumOp>(s, sm);
  const float lse = logf(s);
  const float lo = logf(NEAR_ZERO_F), hi = logf(1.0f - NEAR_ZERO_F);
  float loss = 0.0f;
  float4 *o1 = p.o1 ? reinterpret_cast<float4 *>(p.o1 + off) : nullptr;
  float4 *o2 = p.o2 ? reinterpret_cast<float4 *>(p.o2 + off) : nullptr;
  auto one = [&](float v, float tg, float &logp, float &g) {
    logp = v - lse;
    const float tc = fminf(fmaxf(tg, NEAR_ZERO_F), 1.0f - NEAR_ZERO_F);
    if (tc > NEAR_ZERO_F) loss += -tc * logp;
    g = expf(fminf(fmaxf(logp, lo), hi)) - tg;
  };
#pragma unroll
  for (int j = 0; j < V4; ++j) {
    const int c = t + j * 256;
    if (c < C4) {
      const float4 tg = __ldcs(b4 + c);
      float4 lp, g;
      one(va[j].x, tg.x, lp.x, g.x); one(va[j].y, tg.y, lp.y, g.y); one(va[j].z, tg.z, lp.z, g.z); one(va[j].w, tg.w, lp.w, g.w);
      if (o1) __stcs(o1 + c, lp);
      if (o2) __stcs(o2 + c, g);
    }
  }
  loss = group_allreduce<256, SumOp>(loss, sm);
  if (t == 0 && p.rows) p.rows[row] = loss;
}

template <int KIND>
int launch_rows(b200_ctx *ctx, int M, int C, const RowArgs &p) {
  if (M <= 0 || C <= 0) return B200_OK;
#define ROW_LAUNCH(TPR, VPT)                                                             \
  do {                                                                                   \
    constexpr int RPB = 256 / (TPR);                                                     \
    row_kernel<KIND, TPR, VPT><<<(M + RPB - 1) / RPB, 256, 0, ctx->stream>>>(M, C, p);   \
    LAUNCH_CHECK(ctx);                                                                   \
    return B200_OK;                                                                      \
  } while (0)
  if (KIND == OP_LSM_MCCE && C >= 2048 && (C & 3) == 0 && C <= 256 * 4 * 16 &&
      (((uintptr_t)p.a | (uintptr_t)p.b | (uintptr_t)p.o1 | (uintptr_t)p.o2) & 15) == 0) {
    if (C <= 256 * 4 * 4) lsm_mcce_wide_kernel<4><<<M, 256, 0, ctx->stream>>>(M, C, p);
    else if (C <= 256 * 4 * 10) lsm_mcce_wide_kernel<10><<<M, 256, 0, ctx->stream>>>(M, C, p);
    else lsm_mcce_wide_kernel<16><<<M, 256, 0, ctx->stream>>>(M, C, p);
    LAUNCH_CHECK(ctx);
    return B200_OK;
  }
  if (C <= 32) ROW_LAUNCH(8, 4);
  if (C <= 256) ROW_LAUNCH(32, 8);
  if (C <= 1024) ROW_LAUNCH(32, 32);
  if (C <= 2048) ROW_LAUNCH(256, 8);
  if (C <= 4096) ROW_LAUNCH(256, 16);
  if (C <= 10240) ROW_LAUNCH(256, 40);
  if (C <= 16384) ROW_LAUNCH(256, 64);
#undef ROW_LAUNCH
  b200_set_error("row kernels support up to 16384 classes per row, got %d", C);
  return B200_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" int b200_softmax_fwd(b200_ctx *ctx, int M, int C, const float *x, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && y, "NULL pointer");
  RowArgs p{x, nullptr, y, nullptr, nullptr};
  return launch_rows<OP_SOFTMAX>(ctx, M, C, p);
}
extern "C" int b200_log_softmax_fwd(b200_ctx *ctx, int M, int C, const float *x, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && y, "NULL pointer");
  RowArgs p{x, nullptr, y, nullptr, nullptr};
  return launch_rows<OP_LOG_SOFTMAX>(ctx, M, C, p);
}
extern "C" int b200_softmax_bwd(b200_ctx *ctx, int M, int C, const float *y, const float *dy, float *dx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && y && dy && dx, "NULL pointer");
  RowArgs p{y, dy, dx, nullptr, nullptr};
  return launch_rows<OP_SOFTMAX_BWD>(ctx, M, C, p);
}
extern "C" int b200_mcce_loss_grad(b200_ctx *ctx, int M, int C, const float *logp, const float *target,
                                   float *loss_rows, float *grad) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && logp && target, "NULL pointer");
  RowArgs p{logp, target, grad, nullptr, loss_rows};
  return launch_rows<OP_MCCE>(ctx, M, C, p);
}
extern "C" int b200_mse_loss_grad(b200_ctx *ctx, int M, int C, const float *out, const float *target,
                                  float *loss_rows, float *grad) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && out && target, "NULL pointer");
  RowArgs p{out, target, grad, nullptr, loss_rows};
  return launch_rows<OP_MSE>(ctx, M, C, p);
}
extern "C" int b200_ce_loss_grad(b200_ctx *ctx, int M, int C, const float *log_out, const float *target,
                                 float *loss_rows, float *grad) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && log_out && target, "NULL pointer");
  RowArgs p{log_out, target, grad, nullptr, loss_rows};
  return launch_rows<OP_CE>(ctx, M, C, p);
}
extern "C" int b200_log_softmax_mcce_fused(b200_ctx *ctx, int M, int C, const float *logits,
                                           const float *target, float *logp, float *loss_rows,
                                           float *grad) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && logits && target, "NULL pointer");
  RowArgs p{logits, target, logp, grad, loss_rows};
  return launch_rows<OP_LSM_MCCE>(ctx, M, C, p);
}
