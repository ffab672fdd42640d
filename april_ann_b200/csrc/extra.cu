// The cheap components and optimizers that share the element-wise / multi-tensor kernels:
//   activation functions log_logistic / softplus / softsign / leaky_relu / hardtanh / prelu
//       ann/ann/c_src/activation_function_kernels.cu:52-182, cmath_overloads.h:727-744,993-1023,1073-1109,1143-1153
//   dropout                       ann/ann/c_src/dropout_component.cc:67-134, dropout_kernel.cu:29
//   zero_one loss                 ann/loss/c_src/zero_one_loss_function.cc:39-132
//   global gradient-norm clip     trainable/lua_src/supervised.lua:805-811
//   adagrad / rmsprop / adadelta  ann/optimizer/lua_src/optimizer_{adagrad,rmsprop,adadelta}.lua
// All HBM-bound: float4 grid-stride maps, warp-shuffle reductions, one launch for all tensors.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace {

constexpr int TPB = 256;

inline int grid_for(size_t n, int sm_count) {
  size_t blocks = (n + TPB - 1) / TPB;
  const size_t cap = (size_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}
inline bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

// ---------------------------------------------------------------- parametrised activations
__device__ __forceinline__ float actx_apply(int act, float p0, float p1, float x) {
  switch (act) {
    case B200_ACT_LOG_LOGISTIC: return x < -10.0f ? x : -log1pf(expf(-x));
    case B200_ACT_SOFTPLUS: return x > 10.0f ? x : log1pf(expf(x));
    case B200_ACT_SOFTSIGN: return x / (1.0f + fabsf(x));
    case B200_ACT_LEAKY_RELU: return x > 0.0f ? x : p0 * x;
    case B200_ACT_HARDTANH: return fminf(fmaxf(x, p0), p1);
    default: return act_apply(act, x);
  }
}
// derivative from the INPUT x and/or the OUTPUT y, whichever the reference's functor takes
__device__ __forceinline__ float actx_deriv(int act, float p0, float p1, float x, float y) {
  switch (act) {
    case B200_ACT_LOG_LOGISTIC: return 1.0f;                    // cancelled by the cross-entropy derivative
    case B200_ACT_SOFTPLUS: return 1.0f / (expf(-x) + 1.0f);    // m_softplus_der = logistic(input)
    case B200_ACT_SOFTSIGN: {
      const float v = fminf(fmaxf(y, -1.0f + NEAR_ZERO_F), 1.0f - NEAR_ZERO_F);
      const float a = 1.0f + fabsf(v);
      return 1.0f / (a * a);
    }
    case B200_ACT_LEAKY_RELU: return x > 0.0f ? 1.0f : p0;
    case B200_ACT_HARDTANH: return (x < p0 || x > p1) ? 0.0f : 1.0f;
    default: return act_deriv_from_output(act, y);
  }
}

__global__ void __launch_bounds__(TPB) actx_fwd_kernel(int act, float p0, float p1, size_t n, const float *__restrict__ x,
                                                       float *__restrict__ y, bool vec) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const size_t n4 = n >> 2;
    for (size_t i = tid; i < n4; i += nth) {
      float4 v = reinterpret_cast<const float4 *>(x)[i];
      v.x = actx_apply(act, p0, p1, v.x); v.y = actx_apply(act, p0, p1, v.y);
      v.z = actx_apply(act, p0, p1, v.z); v.w = actx_apply(act, p0, p1, v.w);
      reinterpret_cast<float4 *>(y)[i] = v;
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += nth) y[i] = actx_apply(act, p0, p1, x[i]);
  } else {
    for (size_t i = tid; i < n; i += nth) y[i] = actx_apply(act, p0, p1, x[i]);
  }
}
__global__ void __launch_bounds__(TPB) actx_bwd_kernel(int act, float p0, float p1, size_t n, const float *__restrict__ x,
                                                       const float *__restrict__ y, const float *__restrict__ dy,
                                                       float *__restrict__ dx) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (size_t i = tid; i < n; i += nth)
    dx[i] = actx_deriv(act, p0, p1, x ? x[i] : 0.0f, y ? y[i] : 0.0f) * dy[i];
}

// ---------------------------------------------------------------- PReLU
// y[m,n] = x>0 ? x : a[n]*x   (prelu_actf_component.cc:55-64; scalar: one a for every unit)
__global__ void __launch_bounds__(TPB) prelu_fwd_kernel(size_t total, int N, const float *__restrict__ x,
                                                        const float *__restrict__ a, int scalar, float *__restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = v > 0.0f ? v : __ldg(a + (scalar ? 0 : (i % N))) * v;
  }
}
__global__ void __launch_bounds__(TPB) prelu_bwd_kernel(size_t total, int N, const float *__restrict__ x,
                                                        const float *__restrict__ a, int scalar,
                                                        const float *__restrict__ dy, float *__restrict__ dx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    dx[i] = (x[i] > 0.0f ? 1.0f : __ldg(a + (scalar ? 0 : (i % N)))) * dy[i];
}
// e[m,n] = (x<0) * x * dy  (prelu_actf_component.cc:96-108), written to a temporary whose column sums
// (or total sum) give the gradient
__global__ void __launch_bounds__(TPB) prelu_err_kernel(size_t total, const float *__restrict__ x,
                                                        const float *__restrict__ dy, float *__restrict__ e) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    e[i] = x[i] < 0.0f ? x[i] * dy[i] : 0.0f;
}
__global__ void scalar_axpby_kernel(const float *__restrict__ s, float scale, float beta, float *__restrict__ out) {
  *out = (beta != 0.0f ? beta * *out : 0.0f) + scale * *s;
}

// ---------------------------------------------------------------- dropout
// The mask follows the reference's stream exactly: element i (row-major) is dropped iff the i-th
// draw rand() = randInt32 * (1/4294967295) of the component's MT19937 is < prob
// (dropout_component.cc:91-95).  MT19937's recurrence x[k+624] = x[k+397] ^ twist(x[k], x[k+1]) has a
// shortest dependency distance of 227 words, so one CTA advances the generator 227 words per barrier:
// state[624] lives in shared memory, every block of 624 outputs takes three barriers (227 + 227 + 170).
struct MtDev {
  uint32_t state[624];
  int32_t pos;      // next unread output of the current block (624 = block exhausted)
  int32_t pad_[3];
};
__device__ __forceinline__ uint32_t mt_twist(uint32_t m, uint32_t s0, uint32_t s1) {
  return m ^ (((s0 & 0x80000000u) | (s1 & 0x7fffffffu)) >> 1) ^ ((s1 & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  return y ^ (y >> 18);
}
__global__ void __launch_bounds__(256) dropout_mask_kernel(MtDev *mt, size_t n, double prob, float *__restrict__ mask) {
  __shared__ uint32_t st[624];
  const int t = threadIdx.x;
  for (int i = t; i < 624; i += 256) st[i] = mt->state[i];
  int pos = mt->pos;
  __syncthreads();
  size_t done = 0;
  while (done < n) {
    if (pos >= 624) {
      // reload (MersenneTwister.cc:243-255), in three independent waves
      // (every wave reads its operands, then a barrier, then writes: word i+1 is still the old one when
      // word i is computed)
      uint32_t v = 0;
      if (t < 227) v = mt_twist(st[t + 397], st[t], st[t + 1]);
      __syncthreads();
      if (t < 227) st[t] = v;
      __syncthreads();
      if (t < 227) v = mt_twist(st[t], st[227 + t], st[227 + t + 1]);
      __syncthreads();
      if (t < 227) st[227 + t] = v;
      __syncthreads();
      if (t < 169) v = mt_twist(st[227 + t], st[454 + t], st[454 + t + 1]);
      __syncthreads();
      if (t < 169) st[454 + t] = v;
      __syncthreads();
      if (t == 0) st[623] = mt_twist(st[396], st[623], st[0]);
      __syncthreads();
      pos = 0;
    }
    const size_t take = min((size_t)(624 - pos), n - done);
    for (size_t i = t; i < take; i += 256) {
      const double r = (double)mt_temper(st[pos + i]) * (1.0 / 4294967295.0);
      mask[done + i] = r < prob ? 0.0f : 1.0f;
    }
    pos += (int)take;
    done += take;
  }
  __syncthreads();
  for (int i = t; i < 624; i += 256) mt->state[i] = st[i];
  if (t == 0) mt->pos = pos;
}
// y = mask < 0.5 ? value : x   (m_curried_mask, cmath_overloads.h:1403-1411)
__global__ void __launch_bounds__(TPB) mask_apply_kernel(size_t n, const float *__restrict__ x, const float *__restrict__ mask,
                                                         float value, float *__restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = mask[i] < 0.5f ? value : x[i];
}

// ---------------------------------------------------------------- zero-one loss
// one warp per pattern; first maximum wins (matMax uses '>')
__global__ void __launch_bounds__(TPB) zero_one_kernel(int M, int C, const float *__restrict__ out,
                                                       const float *__restrict__ target, int tcols, float TH,
                                                       float *__restrict__ rows) {
  const int row = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  if (C == 1) {
    if (lane == 0) {
      const bool pred = out[row] > TH, want = target[row] > 0.5f;
      rows[row] = pred != want ? 1.0f : 0.0f;
    }
    return;
  }
  auto argmax = [&](const float *p, int n) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < n; c += 32) {
      const float v = p[c];
      if (v > best || (v == best && c < bi)) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    return bi;
  };
  const int am = argmax(out + (size_t)row * C, C);
  int want;
  if (tcols == C) want = argmax(target + (size_t)row * C, C);
  else want = (int)(target[row] - 1.0f);   // class labels start at 1
  if (lane == 0) rows[row] = am != want ? 1.0f : 0.0f;
}

// ---------------------------------------------------------------- gradient-norm clip
// g *= max_norm / sqrt(norm2sq) when sqrt(norm2sq) > max_norm   (supervised.lua:805-811)
__global__ void __launch_bounds__(TPB) clip_scale_kernel(size_t n, float *__restrict__ g, const float *__restrict__ norm2sq,
                                                         float max_norm) {
  const float nrm = sqrtf(*norm2sq);
  if (!(nrm > max_norm)) return;
  const float ratio = max_norm / nrm;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if ((((uintptr_t)g) & 15) == 0) {
    float4 *g4 = reinterpret_cast<float4 *>(g);
    const size_t n4 = n >> 2;
    for (size_t i = tid; i < n4; i += nth) {
      float4 v = g4[i];
      v.x *= ratio; v.y *= ratio; v.z *= ratio; v.w *= ratio;
      g4[i] = v;
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += nth) g[i] *= ratio;
  } else {
    for (size_t i = tid; i < n; i += nth) g[i] *= ratio;
  }
}

// ---------------------------------------------------------------- adagrad / rmsprop / adadelta
template <int ALGO>
__global__ void __launch_bounds__(128) optimizer_kernel(const b200_opt_tensor *__restrict__ tensors, const int64_t *count_dev) {
  const b200_opt_tensor t = tensors[blockIdx.y];
  const int64_t count = *reinterpret_cast<const volatile int64_t *>(count_dev);
  const bool prune = (count % 100) == 0;
  const float lr = t.lr, mt = t.momentum, decay = t.decay, eps = t.epsilon, l2 = t.weight_decay;
  const float omd = 1.0f - decay;
  const size_t nth = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += nth) {
    float w = t.w[i], g = t.g[i];
    if (l2 > 0.0f) g = fmaf(l2, w, g);
    if (ALGO == B200_OPT_ADAGRAD) {
      // optimizer_adagrad.lua:52-60
      float E = t.s1[i];
      E = (count == 0) ? g * g : decay * E + omd * (g * g);
      const float upd = g * (1.0f / (eps + sqrtf(E)));
      w = fmaf(-lr, upd, w);
      t.s1[i] = E;
    } else if (ALGO == B200_OPT_RMSPROP) {
      // optimizer_rmsprop.lua:69-79 (s1 = Erms, u = Eupdate)
      float Er = t.s1[i];
      Er = decay * Er;
      Er = fmaf(omd, g * g, Er);
      const float tmp = (lr / sqrtf(Er + eps)) * g;
      float Eu;
      if (mt > 0.0f) Eu = fmaf(mt, t.u[i], tmp);
      else Eu = tmp;
      w -= Eu;
      t.s1[i] = Er;
      t.u[i] = Eu;
    } else {
      // optimizer_adadelta.lua:53-84 (s1 = Egradient, s2 = Eupdate, u = lr * last update)
      float u = t.u[i];
      if (mt > 0.0f) w = fmaf(mt, u, w);
      float Eg = t.s1[i], Eu = t.s2[i];
      Eg = decay * Eg + omd * (g * g);
      u = -(g * (sqrtf(Eu + eps) / sqrtf(Eg + eps)));
      Eu = decay * Eu + omd * (u * u);
      w = fmaf(lr, u, w);
      t.s1[i] = Eg;
      t.s2[i] = Eu;
      t.u[i] = u * lr;
    }
    if (prune && fabsf(w) < FLT_MIN) w = 0.0f;
    t.w[i] = w;
    if (t.write_back_grad) t.g[i] = g;
  }
}
// rmsprop's Nesterov look-ahead: w -= mt * Eupdate before the gradient is evaluated (optimizer_rmsprop.lua:44-51)
__global__ void __launch_bounds__(128) lookahead_kernel(const b200_opt_tensor *__restrict__ tensors) {
  const b200_opt_tensor t = tensors[blockIdx.y];
  if (!(t.momentum > 0.0f)) return;
  const size_t nth = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += nth) t.w[i] = fmaf(-t.momentum, t.u[i], t.w[i]);
}
__global__ void __launch_bounds__(128) max_norm_rows_kernel(float *__restrict__ w, int rows, int cols, float mnp) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float *p = w + (size_t)row * cols;
  float s = 0.0f;
  for (int c = lane; c < cols; c += 32) s = fmaf(p[c], p[c], s);
  s = warp_sum(s);
  const float n2 = sqrtf(s);
  if (n2 > mnp) {
    const float r = mnp / n2;
    for (int c = lane; c < cols; c += 32) p[c] *= r;
  }
}

}  // namespace

extern "C" int b200_actf_fwd_ex(b200_ctx *ctx, int act, float p0, float p1, size_t n, const float *x, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && y, "NULL pointer");
  ARG_CHECK(act != B200_ACT_SOFTMAX && act != B200_ACT_LOG_SOFTMAX, "row-wise activations use b200_softmax_fwd / b200_log_softmax_fwd");
  if (n == 0) return B200_OK;
  const bool vec = aligned16(x) && aligned16(y);
  actx_fwd_kernel<<<grid_for(vec ? (n >> 2) + 1 : n, ctx->sm_count), TPB, 0, ctx->stream>>>(
      act == B200_ACT_LINEAR ? B200_ACT_NONE : act, p0, p1, n, x, y, vec);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
extern "C" int b200_actf_bwd_ex(b200_ctx *ctx, int act, float p0, float p1, size_t n, const float *x, const float *y,
                                const float *dy, float *dx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dy && dx, "NULL pointer");
  const bool from_input = act == B200_ACT_SOFTPLUS || act == B200_ACT_LEAKY_RELU || act == B200_ACT_HARDTANH;
  ARG_CHECK(from_input ? x != nullptr : (y != nullptr || act == B200_ACT_LOG_LOGISTIC || act == B200_ACT_LINEAR || act == B200_ACT_NONE),
            "this activation's derivative needs the input (softplus, leaky_relu, hardtanh) / the output (others)");
  if (n == 0) return B200_OK;
  actx_bwd_kernel<<<grid_for(n, ctx->sm_count), TPB, 0, ctx->stream>>>(act == B200_ACT_LINEAR ? B200_ACT_NONE : act, p0, p1, n, x, y, dy, dx);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

extern "C" int b200_prelu_fwd(b200_ctx *ctx, int M, int N, const float *x, const float *a, int scalar, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && a && y, "NULL pointer");
  const size_t total = (size_t)M * N;
  if (!total) return B200_OK;
  prelu_fwd_kernel<<<grid_for(total, ctx->sm_count), TPB, 0, ctx->stream>>>(total, N, x, a, scalar, y);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
extern "C" int b200_prelu_bwd(b200_ctx *ctx, int M, int N, const float *x, const float *a, int scalar, const float *dy,
                              float *dx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && a && dy && dx, "NULL pointer");
  const size_t total = (size_t)M * N;
  if (!total) return B200_OK;
  prelu_bwd_kernel<<<grid_for(total, ctx->sm_count), TPB, 0, ctx->stream>>>(total, N, x, a, scalar, dy, dx);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
extern "C" int b200_prelu_grad(b200_ctx *ctx, int M, int N, const float *x, const float *dy, int scalar, float scale,
                               float beta, float *da, float *tmp) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && dy && da && tmp, "NULL pointer (tmp: M*N floats of workspace)");
  const size_t total = (size_t)M * N;
  if (!total) return B200_OK;
  prelu_err_kernel<<<grid_for(total, ctx->sm_count), TPB, 0, ctx->stream>>>(total, x, dy, tmp);
  LAUNCH_CHECK(ctx);
  if (!scalar) return colsum_scaled(ctx, M, N, tmp, N, scale, beta, da);
  // scalar PReLU: the sum of every element; reuse tmp[0] as the device scalar after the reduction
  int st = b200_sum(ctx, total, tmp, tmp);
  if (st) return st;
  scalar_axpby_kernel<<<1, 1, 0, ctx->stream>>>(tmp, scale, beta, da);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

extern "C" size_t b200_mt_state_bytes(void) { return sizeof(MtDev); }
extern "C" int b200_dropout_mask(b200_ctx *ctx, void *mt_state_dev, size_t n, float prob, float *mask) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && mt_state_dev && mask, "NULL pointer");
  if (n == 0) return B200_OK;
  dropout_mask_kernel<<<1, 256, 0, ctx->stream>>>((MtDev *)mt_state_dev, n, (double)prob, mask);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
extern "C" int b200_mask_apply(b200_ctx *ctx, size_t n, const float *x, const float *mask, float value, float *y) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && x && mask && y, "NULL pointer");
  if (n == 0) return B200_OK;
  mask_apply_kernel<<<grid_for(n, ctx->sm_count), TPB, 0, ctx->stream>>>(n, x, mask, value, y);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

extern "C" int b200_zero_one_loss(b200_ctx *ctx, int M, int C, const float *out, const float *target, int target_cols,
                                  float TH, float *loss_rows) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && out && target && loss_rows, "NULL pointer");
  ARG_CHECK(target_cols == C || target_cols == 1, "Incorrect target matrix bunch_size");
  if (M <= 0) return B200_OK;
  zero_one_kernel<<<(M + TPB / 32 - 1) / (TPB / 32), TPB, 0, ctx->stream>>>(M, C, out, target, target_cols, TH, loss_rows);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

extern "C" int b200_grad_clip(b200_ctx *ctx, size_t n, float *grads, float max_norm, float *norm2sq_dev) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && grads && norm2sq_dev, "NULL pointer");
  ARG_CHECK(max_norm > 0.0f, "max_gradients_norm must be positive");
  if (n == 0) return B200_OK;
  int st = b200_memset_zero(ctx, norm2sq_dev, sizeof(float));
  if (st) return st;
  st = b200_nrm2sq(ctx, n, grads, norm2sq_dev);
  if (st) return st;
  clip_scale_kernel<<<grid_for((n >> 2) + 1, ctx->sm_count), TPB, 0, ctx->stream>>>(n, grads, norm2sq_dev, max_norm);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

extern "C" int b200_optimizer_multi_tensor(b200_ctx *ctx, int algo, int ntensors, const b200_opt_tensor *tensors_dev,
                                           const b200_opt_tensor *tensors_host, int64_t *count_dev, int increment_count) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && tensors_dev && tensors_host && count_dev, "NULL pointer");
  ARG_CHECK(algo == B200_OPT_ADAGRAD || algo == B200_OPT_RMSPROP || algo == B200_OPT_ADADELTA, "unknown optimizer");
  if (ntensors <= 0) return B200_OK;
  size_t max_n = 0;
  for (int i = 0; i < ntensors; ++i) max_n = tensors_host[i].n > max_n ? (size_t)tensors_host[i].n : max_n;
  size_t bx = (max_n + 128 * 4 - 1) / (128 * 4);
  const size_t cap = (size_t)ctx->sm_count * 8;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)ntensors);
  if (algo == B200_OPT_ADAGRAD) optimizer_kernel<B200_OPT_ADAGRAD><<<grid, 128, 0, ctx->stream>>>(tensors_dev, count_dev);
  else if (algo == B200_OPT_RMSPROP) optimizer_kernel<B200_OPT_RMSPROP><<<grid, 128, 0, ctx->stream>>>(tensors_dev, count_dev);
  else optimizer_kernel<B200_OPT_ADADELTA><<<grid, 128, 0, ctx->stream>>>(tensors_dev, count_dev);
  LAUNCH_CHECK(ctx);
  for (int i = 0; i < ntensors; ++i) {
    const b200_opt_tensor &t = tensors_host[i];
    if (t.max_norm_penalty > 0.0f) {
      max_norm_rows_kernel<<<(t.rows + 3) / 4, 128, 0, ctx->stream>>>(t.w, t.rows, t.cols, t.max_norm_penalty);
      LAUNCH_CHECK(ctx);
    }
  }
  if (increment_count) return b200_counter_increment(ctx, count_dev);
  return B200_OK;
}
extern "C" int b200_optimizer_lookahead(b200_ctx *ctx, int ntensors, const b200_opt_tensor *tensors_dev,
                                        const b200_opt_tensor *tensors_host) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && tensors_dev && tensors_host, "NULL pointer");
  if (ntensors <= 0) return B200_OK;
  size_t max_n = 0;
  bool any = false;
  for (int i = 0; i < ntensors; ++i) {
    max_n = tensors_host[i].n > max_n ? (size_t)tensors_host[i].n : max_n;
    any = any || tensors_host[i].momentum > 0.0f;
  }
  if (!any) return B200_OK;
  size_t bx = (max_n + 128 * 4 - 1) / (128 * 4);
  const size_t cap = (size_t)ctx->sm_count * 8;
  if (bx > cap) bx = cap;
  lookahead_kernel<<<dim3((unsigned)(bx < 1 ? 1 : bx), (unsigned)ntensors), 128, 0, ctx->stream>>>(tensors_dev);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
