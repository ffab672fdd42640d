// Multi-tensor SGD with momentum, L2 weight decay, L1 truncation, max-norm penalty and
// the learning-rate decay of the reference -- ONE launch for all parameters instead of the
// 4-7 Lua->C++ matrix calls per tensor of ann.optimizer.sgd:execute
// (ann/optimizer/lua_src/optimizer_sgd.lua:50-100; helpers base_optimizer.lua:28-49).
//
// Per element, in the reference's order:
//   g += l2*w                      (:72, only if l2 > 0)
//   u  = mt>0 ? mt*u : 0           (:74)
//   u += lrd*g                     (:76)   lrd = lr / (1 + decay*count)   (:61-62,69-70)
//   w -= u                         (:78)
//   L1: z=|w|>l1' ; w -= l1'*sign(w) ; u -= l1'*sign(w) ; w *= z   with l1' = lrd*l1  (:80, base_optimizer.lua:28-42)
//   every 100 updates: subnormal weights -> 0   (:85-87)
// HBM traffic: read w,g,u + write w,u = 20 B/parameter (+4 when the regularised gradient
// is written back, which the reference does in place).
#include <float.h>

#include "common.cuh"

namespace {

constexpr int TPB = 128;          // light launches (small tensors, many tensors per launch)
constexpr int TPB_BIG = 1024;     // one big tensor: one CTA per SM
constexpr int UNROLL = 4;
// A big-tensor CTA asks for more (unused) dynamic shared memory than half an SM, so that exactly one is
// resident per SM and none fits beside a contraction CTA: an update that runs next to a contraction then
// owns the SMs the contraction leaves free (ctx->sm_budget tells how many) instead of sharing SMs with it.
// Sharing was measured: the co-resident update competes for the SM's path to L2 and slowed the
// weight-gradient contraction from 17 to 31 us.
constexpr int SGD_BIG_SMEM = 120 * 1024;

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 65536 / (64 * THREADS)) sgd_kernel(const b200_sgd_tensor *__restrict__ tensors, double decay,
                                                                           int64_t *count_dev, int flags) {
  const b200_sgd_tensor t = tensors[blockIdx.y];
  const int64_t count = *reinterpret_cast<volatile int64_t *>(count_dev);
  const int write_back_grad = flags & B200_SGD_WRITE_BACK_GRAD;
  const double dec = 1.0 / (1.0 + decay * (double)count);
  const float lrd = (float)((double)t.lr * dec);
  const float mt = t.momentum, l2 = t.weight_decay;
  const float l1 = (float)(((double)t.lr * dec) * (double)t.l1_norm);
  const bool prune = (count % 100) == 0;
  const bool has_l1 = t.l1_norm > 0.0f;

  auto upd = [&](float &w, float &g, float &u) {
    if (l2 > 0.0f) g = fmaf(l2, w, g);
    u = (mt > 0.0f) ? mt * u : 0.0f;
    u = fmaf(lrd, g, u);
    w -= u;
    if (has_l1) {
      const float z = fabsf(w) > l1 ? 1.0f : 0.0f;
      const float s = (w > 0.0f) ? l1 : (w < 0.0f ? -l1 : 0.0f);
      w -= s;
      u -= s;
      w *= z;
    }
    if (prune && fabsf(w) < FLT_MIN) w = 0.0f;
  };

  const size_t n = t.n;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  const bool vec = ((((uintptr_t)t.w) | ((uintptr_t)t.g) | ((uintptr_t)t.u)) & 15) == 0;
  if (vec) {
    const size_t n4 = n >> 2;
    float4 *w4 = reinterpret_cast<float4 *>(t.w);
    float4 *g4 = reinterpret_cast<float4 *>(t.g);
    float4 *u4 = reinterpret_cast<float4 *>(t.u);
    // UNROLL independent (w, g, u) triples per thread are loaded before the first use: 12 x 16 bytes in
    // flight per thread, so that the one small CTA per SM that fits beside a contraction CTA still keeps
    // the memory system busy
    size_t i = tid;
    for (; i + (UNROLL - 1) * nth < n4; i += UNROLL * nth) {
      float4 w[UNROLL], g[UNROLL], u[UNROLL];
#pragma unroll
      // w stays in L2 for the next forward pass; g and u are touched once per step: streaming loads / stores
      for (int j = 0; j < UNROLL; ++j) { w[j] = w4[i + j * nth]; g[j] = __ldcs(g4 + i + j * nth); u[j] = __ldcs(u4 + i + j * nth); }
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        upd(w[j].x, g[j].x, u[j].x); upd(w[j].y, g[j].y, u[j].y); upd(w[j].z, g[j].z, u[j].z); upd(w[j].w, g[j].w, u[j].w);
        w4[i + j * nth] = w[j];
        __stcs(u4 + i + j * nth, u[j]);
        if (write_back_grad) __stcs(g4 + i + j * nth, g[j]);
      }
    }
    for (; i < n4; i += nth) {
      float4 w = w4[i], g = g4[i], u = u4[i];
      upd(w.x, g.x, u.x); upd(w.y, g.y, u.y); upd(w.z, g.z, u.z); upd(w.w, g.w, u.w);
      w4[i] = w;
      u4[i] = u;
      if (write_back_grad) g4[i] = g;
    }
    for (size_t i2 = (n4 << 2) + tid; i2 < n; i2 += nth) {
      float w = t.w[i2], g = t.g[i2], u = t.u[i2];
      upd(w, g, u);
      t.w[i2] = w; t.u[i2] = u;
      if (write_back_grad) t.g[i2] = g;
    }
  } else {
    for (size_t i = tid; i < n; i += nth) {
      float w = t.w[i], g = t.g[i], u = t.u[i];
      upd(w, g, u);
      t.w[i] = w; t.u[i] = u;
      if (write_back_grad) t.g[i] = g;
    }
  }
  if (flags & B200_SGD_INCREMENT_COUNT) {
    // last update launch of the step: the last CTA to finish bumps the step counter.  Every CTA has read
    // `count` before taking its ticket, and no other update launch of this step is still running.
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned long long *ticket = reinterpret_cast<unsigned long long *>(count_dev + 1);
      const unsigned long long total = (unsigned long long)gridDim.x * gridDim.y;
      if (atomicAdd(ticket, 1ull) == total - 1) {
        *ticket = 0ull;
        *count_dev = count + 1;
      }
    }
  }
}

// max_norm_penalty (base_optimizer.lua:44-49): every row of w with ||row||_2 > mnp is
// rescaled to norm mnp.  One warp per row.
__global__ void __launch_bounds__(TPB) max_norm_kernel(float *__restrict__ w, int rows, int cols, float mnp) {
  const int row = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5);
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

extern "C" int b200_sgd_multi_tensor(b200_ctx *ctx, int ntensors, const b200_sgd_tensor *tensors_dev,
                                     const b200_sgd_tensor *tensors_host, double decay,
                                     const int64_t *count_dev, int write_back_grad) {
  B200_ENTER(ctx);
  return b200_sgd_multi_tensor_ex(ctx, ntensors, tensors_dev, tensors_host, decay, const_cast<int64_t *>(count_dev),
                                  write_back_grad ? B200_SGD_WRITE_BACK_GRAD : 0);
}

extern "C" int b200_sgd_multi_tensor_ex(b200_ctx *ctx, int ntensors, const b200_sgd_tensor *tensors_dev,
                                        const b200_sgd_tensor *tensors_host, double decay, int64_t *count_dev,
                                        int flags) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && tensors_dev && tensors_host && count_dev, "NULL pointer");
  if (ntensors <= 0) return B200_OK;
  size_t max_n = 0;
  for (int i = 0; i < ntensors; ++i) max_n = tensors_host[i].n > max_n ? (size_t)tensors_host[i].n : max_n;
  if (ntensors == 1 && max_n >= (1u << 20)) {
    // one big tensor: one 1024-thread CTA per SM, grid-stride inside (1024 x 12 x 16 B in flight per SM)
    if (ONCE_PER_DEVICE(ctx))
      CUDA_TRY(cudaFuncSetAttribute(sgd_kernel<TPB_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SGD_BIG_SMEM));
    const int sms = ctx->sm_budget > 0 ? ctx->sm_budget : ctx->sm_count;
    size_t blocks = (max_n / 4 + (size_t)TPB_BIG * UNROLL - 1) / ((size_t)TPB_BIG * UNROLL);
    if (blocks > (size_t)sms) blocks = (size_t)sms;
    sgd_kernel<TPB_BIG><<<dim3((unsigned)blocks, 1), TPB_BIG, SGD_BIG_SMEM, ctx->stream>>>(tensors_dev, decay, count_dev, flags);
    LAUNCH_CHECK(ctx);
  } else {
    size_t blocks_x = (max_n / 4 + (size_t)TPB * UNROLL - 1) / ((size_t)TPB * UNROLL);
    // persistent-ish: cap at 4 CTAs per SM worth of blocks per tensor
    size_t cap = (size_t)ctx->sm_count * 4;
    if (blocks_x > cap) blocks_x = cap;
    if (blocks_x < 1) blocks_x = 1;
    dim3 grid((unsigned)blocks_x, (unsigned)ntensors);
    PREFER_MAX_SMEM_ONCE(sgd_kernel<TPB>);
    sgd_kernel<TPB><<<grid, TPB, 0, ctx->stream>>>(tensors_dev, decay, count_dev, flags);
    LAUNCH_CHECK(ctx);
  }
  for (int i = 0; i < ntensors; ++i) {
    const b200_sgd_tensor &t = tensors_host[i];
    if (t.max_norm_penalty > 0.0f) {
      max_norm_kernel<<<(t.rows + TPB / 32 - 1) / (TPB / 32), TPB, 0, ctx->stream>>>(t.w, t.rows, t.cols,
                                                                                 t.max_norm_penalty);
      LAUNCH_CHECK(ctx);
    }
  }
  return B200_OK;
}
