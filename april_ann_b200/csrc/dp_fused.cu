// Data-parallel update as ONE kernel over NVLink peer memory: gradient reduce-scatter + SGD on the
// rank's shard + all-gather of the updated weights.
//
// The reference is single-device (mathcore/c_src/gpu_helper.h:65-68); the NCCL path of csrc/nccl_dp.cu
// (all-reduce of the gradient buckets, then every replica updates every parameter) is the baseline:
// on 2 B200 it moved 17 MB in 55-62 us per bucket and every rank still streamed all of w, g, u.
// Here every rank owns 1/N of each tensor of a bucket:
//   g   = sum over ranks of the peers' gradients of the shard       (P2P loads, fixed rank order)
//   w,u = SGD step of optimizer_sgd.lua:65-95 on the shard          (only the owner keeps u current)
//   w  -> own replica and every peer's replica                       (P2P stores)
// so the wire carries (N-1)/N gradient bytes in and (N-1)/N weight bytes out per rank -- what an
// all-reduce carries -- the reduction needs no second pass over memory, and the update traffic per
// rank drops to 1/N.
//
// Cross-GPU ordering uses one 64-bit tag per (bucket, rank) in every rank's flag block, written with
// release.sys and polled with acquire.sys; the tag is the step number (device step counter + 1), so
// nothing is ever reset and the kernel can live in a replayed CUDA graph:
//   READY[b][r] = s   rank r's gradients of bucket b are final for step s AND r no longer reads the
//                     bucket's weights in step s (its data gradients are done) -> peers may read r's
//                     gradients and overwrite r's weights
//   DONE[b][r]  = s   rank r has written its shard of bucket b into every replica
// b200_dp_wait (end of the step) polls DONE of every bucket and rank: afterwards the replica is
// complete and every peer has finished reading this rank's gradients, which is what the next step's
// forward and backward need.
#include <float.h>

#include "common.cuh"

namespace {

constexpr int DP_THREADS = 1024;
constexpr int DP_UNROLL = 2;
constexpr int DP_SMEM = 120 * 1024;   // one CTA per SM, none beside a contraction CTA (see sgd.cu)
constexpr int DP_MAX_BUCKETS = 16;

__device__ __forceinline__ void st_release_sys(long long *p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long *p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long *ready_slot(long long *flags, int bucket, int rank) {
  return flags + (size_t)bucket * B200_DP_MAX_RANKS + rank;
}
__device__ __forceinline__ long long *done_slot(long long *flags, int bucket, int rank) {
  return flags + (size_t)(DP_MAX_BUCKETS + bucket) * B200_DP_MAX_RANKS + rank;
}

__global__ void __launch_bounds__(DP_THREADS, 1)
dp_fused_update_kernel(const b200_dp_group grp, const b200_sgd_tensor *__restrict__ tensors, int ntensors, double decay,
                       int64_t *count_dev, int bucket) {
  const int rank = grp.rank, nranks = grp.nranks;
  const int64_t count = *reinterpret_cast<volatile int64_t *>(count_dev);
  const long long tag = (long long)count + 1;
  // 1. announce (the launch is ordered after this rank's gradient kernels of the bucket and after the data
  //    gradients that read the bucket's weights), then wait for every peer's announcement
  if (blockIdx.x == 0 && threadIdx.x < nranks) st_release_sys(ready_slot(grp.flags[threadIdx.x], bucket, rank), tag);
  if (threadIdx.x < nranks) {
    const long long *f = ready_slot(grp.flags[rank], bucket, threadIdx.x);
    while (ld_acquire_sys(f) < tag) { }
  }
  __syncthreads();

  const double dec = 1.0 / (1.0 + decay * (double)count);
  const bool prune = (count % 100) == 0;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (int ti = 0; ti < ntensors; ++ti) {
    const b200_sgd_tensor t = tensors[ti];
    const float lrd = (float)((double)t.lr * dec);
    const float mt = t.momentum, l2 = t.weight_decay;
    const float l1 = (float)(((double)t.lr * dec) * (double)t.l1_norm);
    const bool has_l1 = t.l1_norm > 0.0f;
    auto upd = [&](float &w, float &g, float &u) {   // optimizer_sgd.lua:65-95, same order as sgd.cu
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
    // element offset of the tensor inside the arenas (identical layout on every rank); the arena pads
    // every tensor to 128 floats, so working on whole float4 never leaves the tensor's slot
    const size_t off = (size_t)(t.w - grp.weights[rank]);
    const size_t n4 = (t.n + 3) >> 2;
    const size_t per = (n4 + nranks - 1) / nranks;
    const size_t lo = min(n4, (size_t)rank * per), hi = min(n4, lo + per);
    float4 *u4 = reinterpret_cast<float4 *>(t.u);
    for (size_t i = lo + tid; i < hi; i += DP_UNROLL * nth) {
      float4 g[DP_UNROLL], w[DP_UNROLL], u[DP_UNROLL];
      bool ok[DP_UNROLL];
#pragma unroll
      for (int j = 0; j < DP_UNROLL; ++j) {
        const size_t e = i + (size_t)j * nth;
        ok[j] = e < hi;
        const size_t ec = ok[j] ? e : lo;
        // every load of the iteration is issued before the first use: (nranks + 2) x DP_UNROLL independent
        // 16-byte loads per thread, the remote ones with NVLink latency
        g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 part[B200_DP_MAX_RANKS];
#pragma unroll
        for (int p = 0; p < B200_DP_MAX_RANKS; ++p)
          if (p < nranks) part[p] = __ldcs(reinterpret_cast<const float4 *>(grp.grads[p] + off) + ec);
        w[j] = *(reinterpret_cast<const float4 *>(grp.weights[rank] + off) + ec);
        u[j] = __ldcs(u4 + ec);
#pragma unroll
        for (int p = 0; p < B200_DP_MAX_RANKS; ++p)   // fixed rank order: the sum does not depend on who owns the shard
          if (p < nranks) { g[j].x += part[p].x; g[j].y += part[p].y; g[j].z += part[p].z; g[j].w += part[p].w; }
      }
#pragma unroll
      for (int j = 0; j < DP_UNROLL; ++j) {
        if (!ok[j]) continue;
        const size_t e = i + (size_t)j * nth;
        upd(w[j].x, g[j].x, u[j].x); upd(w[j].y, g[j].y, u[j].y); upd(w[j].z, g[j].z, u[j].z); upd(w[j].w, g[j].w, u[j].w);
        __stcs(u4 + e, u[j]);
#pragma unroll
        for (int p = 0; p < B200_DP_MAX_RANKS; ++p)
          if (p < nranks) *(reinterpret_cast<float4 *>(grp.weights[p] + off) + e) = w[j];
      }
    }
  }
  // 2. this rank's shard is in every replica: publish
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long *ticket = reinterpret_cast<unsigned long long *>(count_dev + 1);
    const bool last = atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1;
    if (last) {
      *ticket = 0ull;
      __threadfence_system();
      for (int p = 0; p < nranks; ++p) st_release_sys(done_slot(grp.flags[p], bucket, rank), tag);
    }
  }
}

__global__ void dp_wait_kernel(const b200_dp_group grp, int nbuckets, const int64_t *count_dev) {
  const long long tag = (long long)*count_dev + 1;
  for (int i = threadIdx.x; i < nbuckets * grp.nranks; i += blockDim.x) {
    const long long *f = done_slot(grp.flags[grp.rank], i / grp.nranks, i % grp.nranks);
    while (ld_acquire_sys(f) < tag) { }
  }
}

}  // namespace

extern "C" int b200_ipc_export(b200_ctx *ctx, void *dptr, void *handle64) {
  ARG_CHECK(ctx && dptr && handle64, "NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, dptr));
  memcpy(handle64, &h, 64);
  return B200_OK;
}
extern "C" int b200_ipc_import(b200_ctx *ctx, const void *handle64, void **dptr) {
  ARG_CHECK(ctx && handle64 && dptr, "NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return B200_OK;
}
extern "C" int b200_ipc_close(b200_ctx *ctx, void *dptr) {
  ARG_CHECK(ctx, "NULL pointer");
  if (dptr) CUDA_TRY(cudaIpcCloseMemHandle(dptr));
  return B200_OK;
}
extern "C" size_t b200_dp_flags_bytes(void) { return sizeof(long long) * 2 * DP_MAX_BUCKETS * B200_DP_MAX_RANKS; }

extern "C" int b200_dp_fused_update(b200_ctx *ctx, const b200_dp_group *grp, int ntensors, const b200_sgd_tensor *tensors_dev,
                                    const b200_sgd_tensor *tensors_host, double decay, int64_t *count_dev, int bucket) {
  ARG_CHECK(ctx && grp && tensors_dev && tensors_host && count_dev, "NULL pointer");
  ARG_CHECK(grp->nranks >= 2 && grp->nranks <= B200_DP_MAX_RANKS && grp->rank >= 0 && grp->rank < grp->nranks, "bad replica group");
  ARG_CHECK(bucket >= 0 && bucket < DP_MAX_BUCKETS, "bucket out of range");
  if (ntensors <= 0) return B200_OK;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(dp_fused_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM));
    attr = true;
  }
  size_t shard4 = 0;   // float4 of the biggest shard: enough CTAs to cover it once, at most the SMs planned for
  for (int i = 0; i < ntensors; ++i) {
    ARG_CHECK(tensors_host[i].max_norm_penalty <= 0.0f, "max_norm_penalty needs whole rows: use the all-reduce path");
    const size_t per = (((size_t)tensors_host[i].n + 3) / 4 + grp->nranks - 1) / grp->nranks;
    shard4 = per > shard4 ? per : shard4;
  }
  const int sms = ctx->sm_budget > 0 ? ctx->sm_budget : ctx->sm_count;
  size_t blocks = (shard4 + (size_t)DP_THREADS * DP_UNROLL - 1) / ((size_t)DP_THREADS * DP_UNROLL);
  if (blocks > (size_t)sms) blocks = (size_t)sms;
  if (blocks < 1) blocks = 1;
  dp_fused_update_kernel<<<(unsigned)blocks, DP_THREADS, DP_SMEM, ctx->stream>>>(*grp, tensors_dev, ntensors, decay, count_dev, bucket);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

extern "C" int b200_dp_wait(b200_ctx *ctx, const b200_dp_group *grp, int nbuckets, const int64_t *count_dev) {
  ARG_CHECK(ctx && grp && count_dev, "NULL pointer");
  ARG_CHECK(nbuckets >= 0 && nbuckets <= DP_MAX_BUCKETS, "bucket count out of range");
  if (nbuckets == 0) return B200_OK;
  dp_wait_kernel<<<1, 128, 0, ctx->stream>>>(*grp, nbuckets, count_dev);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
