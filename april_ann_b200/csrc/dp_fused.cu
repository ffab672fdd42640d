// Data-parallel update as ONE kernel over NVLink peer memory: gradient reduce-scatter + SGD on the
// rank's shard + all-gather of the updated weights.
//
// The reference is single-device (mathcore/c_src/gpu_helper.h:65-68); the NCCL path of csrc/nccl_dp.cu
// (all-reduce of the gradient buckets, then every replica updates every parameter) is the baseline:
// on 2 B200 it moved 17 MB in 55-62 us per bucket and every rank still streamed all of w, g, u.
// Here every rank owns 1/N of each tensor of a bucket:
//   g   = sum over ranks of the gradients of the shard               (P2P loads from the peers, fixed rank order)
//   w,u = SGD step of optimizer_sgd.lua:65-95 on the shard          (only the owner keeps u current)
//   w  -> own replica and every peer's replica                       (P2P stores)
// so the wire carries (N-1)/N gradient bytes in and (N-1)/N weight bytes out per rank -- what an
// all-reduce carries -- the reduction needs no second pass over memory, and the update traffic per
// rank drops to 1/N.
//
// Cross-GPU ordering uses one 64-bit tag per (bucket, rank) in every rank's flag block, written with
// release.sys and polled with acquire.sys; the tag is the replica group's epoch + 1 (count_dev[2], advanced by
// b200_dp_wait at the end of every step and never set back -- unlike the optimizer's step count count_dev[0],
// which a checkpoint restore rewrites), so nothing is ever reset and the kernel can live in a replayed CUDA graph:
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

constexpr int DP_THREADS = 512;      // one CTA per SM, 128 registers per thread: (NR + 2) x UNROLL 16-byte loads in flight each
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
__device__ __forceinline__ long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}
// bring-up: per bucket {kernel start, peers ready, shard written, published} globaltimer stamps of this rank
__device__ __forceinline__ long long *dbg_slot(long long *flags, int bucket, int i) {
  return flags + (size_t)2 * DP_MAX_BUCKETS * B200_DP_MAX_RANKS + bucket * 4 + i;
}
__device__ __forceinline__ long long *done_slot(long long *flags, int bucket, int rank) {
  return flags + (size_t)(DP_MAX_BUCKETS + bucket) * B200_DP_MAX_RANKS + rank;
}

// tickets of the two phases live behind the debug stamps of this rank's flag block
__device__ __forceinline__ unsigned long long *ticket_slot(long long *flags, int i) {
  return reinterpret_cast<unsigned long long *>(flags + (size_t)2 * DP_MAX_BUCKETS * B200_DP_MAX_RANKS + 4 * DP_MAX_BUCKETS + i);
}

template <int NR, int DP_UNROLL>
__global__ void __launch_bounds__(DP_THREADS, 1)
dp_fused_update_kernel(const b200_dp_group grp, const b200_sgd_tensor *__restrict__ tensors, int ntensors, double decay,
                       int64_t *count_dev, int bucket, int dbg) {
  const int rank = grp.rank, nranks = grp.nranks;
  const int64_t count = *reinterpret_cast<volatile int64_t *>(count_dev);
  const long long tag = (long long)*reinterpret_cast<volatile int64_t *>(count_dev + 2) + 1;   // replica-group epoch, not the optimizer's count
  if (blockIdx.x == 0 && threadIdx.x == 0) *dbg_slot(grp.flags[rank], bucket, 0) = gtime();

  // Shard q of a tensor = float4 [q*per, (q+1)*per) of its slot; the shards of all tensors of the bucket that
  // belong to one rank form one index space that the grid strides over once.
  constexpr int MAXT = 32;
  __shared__ unsigned long long s_cum[MAXT + 1], s_lo[MAXT], s_off4[MAXT];
  __shared__ float4 *s_u[MAXT];
  __shared__ float s_lrd[MAXT], s_mt[MAXT], s_l2[MAXT], s_l1[MAXT];
  const double dec = 1.0 / (1.0 + decay * (double)count);
  const bool prune = (count % 100) == 0;
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (unsigned long long)gridDim.x * blockDim.x;
  auto plan = [&](int q) {   // index space of rank q's shards
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long cum = 0;
      for (int ti = 0; ti < ntensors; ++ti) {
        const b200_sgd_tensor t = tensors[ti];
        // the arena pads every tensor to 128 floats, so working on whole float4 never leaves the tensor's slot
        const size_t n4 = (t.n + 3) >> 2;
        const size_t per = (n4 + nranks - 1) / nranks;
        const size_t lo = min(n4, (size_t)q * per), hi = min(n4, lo + per);
        s_cum[ti] = cum;
        s_lo[ti] = lo;
        s_off4[ti] = (unsigned long long)(t.w - grp.weights[rank]) >> 2;   // float4 offset of the slot inside the arenas
        s_u[ti] = reinterpret_cast<float4 *>(t.u);
        s_lrd[ti] = (float)((double)t.lr * dec);
        s_mt[ti] = t.momentum;
        s_l2[ti] = t.weight_decay;
        s_l1[ti] = t.l1_norm > 0.0f ? (float)(((double)t.lr * dec) * (double)t.l1_norm) : 0.0f;
        cum += hi - lo;
      }
      s_cum[ntensors] = cum;
    }
    __syncthreads();
  };
  auto locate = [&](unsigned long long gc, int &ti) {   // flat index -> float4 index inside the arenas
    ti = 0;
    while (ti + 1 < ntensors && gc >= s_cum[ti + 1]) ++ti;
    return s_off4[ti] + s_lo[ti] + (gc - s_cum[ti]);
  };
  const float4 *own_g = reinterpret_cast<const float4 *>(grp.grads[rank]);

  // ---- announce: this rank's gradients of the bucket are final and its data gradients no longer read the
  //      bucket's weights (the launch is ordered after both).  Every CTA announces -- the same tag, idempotent --
  //      so the announcement never waits for one particular CTA to be resident.
  if (threadIdx.x < nranks) st_release_sys(ready_slot(grp.flags[threadIdx.x], bucket, rank), tag);
  // ---- wait for every peer's announcement (all CTAs are resident: one per SM, grid <= SM count)
  if (threadIdx.x < nranks) {
    const long long *f = ready_slot(grp.flags[rank], bucket, threadIdx.x);
    while (ld_acquire_sys(f) < tag) { }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) *dbg_slot(grp.flags[rank], bucket, 1) = gtime();

  // ---- this rank's shards: pull the peers' gradients, sum, SGD step, new weights into every replica
  plan(rank);
  const unsigned long long total = s_cum[ntensors];
  for (unsigned long long i = tid; i < total; i += DP_UNROLL * nth) {
    float4 g[DP_UNROLL], w[DP_UNROLL], u[DP_UNROLL];
    int tix[DP_UNROLL];
    unsigned long long el[DP_UNROLL];
    bool ok[DP_UNROLL];
#pragma unroll
    for (int j = 0; j < DP_UNROLL; ++j) {
      const unsigned long long gi = i + (unsigned long long)j * nth;
      ok[j] = gi < total;
      el[j] = locate(ok[j] ? gi : 0ull, tix[j]);
    }
#pragma unroll
    for (int j = 0; j < DP_UNROLL; ++j) {
      // every load of the iteration is issued before the first use: (NR + 2) x DP_UNROLL independent 16-byte loads
      float4 part[NR];
#pragma unroll
      for (int p = 0; p < NR; ++p) {
        const float4 *src = (p == rank || p >= nranks || (dbg & 1)) ? own_g : reinterpret_cast<const float4 *>(grp.grads[p]);
        part[p] = __ldcs(src + el[j]);
      }
      w[j] = *(reinterpret_cast<const float4 *>(grp.weights[rank]) + el[j]);
      u[j] = __ldcs(s_u[tix[j]] + (el[j] - s_off4[tix[j]]));
      g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < NR; ++p)   // fixed rank order
        if (p < nranks) { g[j].x += part[p].x; g[j].y += part[p].y; g[j].z += part[p].z; g[j].w += part[p].w; }
    }
#pragma unroll
    for (int j = 0; j < DP_UNROLL; ++j) {
      if (!ok[j]) continue;
      const float lrd = s_lrd[tix[j]], mt = s_mt[tix[j]], l2 = s_l2[tix[j]], l1 = s_l1[tix[j]];
      auto upd = [&](float &wv, float &gv, float &uv) {   // optimizer_sgd.lua:65-95, same order as sgd.cu
        if (l2 > 0.0f) gv = fmaf(l2, wv, gv);
        uv = (mt > 0.0f) ? mt * uv : 0.0f;
        uv = fmaf(lrd, gv, uv);
        wv -= uv;
        if (l1 > 0.0f) {
          const float z = fabsf(wv) > l1 ? 1.0f : 0.0f;
          const float sg = (wv > 0.0f) ? l1 : (wv < 0.0f ? -l1 : 0.0f);
          wv -= sg;
          uv -= sg;
          wv *= z;
        }
        if (prune && fabsf(wv) < FLT_MIN) wv = 0.0f;
      };
      upd(w[j].x, g[j].x, u[j].x); upd(w[j].y, g[j].y, u[j].y); upd(w[j].z, g[j].z, u[j].z); upd(w[j].w, g[j].w, u[j].w);
      __stcs(s_u[tix[j]] + (el[j] - s_off4[tix[j]]), u[j]);
#pragma unroll
      for (int p = 0; p < NR; ++p)
        if (p < nranks && (p == rank || !(dbg & 2))) *(reinterpret_cast<float4 *>(grp.weights[p]) + el[j]) = w[j];
    }
  }
  // ---- this rank's shard is in every replica: publish.  The last CTA to get here tells every peer, one
  //      thread per peer (a single thread doing nranks release stores to remote GPUs pays each round trip in turn)
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long *ticket = ticket_slot(grp.flags[rank], 1);
    s_last = atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1;
    if (s_last) {
      *ticket = 0ull;
      *dbg_slot(grp.flags[rank], bucket, 2) = gtime();
    }
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < nranks) st_release_sys(done_slot(grp.flags[threadIdx.x], bucket, rank), tag);
    if (threadIdx.x == 0) *dbg_slot(grp.flags[rank], bucket, 3) = gtime();
  }
}

// ---------------------------------------------------------------- NVLS (in-switch) variant
// With the gradient and weight arenas of every rank bound to one multicast object (NVLink SHARP / NVLS, set up by the
// host through CUDA's multicast API: cuMulticastCreate / cuMulticastBindMem, done for us by torch's symmetric memory
// allocator, april_ann_b200/parallel.py), the reduce-scatter and the all-gather become ONE instruction each per
// 16 bytes:
//   multimem.ld_reduce.add.v4.f32 [mc_grads + i]   the NVSwitch reads the 16 bytes from every replica, adds them
//                                                  and returns the sum -- instead of (N-1) P2P loads per element
//   multimem.st.v4.f32 [mc_weights + i], w         the NVSwitch writes the new weights into every replica
//                                                  -- instead of (N-1) P2P stores per element
// so a rank issues 1/N of the loads of the pull kernel above and the switch does the fan-in / fan-out.  Same tags,
// same shard plan, same SGD arithmetic.  The summation order inside the switch is the hardware's; every element
// is reduced exactly once (by its owner), so the replicas stay bit-identical.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4 *mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4 *mc, const float4 &v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <int DP_UNROLL>
__global__ void __launch_bounds__(DP_THREADS, 1)
dp_mc_update_kernel(const b200_dp_group grp, const b200_sgd_tensor *__restrict__ tensors, int ntensors, double decay,
                    int64_t *count_dev, int bucket) {
  const int rank = grp.rank, nranks = grp.nranks;
  const int64_t count = *reinterpret_cast<volatile int64_t *>(count_dev);
  const long long tag = (long long)*reinterpret_cast<volatile int64_t *>(count_dev + 2) + 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) *dbg_slot(grp.flags[rank], bucket, 0) = gtime();
  constexpr int MAXT = 32;
  __shared__ unsigned long long s_cum[MAXT + 1], s_lo[MAXT], s_off4[MAXT];
  __shared__ float4 *s_u[MAXT];
  __shared__ float s_lrd[MAXT], s_mt[MAXT], s_l2[MAXT], s_l1[MAXT];
  const double dec = 1.0 / (1.0 + decay * (double)count);
  const bool prune = (count % 100) == 0;
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (unsigned long long)gridDim.x * blockDim.x;
  if (threadIdx.x == 0) {
    unsigned long long cum = 0;
    for (int ti = 0; ti < ntensors; ++ti) {
      const b200_sgd_tensor t = tensors[ti];
      const size_t n4 = (t.n + 3) >> 2;
      const size_t per = (n4 + nranks - 1) / nranks;
      const size_t lo = min(n4, (size_t)rank * per), hi = min(n4, lo + per);
      s_cum[ti] = cum;
      s_lo[ti] = lo;
      s_off4[ti] = (unsigned long long)(t.w - grp.weights[rank]) >> 2;
      s_u[ti] = reinterpret_cast<float4 *>(t.u);
      s_lrd[ti] = (float)((double)t.lr * dec);
      s_mt[ti] = t.momentum;
      s_l2[ti] = t.weight_decay;
      s_l1[ti] = t.l1_norm > 0.0f ? (float)(((double)t.lr * dec) * (double)t.l1_norm) : 0.0f;
      cum += hi - lo;
    }
    s_cum[ntensors] = cum;
  }
  // announce (gradients final, weights of the bucket no longer read by this rank's data gradients), wait for every peer
  if (threadIdx.x < nranks) st_release_sys(ready_slot(grp.flags[threadIdx.x], bucket, rank), tag);
  if (threadIdx.x < nranks) {
    const long long *f = ready_slot(grp.flags[rank], bucket, threadIdx.x);
    while (ld_acquire_sys(f) < tag) { }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) *dbg_slot(grp.flags[rank], bucket, 1) = gtime();
  auto locate = [&](unsigned long long gc, int &ti) {
    ti = 0;
    while (ti + 1 < ntensors && gc >= s_cum[ti + 1]) ++ti;
    return s_off4[ti] + s_lo[ti] + (gc - s_cum[ti]);
  };
  const float4 *mc_g = reinterpret_cast<const float4 *>(grp.mc_grads);
  float4 *mc_w = reinterpret_cast<float4 *>(grp.mc_weights);
  const float4 *own_w = reinterpret_cast<const float4 *>(grp.weights[rank]);
  const unsigned long long total = s_cum[ntensors];
  for (unsigned long long i = tid; i < total; i += DP_UNROLL * nth) {
    float4 g[DP_UNROLL], w[DP_UNROLL], u[DP_UNROLL];
    int tix[DP_UNROLL];
    unsigned long long el[DP_UNROLL];
    bool ok[DP_UNROLL];
#pragma unroll
    for (int j = 0; j < DP_UNROLL; ++j) {
      const unsigned long long gi = i + (unsigned long long)j * nth;
      ok[j] = gi < total;
      el[j] = locate(ok[j] ? gi : 0ull, tix[j]);
    }
#pragma unroll
    for (int j = 0; j < DP_UNROLL; ++j) {   // 3 x DP_UNROLL independent 16-byte loads in flight
      g[j] = multimem_ld_reduce_add(mc_g + el[j]);
      w[j] = own_w[el[j]];
      u[j] = __ldcs(s_u[tix[j]] + (el[j] - s_off4[tix[j]]));
    }
#pragma unroll
    for (int j = 0; j < DP_UNROLL; ++j) {
      if (!ok[j]) continue;
      const float lrd = s_lrd[tix[j]], mt = s_mt[tix[j]], l2 = s_l2[tix[j]], l1 = s_l1[tix[j]];
      auto upd = [&](float &wv, float &gv, float &uv) {   // optimizer_sgd.lua:65-95, same order as sgd.cu
        if (l2 > 0.0f) gv = fmaf(l2, wv, gv);
        uv = (mt > 0.0f) ? mt * uv : 0.0f;
        uv = fmaf(lrd, gv, uv);
        wv -= uv;
        if (l1 > 0.0f) {
          const float z = fabsf(wv) > l1 ? 1.0f : 0.0f;
          const float sg = (wv > 0.0f) ? l1 : (wv < 0.0f ? -l1 : 0.0f);
          wv -= sg;
          uv -= sg;
          wv *= z;
        }
        if (prune && fabsf(wv) < FLT_MIN) wv = 0.0f;
      };
      upd(w[j].x, g[j].x, u[j].x); upd(w[j].y, g[j].y, u[j].y); upd(w[j].z, g[j].z, u[j].z); upd(w[j].w, g[j].w, u[j].w);
      __stcs(s_u[tix[j]] + (el[j] - s_off4[tix[j]]), u[j]);
      multimem_st(mc_w + el[j], w[j]);      // into every replica, this rank's own included
    }
  }
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long *ticket = ticket_slot(grp.flags[rank], 1);
    s_last = atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1;
    if (s_last) {
      *ticket = 0ull;
      *dbg_slot(grp.flags[rank], bucket, 2) = gtime();
    }
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < nranks) st_release_sys(done_slot(grp.flags[threadIdx.x], bucket, rank), tag);
    if (threadIdx.x == 0) *dbg_slot(grp.flags[rank], bucket, 3) = gtime();
  }
}

// ---------------------------------------------------------------- copy-engine variant
// The same protocol with the NVLink traffic on the copy engines instead of SM loads / stores: the update
// kernel shrinks to a local pass (16 us for the 17 MB bucket on 74 SMs) and the SMs stay free, but every
// peer copy inside the step graph costs ~15-20 us whatever its size, so at the C2 sizes (23 MB of
// gradients) the step is slower (243 us against 191 us on 2 B200) and the SM variant is the default;
// B200_DP_DMA=1 selects this one (big models, where the copies hide behind seconds of contractions).
//   1. cudaMemcpyAsync: this rank's gradients of every peer's slice -> the peer's receive block
//   2. dp_signal_kernel: READY tags to every peer (ordered after the copies on the stream)
//   3. dp_shard_update_kernel: waits for the peers' READY, sums own + received gradients of its slice
//      (local loads, fixed rank order), SGD step, new weights into its own replica
//   4. cudaMemcpyAsync: the updated slice of the weights -> every peer's replica
//   5. dp_signal_kernel: DONE tags to every peer
// A bucket is a contiguous range of the arenas; rank q owns the q-th of nranks slices of it (128-float
// granules), so every copy is one contiguous block.
__global__ void dp_signal_kernel(const b200_dp_group grp, int bucket, int done, const int64_t *count_dev) {
  const long long tag = (long long)count_dev[2] + 1;
  if (threadIdx.x < grp.nranks) {
    long long *f = done ? done_slot(grp.flags[threadIdx.x], bucket, grp.rank) : ready_slot(grp.flags[threadIdx.x], bucket, grp.rank);
    __threadfence_system();
    st_release_sys(f, tag);
  }
}

template <int NR, int UN>
__global__ void __launch_bounds__(DP_THREADS, 1)
dp_shard_update_kernel(const b200_dp_group grp, const b200_sgd_tensor *__restrict__ tensors, int ntensors, double decay,
                       const int64_t *count_dev, int bucket, unsigned long long lo4, unsigned long long hi4) {
  const int rank = grp.rank, nranks = grp.nranks;
  const int64_t count = *count_dev;
  const long long tag = (long long)count_dev[2] + 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) *dbg_slot(grp.flags[rank], bucket, 0) = gtime();
  constexpr int MAXT = 32;
  __shared__ unsigned long long s_beg4[MAXT];       // first float4 of every tensor's slot inside the arenas
  __shared__ float s_lrd[MAXT], s_mt[MAXT], s_l2[MAXT], s_l1[MAXT];
  const double dec = 1.0 / (1.0 + decay * (double)count);
  const bool prune = (count % 100) == 0;
  if (threadIdx.x == 0) {
    for (int ti = 0; ti < ntensors; ++ti) {
      const b200_sgd_tensor t = tensors[ti];
      s_beg4[ti] = (unsigned long long)(t.w - grp.weights[rank]) >> 2;
      s_lrd[ti] = (float)((double)t.lr * dec);
      s_mt[ti] = t.momentum;
      s_l2[ti] = t.weight_decay;
      s_l1[ti] = t.l1_norm > 0.0f ? (float)(((double)t.lr * dec) * (double)t.l1_norm) : 0.0f;
    }
  }
  // the peers' gradients of this slice have landed (their copies are ordered before their READY tags)
  if (threadIdx.x < nranks) {
    const long long *f = ready_slot(grp.flags[rank], bucket, threadIdx.x);
    while (ld_acquire_sys(f) < tag) { }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) *dbg_slot(grp.flags[rank], bucket, 1) = gtime();
  const unsigned long long arena4 = grp.arena_elems >> 2;
  const float4 *own_g = reinterpret_cast<const float4 *>(grp.grads[rank]);
  const float4 *recv = reinterpret_cast<const float4 *>(grp.recv[rank]);
  float4 *w4 = reinterpret_cast<float4 *>(grp.weights[rank]);
  // the three arenas share one layout: the update arena starts where the first tensor's u would be at offset 0
  float4 *u4 = reinterpret_cast<float4 *>(tensors[0].u) - s_beg4[0];
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = lo4 + tid; i < hi4; i += UN * nth) {
    float4 g[UN], w[UN], u[UN];
    bool ok[UN];
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      const unsigned long long e = i + (unsigned long long)j * nth;
      ok[j] = e < hi4;
      const unsigned long long ec = ok[j] ? e : lo4;
      float4 part[NR];
#pragma unroll
      for (int p = 0; p < NR; ++p) {
        const float4 *src = (p == rank || p >= nranks) ? own_g : recv + (unsigned long long)(p < rank ? p : p - 1) * arena4;
        part[p] = __ldcs(src + ec);
      }
      w[j] = w4[ec];
      u[j] = __ldcs(u4 + ec);
      g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < NR; ++p)   // fixed rank order
        if (p < nranks) { g[j].x += part[p].x; g[j].y += part[p].y; g[j].z += part[p].z; g[j].w += part[p].w; }
    }
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      if (!ok[j]) continue;
      const unsigned long long e = i + (unsigned long long)j * nth;
      int ti = 0;
      while (ti + 1 < ntensors && e >= s_beg4[ti + 1]) ++ti;   // (padding behind a tensor takes its options; it holds zeros)
      const float lrd = s_lrd[ti], mt = s_mt[ti], l2 = s_l2[ti], l1 = s_l1[ti];
      auto upd = [&](float &wv, float &gv, float &uv) {   // optimizer_sgd.lua:65-95, same order as sgd.cu
        if (l2 > 0.0f) gv = fmaf(l2, wv, gv);
        uv = (mt > 0.0f) ? mt * uv : 0.0f;
        uv = fmaf(lrd, gv, uv);
        wv -= uv;
        if (l1 > 0.0f) {
          const float z = fabsf(wv) > l1 ? 1.0f : 0.0f;
          const float sg = (wv > 0.0f) ? l1 : (wv < 0.0f ? -l1 : 0.0f);
          wv -= sg;
          uv -= sg;
          wv *= z;
        }
        if (prune && fabsf(wv) < FLT_MIN) wv = 0.0f;
      };
      upd(w[j].x, g[j].x, u[j].x); upd(w[j].y, g[j].y, u[j].y); upd(w[j].z, g[j].z, u[j].z); upd(w[j].w, g[j].w, u[j].w);
      __stcs(u4 + e, u[j]);
      w4[e] = w[j];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *dbg_slot(grp.flags[rank], bucket, 2) = gtime();
}

// Also closes the step for the replica group: the epoch (count_dev[2]) the tags are derived from advances here,
// once every shard of every bucket has landed.  The epoch is separate from the optimizer's step count
// (count_dev[0], which a checkpoint restore may set back): tags must never decrease.
__global__ void dp_wait_kernel(const b200_dp_group grp, int nbuckets, int64_t *count_dev) {
  const long long tag = (long long)count_dev[2] + 1;
  for (int i = threadIdx.x; i < nbuckets * grp.nranks; i += blockDim.x) {
    const long long *f = done_slot(grp.flags[grp.rank], i / grp.nranks, i % grp.nranks);
    while (ld_acquire_sys(f) < tag) { }
  }
  __syncthreads();
  if (threadIdx.x == 0) count_dev[2] = tag;
}

}  // namespace

extern "C" int b200_ipc_export(b200_ctx *ctx, void *dptr, void *handle64) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dptr && handle64, "NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, dptr));
  memcpy(handle64, &h, 64);
  return B200_OK;
}
extern "C" int b200_ipc_import(b200_ctx *ctx, const void *handle64, void **dptr) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && handle64 && dptr, "NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return B200_OK;
}
extern "C" int b200_ipc_close(b200_ctx *ctx, void *dptr) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx, "NULL pointer");
  if (dptr) CUDA_TRY(cudaIpcCloseMemHandle(dptr));
  return B200_OK;
}
extern "C" size_t b200_dp_flags_bytes(void) { return sizeof(long long) * (2 * DP_MAX_BUCKETS * B200_DP_MAX_RANKS + 4 * DP_MAX_BUCKETS + 8); }
extern "C" size_t b200_dp_debug_offset(void) { return sizeof(long long) * 2 * DP_MAX_BUCKETS * B200_DP_MAX_RANKS; }

extern "C" int b200_dp_fused_update(b200_ctx *ctx, const b200_dp_group *grp, int ntensors, const b200_sgd_tensor *tensors_dev,
                                    const b200_sgd_tensor *tensors_host, double decay, int64_t *count_dev, int bucket) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && grp && tensors_dev && tensors_host && count_dev, "NULL pointer");
  ARG_CHECK(grp->nranks >= 2 && grp->nranks <= B200_DP_MAX_RANKS && grp->rank >= 0 && grp->rank < grp->nranks, "bad replica group");
  ARG_CHECK(bucket >= 0 && bucket < DP_MAX_BUCKETS, "bucket out of range");
  if (ntensors <= 0) return B200_OK;
  ARG_CHECK(ntensors <= 32, "at most 32 tensors per bucket");
  for (int i = 0; i < ntensors; ++i)
    ARG_CHECK(tensors_host[i].max_norm_penalty <= 0.0f, "max_norm_penalty needs whole rows: use the all-reduce path");
  const int sms = ctx->sm_budget > 0 ? ctx->sm_budget : ctx->sm_count;
  static int unroll_env = -1, dbg = 0, use_dma = 0;   // bring-up switches
  if (unroll_env < 0) {
    const char *e = getenv("B200_DP_UNROLL");
    unroll_env = e ? atoi(e) : 0;
    const char *d = getenv("B200_DP_DEBUG");   // SM variant: bit 0 = no remote pushes, bit 1 = no remote weight stores
    dbg = d ? atoi(d) : 0;
    const char *m = getenv("B200_DP_DMA");     // 1: NVLink traffic on the copy engines (see below)
    use_dma = m ? atoi(m) : 0;
  }
  const int nr = grp->nranks, me = grp->rank;
  if (use_dma) {
    // the bucket as one contiguous range of the arenas (tensor slots are padded to 128 floats), cut into nr slices
    const size_t b_lo = (size_t)(tensors_host[0].w - grp->weights[me]);
    const b200_sgd_tensor &tl = tensors_host[ntensors - 1];
    const size_t b_hi = (size_t)(tl.w - grp->weights[me]) + (((size_t)tl.n + 127) & ~(size_t)127);
    for (int i = 1; i < ntensors; ++i)
      ARG_CHECK(tensors_host[i].w > tensors_host[i - 1].w, "the tensors of a bucket must be in arena order");
    const size_t granules = (b_hi - b_lo) / 128, per = (granules + nr - 1) / nr * 128;
    auto slice = [&](int q, size_t *lo, size_t *hi) {
      *lo = b_lo + (size_t)q * per < b_hi ? b_lo + (size_t)q * per : b_hi;
      *hi = *lo + per < b_hi ? *lo + per : b_hi;
    };
    const size_t slot = (size_t)grp->arena_elems;
    for (int q = 0; q < nr; ++q) {   // 1. push the gradients of every peer's slice
      if (q == me) continue;
      size_t lo, hi;
      slice(q, &lo, &hi);
      if (hi > lo)
        CUDA_TRY(cudaMemcpyAsync(grp->recv[q] + (size_t)(me < q ? me : me - 1) * slot + lo, grp->grads[me] + lo,
                                 (hi - lo) * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    dp_signal_kernel<<<1, 32, 0, ctx->stream>>>(*grp, bucket, 0, count_dev);   // 2. READY
    LAUNCH_CHECK(ctx);
    size_t lo, hi;
    slice(me, &lo, &hi);
    const size_t n4 = (hi - lo) / 4;
#define DP_LAUNCH_SHARD(NR, UN)                                                                                   \
  do {                                                                                                            \
    auto kern = dp_shard_update_kernel<NR, UN>;                                                                   \
    if (ONCE_PER_DEVICE(ctx)) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM)); \
    size_t blocks = (n4 + (size_t)DP_THREADS * UN - 1) / ((size_t)DP_THREADS * UN);                              \
    if (blocks > (size_t)sms) blocks = (size_t)sms;                                                               \
    if (blocks < 1) blocks = 1;                                                                                   \
    kern<<<(unsigned)blocks, DP_THREADS, DP_SMEM, ctx->stream>>>(*grp, tensors_dev, ntensors, decay, count_dev, bucket, \
                                                                 (unsigned long long)(lo / 4), (unsigned long long)(hi / 4)); \
  } while (0)
    if (nr <= 2) DP_LAUNCH_SHARD(2, 4);   // 3. the slice's update
    else if (nr <= 4) DP_LAUNCH_SHARD(4, 2);
    else DP_LAUNCH_SHARD(8, 2);
#undef DP_LAUNCH_SHARD
    LAUNCH_CHECK(ctx);
    for (int q = 0; q < nr; ++q) {   // 4. the updated slice into every replica
      if (q == me || hi <= lo) continue;
      CUDA_TRY(cudaMemcpyAsync(grp->weights[q] + lo, grp->weights[me] + lo, (hi - lo) * sizeof(float), cudaMemcpyDeviceToDevice,
                               ctx->stream));
    }
    dp_signal_kernel<<<1, 32, 0, ctx->stream>>>(*grp, bucket, 1, count_dev);   // 5. DONE
    LAUNCH_CHECK(ctx);
    return B200_OK;
  }
  size_t shard4 = 0;   // float4 of this rank's shards: enough CTAs to cover them once, at most the SMs planned for
  for (int i = 0; i < ntensors; ++i) shard4 += (((size_t)tensors_host[i].n + 3) / 4 + grp->nranks - 1) / grp->nranks;
  if (grp->mc_grads && grp->mc_weights) {
    // the arenas are bound to a multicast object: the NVSwitch reduces and broadcasts (dp_mc_update_kernel)
    constexpr int UN = 4;
    auto kern = dp_mc_update_kernel<UN>;
    if (ONCE_PER_DEVICE(ctx)) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM));
    size_t blocks = (shard4 + (size_t)DP_THREADS * UN - 1) / ((size_t)DP_THREADS * UN);
    if (blocks > (size_t)sms) blocks = (size_t)sms;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, DP_THREADS, DP_SMEM, ctx->stream>>>(*grp, tensors_dev, ntensors, decay, count_dev, bucket);
    LAUNCH_CHECK(ctx);
    return B200_OK;
  }
#define DP_LAUNCH(NR, UN)                                                                                         \
  do {                                                                                                            \
    auto kern = dp_fused_update_kernel<NR, UN>;                                                                   \
    if (ONCE_PER_DEVICE(ctx)) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM)); \
    size_t blocks = (shard4 + (size_t)DP_THREADS * UN - 1) / ((size_t)DP_THREADS * UN);                          \
    if (blocks > (size_t)sms) blocks = (size_t)sms;                                                               \
    if (blocks < 1) blocks = 1;                                                                                   \
    kern<<<(unsigned)blocks, DP_THREADS, DP_SMEM, ctx->stream>>>(*grp, tensors_dev, ntensors, decay, count_dev, bucket, dbg); \
  } while (0)
  if (grp->nranks <= 2) { if (unroll_env == 2) DP_LAUNCH(2, 2); else DP_LAUNCH(2, 4); }
  else if (grp->nranks <= 4) DP_LAUNCH(4, 2);
  else DP_LAUNCH(8, 2);
#undef DP_LAUNCH
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

extern "C" int b200_dp_wait(b200_ctx *ctx, const b200_dp_group *grp, int nbuckets, int64_t *count_dev) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && grp && count_dev, "NULL pointer");
  ARG_CHECK(nbuckets >= 0 && nbuckets <= DP_MAX_BUCKETS, "bucket count out of range");
  if (nbuckets == 0) return B200_OK;
  dp_wait_kernel<<<1, 128, 0, ctx->stream>>>(*grp, nbuckets, count_dev);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
