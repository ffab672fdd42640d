// Micro-benchmark (GPU box): how fast can ONE CTA per SM push a 64 KiB fp32 tile out of the SM?
//   mode 0: st.global.v4 (row-contiguous, 8 lanes per 128-byte row) from W warps
//   mode 1: cp.async.bulk.global.shared::cta (1-D bulk store) of 4 KiB pieces, W warps, depth D
//   mode 2: st.shared::cluster.v4 into the peer CTA of a 2-CTA cluster, W warps
//   mode 3: cp.async.bulk.shared::cluster.shared::cta (smem -> peer smem, mbarrier complete_tx), W warps
// Prints median cycles per CTA and bytes/cycle/SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int TILE_BYTES = 64 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) store_kernel(float *out, long long *cycles, int warps, int depth, int ld_floats) {
  extern __shared__ __align__(1024) uint8_t smem[];   // [0,64K) source tile, [64K,128K) receive area, then barrier
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 2 * TILE_BYTES;
  for (int i = threadIdx.x; i < TILE_BYTES / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = (float)i;
  uint32_t crank = 0;
  if (MODE >= 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  if (MODE == 3 && threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (MODE >= 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (MODE == 3 && threadIdx.x == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TILE_BYTES) : "memory");
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps) {
    const int pieces = TILE_BYTES / 4096;              // 16 pieces of 4 KiB (32 rows x 128 B)
    if (MODE == 0) {
      // tile = 128 rows x 128 floats; piece = 32 rows x 32 floats
      float *tile = out + (size_t)blockIdx.x * 128 * (size_t)ld_floats;
      for (int pc = warp; pc < pieces; pc += warps) {
        const int r0 = (pc >> 2) * 32, c0 = (pc & 3) * 32;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = r0 + it * 4 + (lane >> 3);
          const float4 v = *reinterpret_cast<const float4 *>(smem + pc * 4096 + (it * 4 + (lane >> 3)) * 128 + (lane & 7) * 16);
          *reinterpret_cast<float4 *>(tile + (size_t)row * ld_floats + c0 + (lane & 7) * 4) = v;
        }
      }
    } else if (MODE == 1) {
      float *tile = out + (size_t)blockIdx.x * (TILE_BYTES / 4);
      int issued = 0;
      for (int pc = warp; pc < pieces; pc += warps) {
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(tile + pc * 1024), "r"(sbase + pc * 4096) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (++issued >= depth) {
            if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
        }
        __syncwarp();
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (MODE == 2) {
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(sbase + TILE_BYTES), "r"(crank ^ 1u));
      for (int pc = warp; pc < pieces; pc += warps) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 v = *reinterpret_cast<const float4 *>(smem + pc * 4096 + lane * 128 + ((g ^ (lane & 7)) << 4));
          asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote + pc * 4096 + lane * 128 + ((g ^ (lane & 7)) << 4)),
                       "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
    } else {
      uint32_t remote, rbar;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(sbase + TILE_BYTES), "r"(crank ^ 1u));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(bar), "r"(crank ^ 1u));
      for (int pc = warp; pc < pieces; pc += warps) {
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], 4096, [%2];" ::"r"(remote + pc * 4096),
                       "r"(sbase + pc * 4096), "r"(rbar) : "memory");
        }
        __syncwarp();
      }
    }
  }
  if (MODE == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (MODE == 3) {
    // wait until the peer's 64 KiB have landed here
    asm volatile(
        "{\n.reg .pred P1;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE;\nbra WAIT_LOOP;\nDONE:\n}" ::"r"(bar)
        : "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps, int depth, float *out, long long *cyc_dev, int grid) {
  auto kern = store_kernel<MODE>;
  const size_t smem = 2 * TILE_BYTES + 64;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<long long> h(grid);
  double med = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = MODE >= 2 ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, out, cyc_dev, warps, depth, 2048));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), cyc_dev, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    med = (double)h[grid / 2];
  }
  printf("%-34s warps %2d depth %d : median %7.0f cycles  %6.1f B/cycle/SM (max CTA %lld)\n", name, warps, depth, med,
         TILE_BYTES / med, h[grid - 1]);
}

int main() {
  const int grid = 128;
  float *out;
  long long *cyc;
  CK(cudaMalloc(&out, (size_t)64 << 20));
  CK(cudaMalloc(&cyc, sizeof(long long) * 256));
  CK(cudaMemset(out, 0, (size_t)64 << 20));
  for (int w : {4, 8, 16}) run<0>("st.global.v4 (strided rows)", w, 0, out, cyc, grid);
  for (int w : {1, 4, 8, 16})
    for (int d : {1, 2}) run<1>("cp.async.bulk smem->global 4 KiB", w, d, out, cyc, grid);
  for (int w : {4, 8, 16}) run<2>("st.shared::cluster.v4 to peer", w, 0, out, cyc, grid);
  for (int w : {1, 4, 16}) run<3>("cp.async.bulk smem->peer smem 4 KiB", w, 0, out, cyc, grid);
  return 0;
}
