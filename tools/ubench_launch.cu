// Fixed cost of one kernel node in a replayed CUDA graph as a function of the launch configuration
// (dynamic shared memory size, cluster dimension, parameter block size, TMEM allocation).
// GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_launch tools/ubench_launch.cu && /tmp/ubench_launch
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

struct Big { uint64_t v[16]; };   // 128 bytes, like a CUtensorMap

template <bool TMEM>
__global__ void __launch_bounds__(192, 1) k_cfg(const __grid_constant__ Big a, const __grid_constant__ Big b,
                                               const __grid_constant__ Big c, int *out) {
  extern __shared__ uint8_t smem[];
  __shared__ uint32_t slot;
  if (TMEM) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
  }
  if (out && threadIdx.x == 0 && a.v[0] == 0x1234567) out[blockIdx.x] = smem[0] + (int)b.v[1] + (int)c.v[2];
}
__global__ void k_small(int *out) {
  if (out && threadIdx.x == 1000) out[0] = 1;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <typename F>
static int time_graph(const char *name, cudaStream_t s, F launch) {
  const int N = 50;
  cudaGraph_t g;
  cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  for (int i = 0; i < N; ++i) launch();
  CK(cudaStreamEndCapture(s, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventRecord(e0, s));
  for (int i = 0; i < 10; ++i) CK(cudaGraphLaunch(ge, s));
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-64s %7.2f us per kernel\n", name, ms * 1e3 / (10 * N));
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  return 0;
}

int main() {
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  Big z = {};
  CK(cudaFuncSetAttribute(k_cfg<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  CK(cudaFuncSetAttribute(k_cfg<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  time_graph("small kernel <<<1,32>>>", s, [&] { k_small<<<1, 32, 0, s>>>(nullptr); });
  time_graph("small kernel <<<1184,256>>>", s, [&] { k_small<<<1184, 256, 0, s>>>(nullptr); });
  const int smems[] = {0, 48 * 1024, 100 * 1024, 226 * 1024};
  for (int grid : {1, 128}) {
    for (int sm : smems) {
      char nm[128];
      snprintf(nm, sizeof nm, "cfg grid %3d, 192 thr, dyn smem %3d KB, 3x128B params", grid, sm / 1024);
      time_graph(nm, s, [&] { k_cfg<false><<<grid, 192, sm, s>>>(z, z, z, nullptr); });
    }
  }
  time_graph("cfg grid 128, 226 KB, + TMEM alloc/dealloc 512 cols", s, [&] { k_cfg<true><<<128, 192, 226 * 1024, s>>>(z, z, z, nullptr); });
  for (int tm = 0; tm < 2; ++tm) {
    char nm[128];
    snprintf(nm, sizeof nm, "cfg grid 128, 226 KB, cluster 2%s", tm ? " + TMEM" : "");
    time_graph(nm, s, [&] {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(128);
      cfg.blockDim = dim3(192);
      cfg.dynamicSmemBytes = 226 * 1024;
      cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int *np = nullptr;
      if (tm) cudaLaunchKernelEx(&cfg, k_cfg<true>, z, z, z, np);
      else cudaLaunchKernelEx(&cfg, k_cfg<false>, z, z, z, np);
    });
  }
  // alternating large-smem and small kernels (carve-out switches)
  time_graph("alternating: cfg 226 KB grid 128  /  small <<<1184,256>>> (per pair / 2)", s, [&] {
    k_cfg<false><<<128, 192, 226 * 1024, s>>>(z, z, z, nullptr);
    k_small<<<1184, 256, 0, s>>>(nullptr);
  });
  // programmatic dependent launch between identical 226 KB kernels
  time_graph("cfg grid 128, 226 KB + TMEM, programmatic stream serialization attr", s, [&] {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = 226 * 1024;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int *np = nullptr;
    cudaLaunchKernelEx(&cfg, k_cfg<true>, z, z, z, np);
  });
  return 0;
}
