// Can a small kernel's CTAs become resident on SMs that a one-CTA-per-SM kernel with a large shared-memory
// footprint already occupies?  K1 spins ~100 us on every SM; K2 (tiny) is launched afterwards on another
// stream; we time K1-start -> K2-end.  ~few us => co-resident, ~100 us => K2 waited for K1.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_coresident tools/ubench_coresident.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int REGS, bool TMEM = false>
__global__ void __launch_bounds__(320, 1) k_spin(long long cycles, float *out) {
  extern __shared__ uint8_t smem[];
  __shared__ uint32_t slot;
  if (TMEM) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
  }
  float r[REGS];
#pragma unroll
  for (int i = 0; i < REGS; ++i) r[i] = threadIdx.x * 0.5f + i;
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
#pragma unroll
    for (int i = 0; i < REGS; ++i) r[i] = r[i] * 1.0001f + r[(i + 1) % REGS];
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < REGS; ++i) s += r[i];
  if (s == 12345.f) out[0] = s + smem[0];
  if (TMEM) {
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
  }
}
__global__ void __launch_bounds__(128) k_small(float *out) {
  if (threadIdx.x == 1000) out[0] = 1;
}
__global__ void __launch_bounds__(128) k_small_smem(float *out) {
  __shared__ float s[1024];
  s[threadIdx.x] = threadIdx.x;
  __syncthreads();
  if (s[(threadIdx.x + 1) & 127] == 12345.f) out[0] = 1;
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <typename K1, typename K2>
int run(const char *name, K1 k1, K2 k2) {
  cudaStream_t a, b;
  CK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, a));
    k1(a);
    CK(cudaEventRecord(e1, a));
    k2(b);
    CK(cudaEventRecord(e2, b));
    CK(cudaDeviceSynchronize());
  }
  float t1, t2;
  cudaEventElapsedTime(&t1, e0, e1);
  cudaEventElapsedTime(&t2, e0, e2);
  printf("%-70s K1 %7.1f us, K2 done after %7.1f us\n", name, t1 * 1e3, t2 * 1e3);
  return 0;
}

int main() {
  const long long cyc = 200000;
  CK(cudaFuncSetAttribute(k_spin<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CK(cudaFuncSetAttribute(k_spin<120>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k_spin<8>); printf("k_spin<8> regs %d\n", fa.numRegs);
  cudaFuncGetAttributes(&fa, k_spin<120>); printf("k_spin<120> regs %d\n", fa.numRegs);
  for (int smem_kb : {210}) {
    for (int heavy = 0; heavy < 2; ++heavy) {
      for (int carve = 0; carve < 2; ++carve) {
        cudaFuncSetAttribute(k_small, cudaFuncAttributePreferredSharedMemoryCarveout, carve ? cudaSharedmemCarveoutMaxShared : cudaSharedmemCarveoutDefault);
        char nm[128];
        snprintf(nm, sizeof nm, "K1 smem %3d KB, %s regs; K2 no smem, carve-out %s", smem_kb, heavy ? "~150" : "few", carve ? "max-shared" : "default");
        run(nm, [&](cudaStream_t s) { if (heavy) k_spin<120><<<148, 320, smem_kb * 1024, s>>>(cyc, nullptr); else k_spin<8><<<148, 320, smem_kb * 1024, s>>>(cyc, nullptr); },
            [&](cudaStream_t s) { k_small<<<592, 128, 0, s>>>(nullptr); });
      }
    }
  }
  CK(cudaFuncSetAttribute(k_spin<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  CK(cudaFuncSetAttribute((k_spin<8, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  CK(cudaFuncSetAttribute((k_spin<160, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  cudaFuncGetAttributes(&fa, k_spin<160>); printf("k_spin<160> regs %d\n", fa.numRegs);
  run("K1 smem 210 KB, k_spin<160>; K2 no smem", [&](cudaStream_t s) { k_spin<160><<<148, 320, 210 * 1024, s>>>(cyc, nullptr); },
      [&](cudaStream_t s) { k_small<<<592, 128, 0, s>>>(nullptr); });
  run("K1 smem 210 KB, few regs + TMEM 512 cols; K2 no smem", [&](cudaStream_t s) { k_spin<8, true><<<148, 320, 210 * 1024, s>>>(cyc, nullptr); },
      [&](cudaStream_t s) { k_small<<<592, 128, 0, s>>>(nullptr); });
  run("K1 smem 210 KB, k_spin<160> + TMEM 512 cols; K2 no smem", [&](cudaStream_t s) { k_spin<160, true><<<148, 320, 210 * 1024, s>>>(cyc, nullptr); },
      [&](cudaStream_t s) { k_small<<<592, 128, 0, s>>>(nullptr); });
  cudaFuncSetAttribute(k_small_smem, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  run("K1 smem 210 KB few regs; K2 4 KB static smem, max-shared", [&](cudaStream_t s) { k_spin<8><<<148, 320, 210 * 1024, s>>>(cyc, nullptr); },
      [&](cudaStream_t s) { k_small_smem<<<592, 128, 0, s>>>(nullptr); });
  return 0;
}
