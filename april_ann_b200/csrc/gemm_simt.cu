// fp32 FFMA contraction with the fused epilogue -- the parity ("strict") mode of
// b200_sgemm / b200_linear_* and the fallback for shapes the tcgen05 path rejects.
// Replaces the cublasSgemm call of mathcore/c_src/gemm.cu:248-327 (row-major, any
// transpose combination, arbitrary leading dimensions).
//
// 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread, register-staged
// double buffering.  A/B tiles are stored k-major in shared memory ([BK][BM]) so the
// inner product reads two float4 per operand per k.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int PAD = 4;

// loads one 128 x 16 operand tile into registers.  CONTIG_K: the k index is the
// contiguous one in global memory (op(A) with transA=0, op(B) with transB=1).
template <bool CONTIG_K>
__device__ __forceinline__ void load_tile(const float *__restrict__ base, int ld, int row0, int rows,
                                          int k0, int K, float (&r)[8]) {
  const int t = threadIdx.x;
  if (CONTIG_K) {
    // element (row, k) at base[row*ld + k]; thread -> k = t%16, row = t/16 + 16*i
    const int k = k0 + (t & 15);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row0 + (t >> 4) + 16 * i;
      r[i] = (row < rows && k < K) ? __ldg(base + (size_t)row * ld + k) : 0.0f;
    }
  } else {
    // element (row, k) at base[k*ld + row]; thread -> row = t%128, k = t/128 + 2*i
    const int row = row0 + (t & 127);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + (t >> 7) + 2 * i;
      r[i] = (row < rows && k < K) ? __ldg(base + (size_t)k * ld + row) : 0.0f;
    }
  }
}

template <bool CONTIG_K>
__device__ __forceinline__ void store_tile(float (*s)[BM + PAD], const float (&r)[8]) {
  const int t = threadIdx.x;
  if (CONTIG_K) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s[t & 15][(t >> 4) + 16 * i] = r[i];
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) s[(t >> 7) + 2 * i][t & 127] = r[i];
  }
}

template <bool A_CONTIG_K, bool B_CONTIG_K>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(int M, int N, int K, const float *__restrict__ A,
                                                       int lda, const float *__restrict__ B, int ldb,
                                                       float *__restrict__ C, int ldc, GemmEpilogue ep) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 thread grid

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float ra[8], rb[8];
  const int nk = (K + BK - 1) / BK;
  load_tile<A_CONTIG_K>(A, lda, m0, M, 0, K, ra);
  load_tile<B_CONTIG_K>(B, ldb, n0, N, 0, K, rb);
  store_tile<A_CONTIG_K>(As[0], ra);
  store_tile<B_CONTIG_K>(Bs[0], rb);
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_tile<A_CONTIG_K>(A, lda, m0, M, (kt + 1) * BK, K, ra);
      load_tile<B_CONTIG_K>(B, ldb, n0, N, (kt + 1) * BK, K, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      // rows ty*4..+3 and 64+ty*4..+3 ; cols tx*4..+3 and 64+tx*4..+3 (conflict-free float4 reads)
      float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile<A_CONTIG_K>(As[cur ^ 1], ra);
      store_tile<B_CONTIG_K>(Bs[cur ^ 1], rb);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float *p = C + (size_t)m * ldc + n;
      const float cold = (ep.beta != 0.0f) ? *p : 0.0f;
      *p = epilogue_apply(ep, acc[i][j], m, n, cold);
    }
  }
}

}  // namespace

int gemm_simt(b200_ctx *ctx, int transA, int transB, int M, int N, int K, const float *A, int lda,
              const float *B, int ldb, float *C, int ldc, const GemmEpilogue &ep) {
  if (M <= 0 || N <= 0) return B200_OK;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  // op(A)[m,k]: transA=0 -> A[m*lda+k] (k contiguous).  op(B)[k,n]: transB=1 -> B[n*ldb+k] (k contiguous).
  const bool a_ck = (transA == 0), b_ck = (transB != 0);
  if (a_ck && b_ck)
    gemm_simt_kernel<true, true><<<grid, NT, 0, ctx->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, ep);
  else if (a_ck && !b_ck)
    gemm_simt_kernel<true, false><<<grid, NT, 0, ctx->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, ep);
  else if (!a_ck && b_ck)
    gemm_simt_kernel<false, true><<<grid, NT, 0, ctx->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, ep);
  else
    gemm_simt_kernel<false, false><<<grid, NT, 0, ctx->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, ep);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}
