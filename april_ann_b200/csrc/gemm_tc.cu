// TF32 contraction on the 5th-generation tensor cores (sm_100a): TMA -> 128B-swizzled shared
// memory -> tcgen05.mma (kind::tf32, fp32 accumulators in TMEM) -> tcgen05.ld -> fused epilogue
// (alpha, bias, activation, activation derivative, beta) -> global.  Replaces the cublasSgemm
// call of mathcore/c_src/gemm.cu:248-327 for the three contractions of a dense layer:
//   forward        Y  = X . W^T      A K-major,  B K-major
//   data gradient  dX = dY . W       A K-major,  B MN-major
//   weight grad.   dW = dY^T . X     A MN-major, B MN-major
//
// Persistent, warp-specialised CTA of 192 threads, one CTA per SM:
//   warp 0      TMA producer (one elected lane), 4-stage ring of {A 128x32, B BNx32} fp32 tiles
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer, 2 accumulator stages in TMEM
//   warps 2..5  epilogue: tcgen05.ld 32x32b.x32 (each warp owns the 32 TMEM lanes of its quadrant)
// K-major operands use SWIZZLE_128B (one TMA box per stage); MN-major fp32 operands must use the
// 128B-swizzle-with-32B-atoms layout (UMMA layout type 1, TMA SWIZZLE_128B_ATOM_32B), loaded as
// 32(MN) x 32(K) boxes of 4 KiB each.
#include <cuda.h>

#include <unordered_map>

#include "common.cuh"

namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 32;           // fp32 elements per 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32: 32 bytes of K per instruction
constexpr int STAGES = 4;
constexpr int MAX_BN = 256;
constexpr int A_BYTES = BM * BK * 4;          // 16 KiB
constexpr int B_BYTES = MAX_BN * BK * 4;      // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NTHREADS = 192;
constexpr int TMEM_COLS = 512;   // two 256-column accumulator stages

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}

struct TcParams {
  int M, N, K;
  int BN;            // N tile (multiple of 16, <= 256)
  int ldc;
  float *C;
  GemmEpilogue ep;
  // descriptor knobs (fixed in production; sweepable from the debug harness)
  uint32_t a_lbo, a_sbo, a_layout, a_kstep;
  uint32_t b_lbo, b_sbo, b_layout, b_kstep;
  uint32_t idesc;
};

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1 KiB alignment
  const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
  // barrier layout (8 bytes each): full[4] empty[4] tmem_full[2] tmem_empty[2] ; then the TMEM base slot
  auto full_bar = [&](int s) { return bars + 8 * s; };
  auto empty_bar = [&](int s) { return bars + 8 * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8 * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8 * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8 * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + p.BN - 1) / p.BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_k = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      // TMA always writes (and signals) whole boxes, also when they are partly out of bounds
      const uint32_t stage_tx = (uint32_t)A_BYTES + (B_KMAJOR ? (uint32_t)p.BN * BK * 4 : (uint32_t)((p.BN + 31) / 32) * 4096u);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * p.BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          mbar_expect_tx(full_bar(stage), stage_tx);
          const int k0 = kb * BK;
          if (A_KMAJOR) {
            tma_load_2d(sa, &map_a, full_bar(stage), k0, m0);                        // box {32 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) tma_load_2d(sa + j * 4096, &map_a, full_bar(stage), m0 + 32 * j, k0);
          }
          if (B_KMAJOR) {
            tma_load_2d(sb, &map_b, full_bar(stage), k0, n0);                        // box {32 k, BN rows}
          } else {
            for (int j = 0; j < (p.BN + 31) / 32; ++j) tma_load_2d(sb + j * 4096, &map_b, full_bar(stage), n0 + 32 * j, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (uint32_t)(local >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * MAX_BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = make_desc(sa + k * p.a_kstep, p.a_lbo, p.a_sbo, p.a_layout);
            const uint64_t bd = make_desc(sb + k * p.b_kstep, p.b_lbo, p.b_sbo, p.b_layout);
            umma_tf32(tmem_d, ad, bd, p.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));             // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));                 // accumulator complete -> epilogue
      }
    }
  } else {
    // ================================ epilogue warps ==============================
    const int quad = warp & 3;                        // TMEM lanes [32*quad, 32*quad+32)
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (uint32_t)(local >> 1) & 1;
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * p.BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int m = m0 + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * MAX_BN;
      float *crow = p.C + (size_t)m * p.ldc;
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        if (m < p.M) {
          const int nbase = n0 + c0;
          const int n_end = min(p.N, n0 + p.BN);     // BN need not be a multiple of the 32-column chunk
          const bool full_vec = (nbase + 32 <= n_end) && ((p.ldc & 3) == 0) && ((nbase & 3) == 0) &&
                                ((((uintptr_t)p.C) & 15) == 0) && p.ep.beta == 0.0f && p.ep.dact == B200_ACT_NONE;
          if (full_vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v;
              v.x = epilogue_apply(p.ep, __uint_as_float(r[j + 0]), m, nbase + j + 0, 0.0f);
              v.y = epilogue_apply(p.ep, __uint_as_float(r[j + 1]), m, nbase + j + 1, 0.0f);
              v.z = epilogue_apply(p.ep, __uint_as_float(r[j + 2]), m, nbase + j + 2, 0.0f);
              v.w = epilogue_apply(p.ep, __uint_as_float(r[j + 3]), m, nbase + j + 3, 0.0f);
              *reinterpret_cast<float4 *>(crow + nbase + j) = v;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = nbase + j;
              if (n < n_end) {
                const float cold = (p.ep.beta != 0.0f) ? crow[n] : 0.0f;
                crow[n] = epilogue_apply(p.ep, __uint_as_float(r[j]), m, n, cold);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapKey {
  const void *ptr;
  uint64_t d0, d1, ld;
  uint32_t b0, b1, sw;
  bool operator==(const MapKey &o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1 && sw == o.sw;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey &k) const {
    size_t h = (size_t)k.ptr;
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.d0); mix(k.d1); mix(k.ld); mix(k.b0); mix(k.b1); mix(k.sw);
    return h;
  }
};
struct TcState {
  EncodeTiledFn encode = nullptr;
  std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps;
  bool attr_set[4] = {false, false, false, false};
  // debug overrides (0 = production values)
  uint32_t dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool dbg_on = false;
  int force_bn = 0;
};

TcState *state(b200_ctx *ctx) {
  if (!ctx->tc_state) {
    TcState *s = new TcState();
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      s->encode = (EncodeTiledFn)fn;
    ctx->tc_state = s;
  }
  return (TcState *)ctx->tc_state;
}

// 2-D fp32 tensor map: dim0 = contiguous extent, dim1 = rows, row pitch ld floats
int get_map(b200_ctx *ctx, TcState *s, const float *ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1,
            CUtensorMapSwizzle sw, CUtensorMap *out) {
  MapKey key{ptr, d0, d1, ld, b0, b1, (uint32_t)sw};
  auto it = s->maps.find(key);
  if (it != s->maps.end()) { *out = it->second; return B200_OK; }
  cuuint64_t gdim[2] = {d0, d1};
  cuuint64_t gstride[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = s->encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200_set_error("cuTensorMapEncodeTiled failed (%d) for dims {%llu,%llu} ld %llu box {%u,%u}", (int)r,
                   (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, b0, b1);
    return B200_ERR_CUDA;
  }
  if (s->maps.size() > 4096) s->maps.clear();
  s->maps.emplace(key, m);
  *out = m;
  return B200_OK;
}

int pick_bn(int M, int N, int sm_count) {
  // N tile: multiple of 16 in [16, 256].  Prefer the widest tile that still yields >= ~0.85 wave
  // of the machine; otherwise the tile that maximises SM coverage.
  if (N <= 16) return 16;
  const int tiles_m = (M + BM - 1) / BM;
  int best = 16;
  double best_score = -1.0;
  for (int bn = 256; bn >= 32; bn -= 16) {
    const int tn = (N + bn - 1) / bn;
    const long tiles = (long)tiles_m * tn;
    const long waves = (tiles + sm_count - 1) / sm_count;
    // useful work fraction: real columns / padded columns, times wave occupancy; wide tiles are
    // cheaper per flop on shared-memory bandwidth (A is re-read per N tile)
    const double col_eff = (double)N / ((double)tn * bn);
    const double occ = (double)tiles / ((double)waves * sm_count);
    const double width = bn >= 192 ? 1.0 : (bn >= 128 ? 0.92 : (bn >= 64 ? 0.75 : 0.5));
    const double score = col_eff * occ * width;
    if (score > best_score + 1e-9) { best_score = score; best = bn; }
  }
  return best;
}

template <bool AK, bool BKM>
int launch(b200_ctx *ctx, TcState *s, int idx, const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &p, int grid) {
  auto kern = gemm_tc_kernel<AK, BKM>;
  if (!s->attr_set[idx]) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    s->attr_set[idx] = true;
  }
  kern<<<grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(ma, mb, p);
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

}  // namespace

void gemm_tc_destroy(b200_ctx *ctx) {
  if (ctx->tc_state) {
    delete (TcState *)ctx->tc_state;
    ctx->tc_state = nullptr;
  }
}

// debug hook: override descriptor fields {a_lbo,a_sbo,a_layout,a_kstep,b_lbo,b_sbo,b_layout,b_kstep}
extern "C" int b200_debug_tc_override(b200_ctx *ctx, int enable, const uint32_t *vals8, int force_bn) {
  TcState *s = state(ctx);
  s->dbg_on = enable != 0;
  if (vals8) memcpy(s->dbg, vals8, sizeof(s->dbg));
  s->force_bn = force_bn;
  return B200_OK;
}

int gemm_tc(b200_ctx *ctx, int transA, int transB, int M, int N, int K, const float *A, int lda, const float *B,
            int ldb, float *C, int ldc, const GemmEpilogue &ep) {
  if (M <= 0 || N <= 0) return B200_OK;
  TcState *s = state(ctx);
  if (!s->encode) return B200_ERR_UNSUPPORTED;
  // TMA: 16-byte aligned base and row pitch; tiny problems are not worth a persistent launch
  if ((((uintptr_t)A) & 15) || (((uintptr_t)B) & 15) || (lda & 3) || (ldb & 3) || K < 8) return B200_ERR_UNSUPPORTED;
  if ((long)M * N < 64 * 16) return B200_ERR_UNSUPPORTED;
  const bool a_k = (transA == 0);   // op(A)[m,k] = A[m*lda + k]
  const bool b_k = (transB != 0);   // op(B)[k,n] = B[n*ldb + k]

  TcParams p;
  p.M = M; p.N = N; p.K = K;
  p.BN = s->force_bn ? s->force_bn : pick_bn(M, N, ctx->sm_count);
  p.ldc = ldc;
  p.C = C;
  p.ep = ep;
  // K-major, SWIZZLE_128B: rows of 128 B, 8-row atoms 1 KiB apart (SBO), +32 B per UMMA_K step
  // MN-major fp32, 128B swizzle / 32 B atoms: 4-row k-groups 512 B apart (SBO), 32-wide MN chunks
  // one 4 KiB TMA box apart (LBO), two k-groups (1 KiB) per UMMA_K step
  p.a_layout = a_k ? 2u : 1u; p.a_lbo = a_k ? 16u : 4096u; p.a_sbo = a_k ? 1024u : 512u; p.a_kstep = a_k ? 32u : 1024u;
  p.b_layout = b_k ? 2u : 1u; p.b_lbo = b_k ? 16u : 4096u; p.b_sbo = b_k ? 1024u : 512u; p.b_kstep = b_k ? 32u : 1024u;
  if (s->dbg_on) {
    p.a_lbo = s->dbg[0]; p.a_sbo = s->dbg[1]; p.a_layout = s->dbg[2]; p.a_kstep = s->dbg[3];
    p.b_lbo = s->dbg[4]; p.b_sbo = s->dbg[5]; p.b_layout = s->dbg[6]; p.b_kstep = s->dbg[7];
  }
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A/B=tf32 [7,10)/[10,13)=2,
  // a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_k ? 0u : 1u) << 15) | ((b_k ? 0u : 1u) << 16) |
            ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  CUtensorMap ma, mb;
  int st;
  if (a_k) st = get_map(ctx, s, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B, &ma);
  else st = get_map(ctx, s, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, &ma);
  if (st) return st;
  if (b_k) st = get_map(ctx, s, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)p.BN, CU_TENSOR_MAP_SWIZZLE_128B, &mb);
  else st = get_map(ctx, s, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, &mb);
  if (st) return st;

  const int tiles = ((M + BM - 1) / BM) * ((N + p.BN - 1) / p.BN);
  const int grid = tiles < ctx->sm_count ? tiles : ctx->sm_count;
  if (a_k && b_k) return launch<true, true>(ctx, s, 0, ma, mb, p, grid);
  if (a_k && !b_k) return launch<true, false>(ctx, s, 1, ma, mb, p, grid);
  if (!a_k && b_k) return launch<false, true>(ctx, s, 2, ma, mb, p, grid);
  return launch<false, false>(ctx, s, 3, ma, mb, p, grid);
}
