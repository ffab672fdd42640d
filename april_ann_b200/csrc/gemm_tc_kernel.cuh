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
#include <math.h>

#pragma once
#include <set>
#include <unordered_map>

#include "common.cuh"

namespace b200tc {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 32;           // fp32 elements per 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32: 32 bytes of K per instruction
constexpr int MAX_STAGES = 8;
constexpr int MAX_BN = 256;
constexpr int A_BYTES = BM * BK * 4;          // 16 KiB
constexpr int EPI_WARPS = 8;         // two per TMEM lane quadrant, each takes half of the tile's column chunks
// Staging for the TMA stores.  On the last tile of a CTA (the only one at the BASELINE MLP sizes) the operand
// ring is idle once the accumulator is complete and every epilogue warp stages there; on earlier tiles of a
// persistent CTA the ring is busy with the next tile's loads, the epilogue is hidden behind that main loop,
// and one warp per quadrant works through a dedicated 4 KiB buffer.  Keeping the dedicated part small leaves
// ~16 KiB of shared memory per SM, so the light kernels of a training step (SGD, bias gradients, loss
// statistics) can be co-resident with a contraction CTA instead of waiting for it.
constexpr int STAGING_BYTES = 4 * 4096;
constexpr int BIAS_BYTES = MAX_BN * 4;               // the tile's slice of the bias vector
constexpr int SMEM_EXTRA = STAGING_BYTES + 256 /*barriers*/ + BIAS_BYTES + 1024 /*align slack*/;
constexpr int SMEM_MAX = 227 * 1024;
constexpr int SMEM_BUDGET = 211 * 1024;   // ring + extras: four 48 KiB stages of a 128x256 tile, ~16 KiB per SM left for co-resident kernels
constexpr int NTHREADS = 64 + 32 * EPI_WARPS;        // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int TMEM_COLS = 512;   // two 256-column accumulator stages
constexpr int STAMP_STRIDE = 32;  // bring-up: int64 stamps per CTA

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
// programmatic dependent launch: the next kernel of the stream may start its prologue (barriers, TMEM, tensor-map
// prefetch) while this one drains; nothing a previous kernel wrote is read before grid_dependency_wait()
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// C[box] += smem tile (fp32 add performed at the L2; commutative, so two contributors give a bit-exact sum)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
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

// the same load without the wait, and the wait as a separate step that names the registers (so that nothing
// reads them earlier): lets the load of the next chunk run under the processing of the current one
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// thread-block cluster helpers (split-K pairs)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
// arrive on an mbarrier of another CTA of the cluster (address from mapa), releasing this thread's
// earlier (remote) stores at cluster scope
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bulk copy shared::cta -> shared memory of another CTA of the cluster, completion counted in bytes on an
// mbarrier of the destination CTA
__device__ __forceinline__ void dsm_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP_C:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE_C;\n"
      "bra WAIT_LOOP_C;\n"
      "DONE_C:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}

struct TcParams {
  int M, N, K;
  int BN;            // N tile (multiple of 16, <= 256)
  int stages;        // smem ring depth (<= MAX_STAGES)
  uint32_t stage_bytes;   // A_BYTES + B tile bytes (multiple of 1 KiB)
  uint32_t staging_bytes; // dedicated store staging: 16 KiB (one 4 KiB tile per quadrant) or 32 KiB (one per epilogue warp)
  int ldc;
  float *C;
  GemmEpilogue ep;
  // descriptor knobs (fixed in production; sweepable from the debug harness)
  uint32_t a_lbo, a_sbo, a_layout, a_kstep;
  uint32_t b_lbo, b_sbo, b_layout, b_kstep;
  uint32_t idesc;
  int splitk;          // 1, or 2: a cluster of two CTAs shares one output tile, each contracting half of K;
                       //          the partial accumulators meet through distributed shared memory
  int reduce_add;      // 1 (lean kernels, beta == 1): every unit ADDS its tile into C with TMA reduce-add stores; with
                       //    splitk == 2 the k slices are independent units (no cluster, no exchange)
  int tma_store;       // 1: the epilogue writes C through map_c (cp.async.bulk.tensor store), 0: direct stores
  uint32_t dbg_flags;  // bring-up only: 8 = force direct stores instead of TMA stores, 16 = no split-K, 32 = epilogue phase clocks
  long long *stamps;   // bring-up only: per-CTA clock64 stamps [grid][16] (nullptr in production)
};

// v = alpha*acc (+ bias) ; v = ACT(v) ; v *= DACT'(dsrc) ; v += beta*cold  -- ACT / DACT fixed at compile time
template <int ACT, int DACT>
__device__ __forceinline__ float epi_value(float alpha, float beta, bool has_bias, bool has_c, float acc, float bias,
                                           float dsrc, float cold) {
  float v = alpha * acc;
  if (has_bias) v += bias;
  if (ACT != B200_ACT_NONE) v = act_apply(ACT, v);
  if (DACT != B200_ACT_NONE) v *= act_deriv_from_output(DACT, dsrc);
  if (has_c) v = fmaf(beta, cold, v);
  return v;
}

__device__ __forceinline__ void bulk_wait_read(int pending) {
  // cp.async.bulk.wait_group.read takes an immediate
  switch (pending) {
    case 0: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
    default: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory"); }

// staging buffers of one epilogue warp: `nbuf` 4 KiB tiles, `stride` bytes apart
struct Staging {
  uint32_t base, stride;
  int nbuf;
  __device__ __forceinline__ uint32_t buf(int j) const { return base + (uint32_t)j * stride; }
};

// Epilogue of the columns [c_begin, c_end) of one 128 x BN accumulator for one warp (32 TMEM lanes = 32 rows
// starting at m_base): tcgen05.ld 32 columns at a time -> shared memory (transpose) -> global through a TMA
// store, or direct row-contiguous stores when C cannot be described by a tensor map.
// Every field of the parameter block is copied into a register up front: left in the struct they
// were re-read from local memory for every element (ptxas kept a stack copy of the kernel parameters).
template <int ACT, int DACT>
__device__ __forceinline__ void epilogue_tile(const TcParams &p, const CUtensorMap *map_c, uint32_t taddr, const Staging &stg,
                                              int m_base, int n0, int lane, int &issued, int c_begin, int c_end,
                                              uint32_t recv, uint32_t bias_smem, const uint32_t (&relu_mask)[8], bool have_mask) {
  const float alpha = p.ep.alpha, beta = p.ep.beta;
  const float *__restrict__ dsrc = p.ep.dsrc;
  const int ld_dsrc = p.ep.ld_dsrc, ldc = p.ldc, M = p.M;
  float *__restrict__ C = p.C;
  const bool tma_store = p.tma_store != 0;
  const int n_end = min(p.N, n0 + p.BN);
  const int cg = lane & 7;                       // 16-byte column group this lane handles after the transpose
  const bool has_bias = p.ep.bias != nullptr, has_c = beta != 0.0f;
  constexpr bool has_d = DACT != B200_ACT_NONE;
  const bool c_vec = ((ldc & 3) == 0) && ((((uintptr_t)C) & 15) == 0);
  const bool d_vec = !has_d || (((ld_dsrc & 3) == 0) && ((((uintptr_t)dsrc) & 15) == 0));
  // simple epilogues (alpha, bias, activation) are applied in the accumulator layout (lane = row);
  // the ones that read global memory per element (derivative source, old C) run after the transpose
  // (ReLU derivatives arrive as per-chunk bit masks in the accumulator layout: simple as well)
  const bool masked = (DACT == B200_ACT_RELU) && have_mask && !has_bias;
  const bool simple = tma_store && !has_c && (!has_d || masked);
  // bring-up: phase durations of warp 2 lane 0, accumulated in registers, flushed once at the end
  const bool timing = p.stamps != nullptr && (p.dbg_flags & 32u) && threadIdx.x == 64;   // per-phase clocks perturb the epilogue: opt-in
  long long tacc[6] = {0, 0, 0, 0, 0, 0};
  long long tq = timing ? clock64() : 0;
#define EPI_STAMP(i) do { if (timing) { const long long tn = clock64(); tacc[i] += tn - tq; tq = tn; } } while (0)
#define EPI_MARK(slot) do { if (p.stamps != nullptr && threadIdx.x == 64) p.stamps[STAMP_STRIDE * blockIdx.x + (slot)] = clock64(); } while (0)
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    if (n0 + c0 >= n_end) break;
    if (p.stamps != nullptr && threadIdx.x == 64 && ((c0 - c_begin) >> 5) < 8)
      p.stamps[STAMP_STRIDE * blockIdx.x + 19 + ((c0 - c_begin) >> 5)] = clock64();   // bring-up: start of each chunk
    const uint32_t buf = stg.buf(issued % stg.nbuf);
    if (tma_store && issued >= stg.nbuf) {
      // the bulk store that last read this buffer (nbuf chunks ago) must have finished reading it
      if (lane == 0) bulk_wait_read(stg.nbuf - 1);
      __syncwarp();
    }
    EPI_STAMP(0);
    uint32_t r[32];
    tmem_ld32(taddr + c0, r);                    // lane = accumulator row, 32 consecutive columns
    EPI_STAMP(1);
    EPI_MARK(27);
    if (recv) {
      // split-K: add the peer CTA's partial sums of this chunk (same row-per-lane, XOR-swizzled layout)
      const uint32_t rb = recv + (uint32_t)((c0 - c_begin) >> 5) * 4096u + (uint32_t)lane * 128u;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float v[4];
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                     : "r"(rb + (uint32_t)((g ^ (lane & 7)) << 4)));
#pragma unroll
        for (int e = 0; e < 4; ++e) r[4 * g + e] = __float_as_uint(__uint_as_float(r[4 * g + e]) + v[e]);
      }
    }
    EPI_STAMP(2);
    if (simple && masked) {
      // chunk index within this warp's range selects the mask (unrolled select keeps the array in registers)
      const int q = (c0 - c_begin) >> 5;
      uint32_t mk = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) mk = (i == q) ? relu_mask[i] : mk;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        r[j] = ((mk >> j) & 1u) ? __float_as_uint(alpha * __uint_as_float(r[j])) : 0u;
    } else if (simple) {
      if (has_bias) {
        // the tile's bias slice was staged in shared memory before the accumulator was ready:
        // 8 broadcast float4 reads per chunk, no global latency on this path
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float b[4];
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b[0]), "=f"(b[1]), "=f"(b[2]), "=f"(b[3])
                       : "r"(bias_smem + (uint32_t)(c0 + 4 * g) * 4u));
#pragma unroll
          for (int e = 0; e < 4; ++e)
            r[4 * g + e] = __float_as_uint(epi_value<ACT, DACT>(alpha, beta, true, false, __uint_as_float(r[4 * g + e]), b[e], 1.0f, 0.0f));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = __float_as_uint(epi_value<ACT, DACT>(alpha, beta, false, false, __uint_as_float(r[j]), 0.0f, 1.0f, 0.0f));
      }
    }
    // 32x32 tile -> shared memory, 16-byte groups XOR-swizzled by the row (= TMA SWIZZLE_128B, and
    // conflict-free for the row-wise reads below)
#pragma unroll
    for (int g = 0; g < 8; ++g)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + (uint32_t)lane * 128u + (uint32_t)((g ^ (lane & 7)) << 4)),
                   "r"(r[4 * g]), "r"(r[4 * g + 1]), "r"(r[4 * g + 2]), "r"(r[4 * g + 3]) : "memory");
    EPI_STAMP(3);
    EPI_MARK(28);
    if (!simple) {
      __syncwarp();
      // row-contiguous domain: 8 lanes cover the 128 bytes of one row, global accesses coalesce
      const int n = n0 + c0 + 4 * cg;
      float bv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      if (has_bias)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bv[0]), "=f"(bv[1]), "=f"(bv[2]), "=f"(bv[3])
                     : "r"(bias_smem + (uint32_t)(c0 + 4 * cg) * 4u));
      const bool full4 = n + 4 <= n_end;
      // all global reads of the chunk are issued before the first use (8 independent 16-byte loads per lane)
      // (one source per epilogue kind: the derivative source for data gradients, old C for beta != 0)
      const bool pre = full4 && d_vec && c_vec && (has_d != has_c);
      float4 gq[8];
      if (pre) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int m = m_base + it * 4 + (lane >> 3);
          if (m < M)
            gq[it] = has_d ? __ldg(reinterpret_cast<const float4 *>(dsrc + (size_t)m * ld_dsrc + n))
                           : *reinterpret_cast<const float4 *>(C + (size_t)m * ldc + n);
        }
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3);
        const uint32_t saddr = buf + (uint32_t)row * 128u + (uint32_t)((cg ^ (row & 7)) << 4);
        float v[4];
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(saddr));
        const int m = m_base + row;
        if (m >= M || n >= n_end) continue;
        float *cp = C + (size_t)m * ldc + n;
        const float *dp = dsrc + (size_t)m * ld_dsrc + n;
        float d[4] = {1.0f, 1.0f, 1.0f, 1.0f}, c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (pre) {
          if (has_d) { d[0] = gq[it].x; d[1] = gq[it].y; d[2] = gq[it].z; d[3] = gq[it].w; }
          else { c[0] = gq[it].x; c[1] = gq[it].y; c[2] = gq[it].z; c[3] = gq[it].w; }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (n + e < n_end) {
              if (has_d) d[e] = __ldg(dp + e);
              if (has_c) c[e] = cp[e];
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = epi_value<ACT, DACT>(alpha, beta, has_bias, has_c, v[e], bv[e], d[e], c[e]);
        if (tma_store) {
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
        } else if (full4 && c_vec) {
          *reinterpret_cast<float4 *>(cp) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (n + e < n_end) cp[e] = v[e];
        }
      }
    }
    EPI_STAMP(4);
    EPI_MARK(29);
    if (tma_store) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA
      __syncwarp();
      EPI_MARK(30);
      if (lane == 0) tma_store_2d(map_c, buf, n0 + c0, m_base);      // rows/columns past M/N are clipped
      ++issued;
      EPI_MARK(31);
    } else {
      __syncwarp();
    }
    EPI_STAMP(5);
  }
#undef EPI_STAMP
#undef EPI_MARK
  if (timing)
    for (int i = 0; i < 6; ++i) p.stamps[STAMP_STRIDE * blockIdx.x + 8 + i] += tacc[i];
}


// Lean epilogue of the columns [c_begin, c_end) of one 128 x BN accumulator for one warp: the only cases a
// training step of the BASELINE networks produces -- alpha (+ bias) (+ activation), or alpha x ReLU gate from
// the per-chunk bit masks -- applied in the accumulator layout and written with TMA stores.  No global loads,
// no transposed pass, no run-time epilogue switches: the instruction stream of a chunk is ~250 instructions
// (the generic epilogue_tile above is ~3500 per activation kind, and the eight epilogue warps of a CTA were
// instruction-fetch bound in it: 13-30 % of the stall samples of the round-1 ncu captures were "no instruction").
template <int ACT, int DACT>
__device__ __forceinline__ void epilogue_lean(const TcParams &p, const CUtensorMap *map_c, uint32_t taddr, const Staging &stg,
                                              int m_base, int n0, int lane, int &issued, int c_begin, int c_end,
                                              uint32_t recv, uint32_t bias_smem, const uint32_t (&relu_mask)[8]) {
  const float alpha = p.ep.alpha;
  const bool has_bias = p.ep.bias != nullptr;
  const bool reduce_add = p.reduce_add != 0;
  const int n_end = min(p.N, n0 + p.BN);
  c_end = min(c_end, ((n_end - n0 + 31) >> 5) << 5);     // chunks that hold at least one valid column
  if (c_begin >= c_end) return;
  // everything after the TMEM load of one chunk: (+ peer's partial sums) -> epilogue math -> swizzled shared-memory
  // tile -> TMA store
  auto finish = [&](uint32_t (&r)[32], int c0) {
    const uint32_t buf = stg.buf(issued % stg.nbuf);
    if (issued >= stg.nbuf) {
      if (lane == 0) bulk_wait_read(stg.nbuf - 1);   // the store that last read this buffer has finished reading it
      __syncwarp();
    }
    if (recv) {
      const uint32_t rb = recv + (uint32_t)((c0 - c_begin) >> 5) * 4096u + (uint32_t)lane * 128u;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float v[4];
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                     : "r"(rb + (uint32_t)((g ^ (lane & 7)) << 4)));
#pragma unroll
        for (int e = 0; e < 4; ++e) r[4 * g + e] = __float_as_uint(__uint_as_float(r[4 * g + e]) + v[e]);
      }
    }
    if (DACT == B200_ACT_RELU) {
      const int q = (c0 - c_begin) >> 5;
      uint32_t mk = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) mk = (i == q) ? relu_mask[i] : mk;
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = ((mk >> j) & 1u) ? __float_as_uint(alpha * __uint_as_float(r[j])) : 0u;
    } else if (has_bias) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float b[4];
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b[0]), "=f"(b[1]), "=f"(b[2]), "=f"(b[3])
                     : "r"(bias_smem + (uint32_t)(c0 + 4 * g) * 4u));
#pragma unroll
        for (int e = 0; e < 4; ++e)
          r[4 * g + e] = __float_as_uint(epi_value<ACT, B200_ACT_NONE>(alpha, 0.0f, true, false, __uint_as_float(r[4 * g + e]), b[e], 1.0f, 0.0f));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        r[j] = __float_as_uint(epi_value<ACT, B200_ACT_NONE>(alpha, 0.0f, false, false, __uint_as_float(r[j]), 0.0f, 1.0f, 0.0f));
    }
#pragma unroll
    for (int g = 0; g < 8; ++g)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + (uint32_t)lane * 128u + (uint32_t)((g ^ (lane & 7)) << 4)),
                   "r"(r[4 * g]), "r"(r[4 * g + 1]), "r"(r[4 * g + 2]), "r"(r[4 * g + 3]) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      if (reduce_add) tma_reduce_add_2d(map_c, buf, n0 + c0, m_base);
      else tma_store_2d(map_c, buf, n0 + c0, m_base);
    }
    ++issued;
  };
  // two register sets: the TMEM load of chunk i+1 is in flight while chunk i is finished (the TMEM read port, 64 B
  // per cycle for the whole SM, is the floor of the epilogue; eight warps that load, compute and store in
  // lockstep left it idle two thirds of the time)
  uint32_t ra[32], rb2[32];
  tmem_ld32_issue(taddr + c_begin, ra);
  for (int c0 = c_begin; c0 < c_end; c0 += 64) {
    tmem_ld32_wait(ra);
    if (c0 + 32 < c_end) tmem_ld32_issue(taddr + c0 + 32, rb2);
    finish(ra, c0);
    if (c0 + 32 < c_end) {
      tmem_ld32_wait(rb2);
      if (c0 + 64 < c_end) tmem_ld32_issue(taddr + c0 + 64, ra);
      finish(rb2, c0 + 32);
    }
  }
}

// Kernel variants.  LEAN: the epilogue is fixed at compile time to <V_ACT, V_DACT> in its "simple" form (TMA
// stores, beta == 0, derivative only as ReLU gate masks); the host picks a lean variant whenever the call
// qualifies.  !LEAN: the generic kernel, every epilogue kind behind run-time switches (beta != 0, tanh / logistic
// derivatives, unaligned C).  STAMPS: bring-up clocks (tools/gemm_stamps.py), compiled out otherwise.
// 128 registers per thread: the 10 warps land 3/3/2/2 on the four SM sub-partitions, and a sub-partition
// with three contraction warps must still have room for a warp of a light kernel (SGD, bias gradient)
// that runs beside it; at the 168 registers a 320-thread launch bound allows nothing else became resident
// (tools/ubench_coresident.cu)
template <bool A_KMAJOR, bool B_KMAJOR, int V_ACT, int V_DACT, bool LEAN, bool STAMPS>
__global__ void __maxnreg__(128)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1 KiB alignment
  const int STAGES = p.stages;
  const uint32_t STAGE_BYTES = p.stage_bytes;
  const uint32_t staging = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bars = staging + p.staging_bytes;
  // barrier layout (8 bytes each): full[8] empty[8] tmem_full[2] tmem_empty[2] ; then the TMEM base slot
  auto full_bar = [&](int s) { return bars + 8 * s; };
  auto empty_bar = [&](int s) { return bars + 8 * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8 * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8 * (2 * MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8 * (2 * MAX_STAGES + 4);
  // split-K exchange: xready = "the peer's ring is idle, send", xdone = "the peer's partial sums have landed here"
  // xack = "the peer has everything it needs from this CTA's shared memory" (the source of the bulk copies)
  const uint32_t xready_bar = bars + 8 * (2 * MAX_STAGES + 5), xdone_bar = bars + 8 * (2 * MAX_STAGES + 6);
  const uint32_t xack_bar = bars + 8 * (2 * MAX_STAGES + 7);
  const uint32_t bias_smem = bars + 256u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long *stamp = (STAMPS && p.stamps) ? p.stamps + STAMP_STRIDE * blockIdx.x : nullptr;
  if (stamp && threadIdx.x == 0) {
    stamp[0] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    stamp[14] = (long long)gt;
  }
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + p.BN - 1) / p.BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_k = (p.K + BK - 1) / BK;
  // Work units.  A unit is one output tile contracted over one of `splitk` slices of K.
  //  * splitk == 1: unit = tile.
  //  * exchange mode (splitk == 2, !reduce_add): a cluster of two CTAs shares a tile, CTA rank r contracts slice r,
  //    the partial sums meet through distributed shared memory and each CTA finishes half of the tile.
  //  * reduce-add mode (reduce_add): C already holds what the result is added to (beta == 1), every unit adds its
  //    partial tile with TMA reduce-add stores; units are independent and the persistent CTAs simply stride over them.
  const bool xchg = (p.splitk == 2) && !p.reduce_add;
  const uint32_t crank = xchg ? cluster_ctarank() : 0u;
  const int S = p.splitk;
  const int kb_per = (num_k + S - 1) / S;
  const int unit_first = xchg ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = xchg ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_units = xchg ? num_tiles : num_tiles * S;
  auto unit_tile = [&](int u) { return xchg ? u : u / S; };
  auto unit_kb_lo = [&](int u) { return (xchg ? (int)crank : u % S) * kb_per; };
  auto unit_kb_hi = [&](int u) { return min(num_k, unit_kb_lo(u) + kb_per); };
  // TMA always writes (and signals) whole boxes, also when they are partly out of bounds
  const uint32_t stage_tx = (uint32_t)A_BYTES + (B_KMAJOR ? (uint32_t)p.BN * BK * 4 : (uint32_t)((p.BN + 31) / 32) * 4096u);

  // one TMA per operand per stage.  MN-major operands are described as 3-D tensors
  // {32 contiguous elements, K rows, MN/32 chunks}: a box {32, 32, tile/32} lands as consecutive
  // 4 KiB [32 k][32 mn] blocks, the layout the MN-major UMMA descriptor walks (LBO = 4 KiB).
  // (issuing the 4 + 8 separate 4 KiB boxes of a 128x256 tile cost ~100 cycles each and made the
  // weight-gradient main loop TMA-issue bound: 1150 instead of 610 cycles per k-block)
  auto load_stage = [&](int stage, int m0, int n0, int kb) {
    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
    mbar_expect_tx(full_bar(stage), stage_tx);
    const int k0 = kb * BK;
    if (A_KMAJOR) tma_load_2d(sa, &map_a, full_bar(stage), k0, m0);            // box {32 k, 128 rows}
    else tma_load_3d(sa, &map_a, full_bar(stage), 0, k0, m0 >> 5);             // box {32 m, 32 k, 4}
    if (B_KMAJOR) tma_load_2d(sb, &map_b, full_bar(stage), k0, n0);            // box {32 k, BN rows}
    else tma_load_3d(sb, &map_b, full_bar(stage), 0, k0, n0 >> 5);             // box {32 n, 32 k, BN/32}
  };

  // the first pass over the ring needs no empty-slot wait: thread 0 initialises the barriers and starts
  // the first loads right away, so that their latency overlaps the TMEM allocation and the CTA-wide sync
  int pre_issued = 0;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 32 * EPI_WARPS); }
    mbar_init(xready_bar, 1);
    mbar_init(xdone_bar, 1);
    mbar_init(xack_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (xchg) {
      // the peer ships 4 quadrants x (chunks this CTA owns) x 4 KiB of partial sums
      const int nch = (p.BN + 31) >> 5, own0 = (nch + 1) >> 1;
      mbar_expect_tx(xdone_bar, (uint32_t)(4 * (crank ? nch - own0 : own0)) * 4096u);
    }
    // everything above overlapped the tail of the previous kernel of the stream (programmatic dependent
    // launch); its results -- this kernel's operands -- are only touched from here on
    grid_dependency_wait();
    if (unit_first < num_units) {
      const int tile = unit_tile(unit_first), kb_lo = unit_kb_lo(unit_first), kb_hi = unit_kb_hi(unit_first);
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * p.BN;
      const int npre = min(STAGES, kb_hi - kb_lo);
      for (; pre_issued < npre; ++pre_issued) load_stage(pre_issued, m0, n0, kb_lo + pre_issued);
    }
    if (stamp) stamp[2] = clock64();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  grid_dependency_wait();   // every thread: the epilogue warps read the bias, the derivative source and C
  // split-K pair: nobody touches the peer's barriers before the peer has initialised them.  The arrive is
  // here, the matching wait sits right before the first remote access (epilogue warps) or at the end of
  // the role (producer / MMA warps), so the handshake costs nothing.
  if (xchg) cluster_arrive_relaxed();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (stamp && threadIdx.x == 0) stamp[1] = clock64();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int skip = pre_issued;   // k-blocks of the first tile already in flight
      for (int u = unit_first; u < num_units; u += unit_step) {
        const int tile = unit_tile(u), kb_lo = unit_kb_lo(u), kb_hi = unit_kb_hi(u);
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * p.BN;
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          if (skip > 0) {
            --skip;
          } else {
            mbar_wait(empty_bar(stage), phase ^ 1);
            load_stage(stage, m0, n0, kb);
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
      for (int u = unit_first; u < num_units; u += unit_step, ++local) {
        const int kb_lo = unit_kb_lo(u), kb_hi = unit_kb_hi(u);
        const int acc = local & 1;
        const uint32_t acc_phase = (uint32_t)(local >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * MAX_BN;
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (stamp && kb == kb_lo && local == 0) stamp[3] = clock64();
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = make_desc(sa + k * p.a_kstep, p.a_lbo, p.a_sbo, p.a_layout);
            const uint64_t bd = make_desc(sb + k * p.b_kstep, p.b_lbo, p.b_sbo, p.b_layout);
            umma_tf32(tmem_d, ad, bd, p.idesc, (kb != kb_lo || k != 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));             // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));                 // accumulator complete -> epilogue
        if (stamp) stamp[4] = clock64();
      }
      // this CTA's contraction work is issued: once every CTA is here (or gone) the next kernel of the stream
      // may be launched, so that its launch latency and prologue run under this kernel's epilogues
      grid_launch_dependents();
    }
  } else {
    // ================================ epilogue warps ==============================
    const int quad = warp & 3;                        // TMEM lanes [32*quad, 32*quad+32)
    const int half = (warp - 2) >> 2;                 // which half of the tile's column chunks
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;                  // 0 .. 32*EPI_WARPS-1
    int local = 0;
    int issued = 0;
    for (int u = unit_first; u < num_units; u += unit_step, ++local) {
      const int tile = unit_tile(u);
      const bool last_unit = u + unit_step >= num_units;
      const int acc = local & 1;
      const uint32_t acc_phase = (uint32_t)(local >> 1) & 1;
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * p.BN;
      // the tile's bias slice -> shared memory while the contraction is still running
      if (p.ep.bias != nullptr) {
        epi_bar_sync();                               // every warp is done with the previous tile's slice
        for (int j = et; j < p.BN; j += 32 * EPI_WARPS) {
          const float bvv = (n0 + j < p.N) ? __ldg(p.ep.bias + n0 + j) : 0.0f;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_smem + (uint32_t)j * 4u), "f"(bvv) : "memory");
        }
        epi_bar_sync();
      }
      // ReLU data gradients: the derivative is one bit per element.  While the contraction is still running
      // this warp reads the activation tile of the layer below for the chunks it will finish (lane = its
      // accumulator row, 8 x 16 bytes per chunk) and keeps one 32-bit mask per chunk, so that the epilogue
      // needs no global load and stays in the accumulator layout.  (Loading those values inside the
      // epilogue, after the transpose, cost ~5500 cycles per 32x32 chunk of exposed latency.)
      uint32_t relu_mask[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      bool have_mask = false;
      if ((LEAN ? V_DACT == B200_ACT_RELU : (p.ep.dact == B200_ACT_RELU && p.tma_store && p.ep.beta == 0.0f)) && p.BN <= 256) {
        have_mask = true;
        const int nch_p = (p.BN + 31) >> 5;
        int lo = 0, hi = nch_p;
        if (xchg) {
          const int own0 = (nch_p + 1) >> 1;
          lo = crank ? own0 : 0;
          hi = crank ? nch_p : own0;
        }
        if (last_unit || p.staging_bytes >= (uint32_t)(EPI_WARPS * 4096)) {
          const int mid = lo + ((hi - lo + 1) >> 1);
          if (half) lo = mid; else hi = mid;
        } else if (half) {
          lo = hi;
        }
        const int row = m0 + quad * 32 + lane;
        const float *drow = p.ep.dsrc + (size_t)min(row, p.M - 1) * p.ep.ld_dsrc + n0;
        const bool vec = ((p.ep.ld_dsrc & 3) == 0) && ((((uintptr_t)p.ep.dsrc) & 15) == 0) && ((n0 & 3) == 0);
        const int n_lim = min(p.N - n0, p.BN);          // valid columns of this tile
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int ch = lo + q;
          if (ch < hi) {
            uint32_t mk = 0u;
            if (vec && ch * 32 + 32 <= n_lim) {
              float4 v[8];
#pragma unroll
              for (int g = 0; g < 8; ++g) v[g] = __ldg(reinterpret_cast<const float4 *>(drow + ch * 32 + 4 * g));
#pragma unroll
              for (int g = 0; g < 8; ++g)
                mk |= ((v[g].x > 0.0f ? 1u : 0u) | (v[g].y > 0.0f ? 2u : 0u) | (v[g].z > 0.0f ? 4u : 0u) |
                       (v[g].w > 0.0f ? 8u : 0u)) << (4 * g);
            } else {
              for (int j = 0; j < 32; ++j)
                if (ch * 32 + j < n_lim && __ldg(drow + ch * 32 + j) > 0.0f) mk |= 1u << j;
            }
            relu_mask[q] = mk;
          }
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (stamp && threadIdx.x == 64) stamp[5] = clock64();
      // column chunks (32 wide) of the tile this warp finishes: [c_begin, c_end)
      const int nch = (p.BN + 31) >> 5;
      int ch_lo = 0, ch_hi = nch;
      uint32_t recv = 0;
      uint32_t ring_free = smem_base;                 // ring bytes from here on are free for staging (if any)
      if (xchg) {
        // Exchange: this CTA finishes the chunks [ch_lo, ch_hi) of the tile and ships its partial sums of
        // the others to the peer.  The receive area is the (now idle) operand ring of the peer,
        // which is only safe to overwrite once the peer's MMAs have retired: cluster barrier #1.
        const int own0 = (nch + 1) >> 1;                          // chunks owned by rank 0
        ch_lo = crank ? own0 : 0;
        ch_hi = crank ? nch : own0;
        const int s_lo = crank ? 0 : own0, s_hi = crank ? own0 : nch;
        const int peer_chunks = s_hi - s_lo, own_chunks = ch_hi - ch_lo;
        __syncwarp();
        cluster_wait();                                // the peer's barriers exist (arrive: after the setup sync)
        // this CTA's accumulator is complete, so its ring is idle: tell the peer it may send
        if (ew == 0 && lane == 0) mbar_arrive_remote(mapa_shared(xready_bar, crank ^ 1u));
        // Stage this warp's share of the peer's chunks in the local ring (behind the receive area), in the
        // row-per-lane XOR-swizzled layout the receiver reads; then one bulk copy per 4 KiB chunk moves it
        // into the peer's receive area.  (Per-lane st.shared::cluster of 16-byte pieces moved < 10 B/cycle.)
        const uint32_t taddr_s = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * MAX_BN;
        const int max_send = (peer_chunks + 1) >> 1;
        const int my_lo = s_lo + (half ? max_send : 0), my_hi = half ? s_hi : s_lo + max_send;
        const uint32_t send_base = smem_base + (uint32_t)(4 * own_chunks + ew * max_send) * 4096u;
        int ns = 0;
        for (int ch = my_lo; ch < my_hi; ++ch, ++ns) {
          uint32_t r[32];
          tmem_ld32(taddr_s + ch * 32, r);
          const uint32_t dst = send_base + (uint32_t)ns * 4096u + (uint32_t)lane * 128u;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)((g ^ (lane & 7)) << 4)),
                         "r"(r[4 * g]), "r"(r[4 * g + 1]), "r"(r[4 * g + 2]), "r"(r[4 * g + 3]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        mbar_wait_cluster(xready_bar, 0);              // the peer's ring is idle
        if (stamp && threadIdx.x == 64) stamp[16] = clock64();
        if (lane == 0) {
          const uint32_t remote = mapa_shared(smem_base + (uint32_t)(quad * peer_chunks) * 4096u, crank ^ 1u);
          const uint32_t remote_bar = mapa_shared(xdone_bar, crank ^ 1u);
          for (int i = 0; i < ns; ++i)
            dsm_bulk_copy(remote + (uint32_t)(my_lo + i - s_lo) * 4096u, send_base + (uint32_t)i * 4096u, 4096u, remote_bar);
        }
        if (stamp && threadIdx.x == 64) stamp[17] = clock64();
        mbar_wait_cluster(xdone_bar, 0);               // the peer's partial sums of this CTA's chunks have landed
        if (stamp && threadIdx.x == 64) stamp[18] = clock64();
        if (ew == 0 && lane == 0) mbar_arrive_remote(mapa_shared(xack_bar, crank ^ 1u));
        recv = smem_base + (uint32_t)(quad * own_chunks) * 4096u;
        ring_free = smem_base + (uint32_t)(4 * own_chunks + EPI_WARPS * max_send) * 4096u;
      }
      Staging stg;
      int w_lo, w_hi;
      if (last_unit) {
        // last tile of this CTA: the producer has nothing more to load, the ring behind ring_free is idle.
        // The two warps of a quadrant split the chunks; each warp gets up to 4 staging tiles of its own.
        const int ch_mid = ch_lo + ((ch_hi - ch_lo + 1) >> 1);
        w_lo = half ? ch_mid : ch_lo;
        w_hi = half ? ch_hi : ch_mid;
        const uint32_t ring_end = smem_base + (uint32_t)STAGES * STAGE_BYTES;
        const int avail = (int)((ring_end - ring_free) / (uint32_t)(EPI_WARPS * 4096));
        stg.base = ring_free + (uint32_t)ew * 4096u;
        stg.stride = (uint32_t)(EPI_WARPS * 4096);
        stg.nbuf = min(4, avail);
        if (avail < 1) {   // cannot happen with the rings gemm_tc() sizes; stay correct anyway
          have_mask = false;   // the masks were gathered for the split ranges
          w_lo = half ? ch_hi : ch_lo;
          w_hi = ch_hi;
          stg.base = staging + (uint32_t)quad * 4096u;
          stg.nbuf = 1;
        }
      } else if (p.staging_bytes >= (uint32_t)(EPI_WARPS * 4096)) {
        // the ring is busy with the next tile; launches with more tiles than CTAs carry a dedicated 4 KiB tile
        // per epilogue warp, so all eight warps work (short-K shapes are epilogue bound otherwise)
        const int ch_mid = ch_lo + ((ch_hi - ch_lo + 1) >> 1);
        w_lo = half ? ch_mid : ch_lo;
        w_hi = half ? ch_hi : ch_mid;
        stg.base = staging + (uint32_t)ew * 4096u;
        stg.stride = 0;
        stg.nbuf = 1;
      } else {
        // the ring is busy with the next tile: one warp per quadrant, dedicated buffer
        w_lo = half ? ch_hi : ch_lo;
        w_hi = ch_hi;
        stg.base = staging + (uint32_t)quad * 4096u;
        stg.stride = 0;
        stg.nbuf = 1;
      }
      if (recv) recv += (uint32_t)(w_lo - ch_lo) * 4096u;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * MAX_BN;
      const int c_begin = w_lo * 32, c_end = min(w_hi * 32, p.BN);
      // the activation / derivative kind is a compile-time constant inside each instantiation: a run-time
      // switch per element made the unrolled epilogue ~4500 instructions per 32-column chunk
#define EPI_CALL(A, D) epilogue_tile<A, D>(p, &map_c, taddr, stg, m0 + quad * 32, n0, lane, issued, c_begin, c_end, recv, bias_smem, relu_mask, have_mask)
      if (LEAN) {
        epilogue_lean<V_ACT, V_DACT>(p, &map_c, taddr, stg, m0 + quad * 32, n0, lane, issued, c_begin, c_end, recv, bias_smem, relu_mask);
      } else if (p.ep.dact != B200_ACT_NONE) {
        switch (p.ep.dact) {
          case B200_ACT_LOGISTIC: EPI_CALL(B200_ACT_NONE, B200_ACT_LOGISTIC); break;
          case B200_ACT_TANH: EPI_CALL(B200_ACT_NONE, B200_ACT_TANH); break;
          default: EPI_CALL(B200_ACT_NONE, B200_ACT_RELU); break;
        }
      } else {
        switch (p.ep.act) {
          case B200_ACT_LOGISTIC: EPI_CALL(B200_ACT_LOGISTIC, B200_ACT_NONE); break;
          case B200_ACT_TANH: EPI_CALL(B200_ACT_TANH, B200_ACT_NONE); break;
          case B200_ACT_RELU: EPI_CALL(B200_ACT_RELU, B200_ACT_NONE); break;
          default: EPI_CALL(B200_ACT_NONE, B200_ACT_NONE); break;
        }
      }
#undef EPI_CALL
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (stamp && threadIdx.x == 64) stamp[6] = clock64();
    }
    if (p.tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    // the bulk copies of the exchange read this CTA's shared memory: stay until the peer has acknowledged
    if (xchg && ew == 0) mbar_wait_cluster(xack_bar, 0);
  }
  if (warp < 2 && xchg) {
    __syncwarp();
    cluster_wait();   // pairs with the arrive after the setup sync
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  if (stamp && threadIdx.x == 0) {
    stamp[7] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    stamp[15] = (long long)gt;
  }
}


// ------------------------------------------------------------------ host side shared by the instantiation units
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TcState {
  EncodeTiledFn encode = nullptr;
  std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps;
  std::set<const void *> attr_set;     // kernels whose shared-memory limit has been raised on this context's device
  bool pdl = true;                     // programmatic dependent launch (B200_PDL=0 switches it off)
  // debug overrides (0 = production values)
  uint32_t dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool dbg_on = false;
  int force_bn = 0;
  long long *stamps = nullptr;
  int force_stages = 0;
  uint32_t dbg_flags = 0;
};

template <bool AK, bool BKM, int ACT, int DACT, bool LEAN, bool STAMPS>
int launch(b200_ctx *ctx, TcState *s, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mc, const TcParams &p,
           int grid) {
  auto kern = gemm_tc_kernel<AK, BKM, ACT, DACT, LEAN, STAMPS>;
  if (!s->attr_set.count((const void *)kern)) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    s->attr_set.insert((const void *)kern);
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + SMEM_EXTRA - STAGING_BYTES + p.staging_bytes;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (p.splitk == 2 && !p.reduce_add) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (s->pdl) {
    // the kernel may be launched while its predecessor on the stream is still draining: it does not touch
    // anything the predecessor writes before griddepcontrol.wait
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, p));
  LAUNCH_CHECK(ctx);
  return B200_OK;
}

// one per operand layout (gemm_tc_inst_*.cu): picks the variant <act, dact, lean, stamps> of that layout
int tc_launch_kk(b200_ctx *ctx, TcState *s, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mc, const TcParams &p, int grid, bool lean);
int tc_launch_kmn(b200_ctx *ctx, TcState *s, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mc, const TcParams &p, int grid, bool lean);
int tc_launch_mnk(b200_ctx *ctx, TcState *s, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mc, const TcParams &p, int grid, bool lean);
int tc_launch_mnmn(b200_ctx *ctx, TcState *s, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mc, const TcParams &p, int grid, bool lean);

}  // namespace b200tc
