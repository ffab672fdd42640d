// TF32 contraction on the 5th-generation tensor cores (sm_100a), host side: tensor-map cache, tile / split-K
// choice, variant dispatch.  The kernel lives in gemm_tc_kernel.cuh and is instantiated per operand layout in
// gemm_tc_inst_{kk,kmn,mnk,mnmn}.cu (forward = K-major x K-major, data gradient = K-major x MN-major, weight
// gradient = MN-major x MN-major).  Replaces the cublasSgemm call of mathcore/c_src/gemm.cu:248-327.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "gemm_tc_kernel.cuh"

using namespace b200tc;

namespace {

// ------------------------------------------------------------------ host side
TcState *state(b200_ctx *ctx) {
  if (!ctx->tc_state) {
    TcState *s = new TcState();
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      s->encode = (EncodeTiledFn)fn;
    if (const char *e = getenv("B200_PDL")) s->pdl = atoi(e) != 0;
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

// 3-D fp32 tensor map of an MN-major operand stored as [K rows][MN contiguous], row pitch ld floats:
// dim0 = 32 elements inside a chunk, dim1 = K rows, dim2 = MN/32 chunks (128 bytes apart).  The last chunk
// may run past MN: what it reads there (the row tail / the next row; the pool pads every block by 512 B)
// only reaches accumulator rows/columns >= M/N, which the epilogue never stores.  K is bounded exactly, so
// the contraction tail is zero-filled.
int get_map_mn(b200_ctx *ctx, TcState *s, const float *ptr, uint64_t mn, uint64_t k, uint64_t ld, uint32_t chunks_per_box,
               CUtensorMap *out) {
  MapKey key{ptr, mn, k, ld, 0x80000000u | chunks_per_box, 32, (uint32_t)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B};
  auto it = s->maps.find(key);
  if (it != s->maps.end()) { *out = it->second; return B200_OK; }
  cuuint64_t gdim[3] = {32, k, (mn + 31) / 32};
  cuuint64_t gstride[2] = {ld * sizeof(float), 128};
  cuuint32_t box[3] = {32, BK, chunks_per_box};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = s->encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)ptr, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200_set_error("cuTensorMapEncodeTiled (3-D, MN-major) failed (%d) for mn %llu k %llu ld %llu box {32,32,%u}", (int)r,
                   (unsigned long long)mn, (unsigned long long)k, (unsigned long long)ld, chunks_per_box);
    return B200_ERR_CUDA;
  }
  if (s->maps.size() > 4096) s->maps.clear();
  s->maps.emplace(key, m);
  *out = m;
  return B200_OK;
}

// Tile width and split-K decision.
// Measured on B200 (tools/gemm_stamps.py): one kind::tf32 M=128 instruction costs ~150 cycles for every
// N <= 256 (the A read from shared memory is the floor), so a 32-wide k-block costs >= ~610 cycles per
// CTA whatever the tile width; when many CTAs stream at once the L2->SM fabric (~7.3 KB/cycle for the
// whole chip) becomes the bound instead.  Candidates: N tile in {32..256 step 32} (16 for N <= 16),
// alone or as a split-K pair (two CTAs per tile, half of K each, partials exchanged through distributed
// shared memory) when that fits in one wave.  Cost in cycles:
//   waves * k-blocks * max(610, active*(16 KiB + 128 B * BN)/7300) + exposed epilogue (+ exchange).
void pick_tile(int M, int N, int K, int sm_count, bool allow_split, bool reduce_add, int *bn_out, int *split_out) {
  *split_out = 1;
  if (N <= 16) { *bn_out = 16; return; }
  const int tiles_m = (M + BM - 1) / BM;
  const int num_k = (K + BK - 1) / BK;
  int best = 32, best_split = 1;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 32; bn -= 32) {
    const int tn = (N + bn - 1) / bn;
    const long tiles = (long)tiles_m * tn;
    // exchange mode: one tile per cluster pair, both CTAs resident at once.  reduce-add mode: the k slices are
    // independent units of the persistent CTAs, so a problem with few output tiles and a long contraction (the
    // weight gradient of a convolution: 2 tiles, 1024 k-blocks) is cut into as many slices as it takes to fill
    // the device.  More than two slices per tile make the fp32 summation order run-dependent (last bits only).
    const int max_split = reduce_add ? 64 : 2;
    for (int split = 1; split <= max_split; split *= 2) {
      if (split >= 2 && (!allow_split || num_k < 4 * split)) continue;
      if (split == 2 && !reduce_add && 2 * tiles > sm_count) continue;
      if (split > 2 && tiles * (split / 2) >= sm_count) continue;     // already more units than SMs with fewer slices
      const long ctas = tiles * split;
      const long waves = (ctas + sm_count - 1) / sm_count;
      const double active = (double)(ctas < sm_count ? ctas : sm_count);
      const double kblock = fmax(610.0, active * (16384.0 + 128.0 * bn) / 7300.0);
      const double kblocks = (double)((num_k + split - 1) / split);
      double cost;
      if (reduce_add) {
        // persistent CTAs: the epilogue of a unit runs under the next unit's main loop, only the last one is exposed
        // (every CTA drains a whole 128 x bn tile)
        cost = (double)waves * kblocks * kblock + (bn / 32) * 350.0 + (waves > 1 ? 600.0 : 0.0) + 50.0 * split;
      } else {
        cost = (double)waves * kblocks * kblock + (bn / 32) * 500.0 / split;
        if (split == 2) cost += 1500.0 + (bn / 64) * 450.0;   // two cluster barriers + DSMEM exchange of half a tile
      }
      if (cost < best_cost - 1e-6) { best_cost = cost; best = bn; best_split = split; }
    }
  }
  *bn_out = best;
  *split_out = best_split;
}


// Second half of the launch plan, after the tile width and split are chosen (pick_tile or a forced value): split
// sanity, bytes per ring stage, staging tiles and ring depth.  Pure host arithmetic (also behind b200_debug_tc_plan,
// which the CPU tests sweep over shapes).
void finish_plan(TcParams &p, int M, int N, int K, bool b_k, int sms) {
  if (p.splitk > 2 && !p.reduce_add) p.splitk = 2;
  {
    // no empty k slice: every unit must issue at least one MMA (its accumulator is added as it stands)
    const int nk = (K + BK - 1) / BK;
    while (p.splitk > 1 && (p.splitk - 1) * ((nk + p.splitk - 1) / p.splitk) >= nk) p.splitk /= 2;
  }
  // the ring is as deep as shared memory allows: loads are latency/bandwidth bound, so bytes in flight matter
  p.stage_bytes = (uint32_t)A_BYTES + (b_k ? (uint32_t)p.BN * BK * 4 : (uint32_t)((p.BN + 31) / 32) * 4096u);
  {
    // launches with more tiles than CTAs (persistent CTAs, the epilogue of a tile runs under the next main loop)
    // fill the device by themselves: they take the whole shared memory and a staging tile per epilogue warp;
    // single-tile launches stay under SMEM_BUDGET so that light kernels can be resident beside them
    const long tiles_all = (long)((M + BM - 1) / BM) * ((N + p.BN - 1) / p.BN);
    const bool multi = p.reduce_add ? tiles_all * p.splitk > sms : (p.splitk != 2 && tiles_all > sms);
    p.staging_bytes = multi ? (uint32_t)(EPI_WARPS * 4096) : (uint32_t)STAGING_BYTES;
    const int budget = multi ? SMEM_MAX : SMEM_BUDGET;
    p.stages = (budget - (SMEM_EXTRA - STAGING_BYTES + (int)p.staging_bytes)) / (int)p.stage_bytes;
  }
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
}

}  // namespace

void gemm_tc_destroy(b200_ctx *ctx) {
  if (ctx->tc_state) {
    delete (TcState *)ctx->tc_state;
    ctx->tc_state = nullptr;
  }
}

// debug hook: per-CTA clock64 stamps (device buffer of 8*grid int64, or NULL to switch off)
extern "C" int b200_debug_tc_stamps(b200_ctx *ctx, long long *stamps_dev) {
  B200_ENTER(ctx);
  state(ctx)->stamps = stamps_dev;
  return B200_OK;
}
extern "C" int b200_debug_tc_stages(b200_ctx *ctx, int stages) {
  B200_ENTER(ctx);
  state(ctx)->force_stages = stages & 0xff;
  state(ctx)->dbg_flags = (uint32_t)stages >> 8;
  return B200_OK;
}

// debug hook: override descriptor fields {a_lbo,a_sbo,a_layout,a_kstep,b_lbo,b_sbo,b_layout,b_kstep}
extern "C" int b200_debug_tc_override(b200_ctx *ctx, int enable, const uint32_t *vals8, int force_bn) {
  B200_ENTER(ctx);
  TcState *s = state(ctx);
  s->dbg_on = enable != 0;
  if (vals8) memcpy(s->dbg, vals8, sizeof(s->dbg));
  s->force_bn = force_bn;
  return B200_OK;
}

// The launch plan of a contraction as pure host logic (no device needed): tile width, k slices, ring depth, whether
// the launch is planned as persistent multi-tile CTAs, and the grid.  reduce_add = the beta == 1 lean form.
extern "C" int b200_debug_tc_plan(int M, int N, int K, int b_kmajor, int sm_count, int reduce_add, int *bn, int *splitk,
                                  int *stages, int *stage_bytes, int *staging_bytes) {
  if (M <= 0 || N <= 0 || K <= 0 || sm_count <= 0) return B200_ERR_BAD_ARG;
  TcParams p{};
  p.reduce_add = reduce_add ? 1 : 0;
  pick_tile(M, N, K, sm_count, true, reduce_add != 0, &p.BN, &p.splitk);
  finish_plan(p, M, N, K, b_kmajor != 0, sm_count);
  if (bn) *bn = p.BN;
  if (splitk) *splitk = p.splitk;
  if (stages) *stages = p.stages;
  if (stage_bytes) *stage_bytes = (int)p.stage_bytes;
  if (staging_bytes) *staging_bytes = (int)p.staging_bytes;
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

  // SMs this launch plans for: the whole device, or the budget set while two contractions of a step are
  // meant to run side by side (b200_set_sm_budget)
  const int sms = ctx->sm_budget > 0 ? ctx->sm_budget : ctx->sm_count;
  TcParams p;
  p.M = M; p.N = N; p.K = K;
  // Lean variant: epilogue fixed at compile time, applied in the accumulator layout, TMA stores only.  Covers
  // what a training step launches: forward (bias + element-wise activation), data gradient (plain or with the
  // ReLU gate masks), weight gradient (alpha).  With beta == 1 the lean kernels ADD their tiles into C with TMA
  // reduce-add stores, which also makes split-K free of any exchange (the k slices are independent units).
  // Everything else (other beta, tanh / logistic derivatives, C that a tensor map cannot describe, an activation on
  // a layout that has no lean instance) takes the generic kernel.
  const bool c_tma_ok = (((uintptr_t)C) & 15) == 0 && (ldc & 3) == 0 && !(s->dbg_flags & 8);
  bool lean = c_tma_ok && (ep.beta == 0.0f || ep.beta == 1.0f) && !(s->dbg_flags & 64);
  if (ep.dact != B200_ACT_NONE) lean = lean && ep.dact == B200_ACT_RELU && ep.bias == nullptr && ep.act == B200_ACT_NONE && a_k && !b_k;
  if (ep.act != B200_ACT_NONE) lean = lean && a_k && b_k;
  if (ep.beta == 1.0f && (ep.act != B200_ACT_NONE || ep.bias != nullptr)) lean = false;   // act(sum) is not a sum of act(partials)
  p.reduce_add = (lean && ep.beta == 1.0f) ? 1 : 0;
  pick_tile(M, N, K, sms, !(s->dbg_flags & 16), p.reduce_add != 0, &p.BN, &p.splitk);
  if (s->force_bn) {
    p.BN = s->force_bn & 0xfff;
    p.splitk = (s->force_bn & 0x1000) ? 2 : 1;
  }
  finish_plan(p, M, N, K, b_k, sms);
  if (s->force_stages > 0 && s->force_stages < p.stages) p.stages = s->force_stages;
  p.ldc = ldc;
  p.C = C;
  p.ep = ep;
  p.stamps = s->stamps;
  p.dbg_flags = s->dbg_flags;
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
  else st = get_map_mn(ctx, s, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM / 32, &ma);
  if (st) return st;
  if (b_k) st = get_map(ctx, s, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)p.BN, CU_TENSOR_MAP_SWIZZLE_128B, &mb);
  else st = get_map_mn(ctx, s, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)((p.BN + 31) / 32), &mb);
  if (st) return st;

  // C goes out through TMA stores of 32x32 boxes when its layout allows (16-byte aligned base and pitch);
  // a box may only spill over the right edge of its N tile if that is also the edge of the matrix
  CUtensorMap mc = ma;
  const int tiles_n = (N + p.BN - 1) / p.BN;
  p.tma_store = (c_tma_ok && ((p.BN & 31) == 0 || tiles_n == 1)) ? 1 : 0;
  if (!p.tma_store) {
    if (p.reduce_add) return B200_ERR_UNSUPPORTED;   // (cannot happen: BN is a multiple of 32 or the tile spans N)
    lean = false;
  }
  if (lean && ep.dact == B200_ACT_RELU && p.BN > 256) lean = false;
  if (p.tma_store) {
    st = get_map(ctx, s, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B, &mc);
    if (st) return st;
  }

  const int tiles = ((M + BM - 1) / BM) * tiles_n;
  if (p.splitk >= 2 && ((!p.reduce_add && 2 * tiles > sms) || (p.BN & 31) != 0 || (K + BK - 1) / BK < p.splitk)) p.splitk = 1;
  // the exchange area of a split-K pair (half a tile per CTA) lives in the operand ring
  if (p.splitk == 2 && !p.reduce_add) {
    // receive area (4 quadrants x own chunks) + send staging (8 warps x half of the peer's chunks), 4 KiB each
    const int nch = (p.BN + 31) / 32, own0 = (nch + 1) / 2, own1 = nch - own0;
    const size_t need0 = (size_t)(4 * own0 + EPI_WARPS * ((own1 + 1) / 2)) * 4096u;
    const size_t need1 = (size_t)(4 * own1 + EPI_WARPS * ((own0 + 1) / 2)) * 4096u;
    if ((size_t)p.stages * p.stage_bytes < (need0 > need1 ? need0 : need1)) p.splitk = 1;
  }
  const int units = tiles * p.splitk;
  const int grid = p.reduce_add ? (units < sms ? units : sms) : (p.splitk == 2 ? 2 * tiles : (tiles < sms ? tiles : sms));
  if (a_k && b_k) return tc_launch_kk(ctx, s, ma, mb, mc, p, grid, lean);
  if (a_k && !b_k) return tc_launch_kmn(ctx, s, ma, mb, mc, p, grid, lean);
  if (!a_k && b_k) return tc_launch_mnk(ctx, s, ma, mb, mc, p, grid, lean);
  return tc_launch_mnmn(ctx, s, ma, mb, mc, p, grid, lean);
}
