// tcgen05 contraction, operand layout "kmn": the kernel variants of this layout (gemm_tc_kernel.cuh).
#include <cuda.h>
#include <math.h>

#include "gemm_tc_kernel.cuh"

namespace b200tc {

int tc_launch_kmn(b200_ctx *ctx, TcState *s, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mc, const TcParams &p,
                 int grid, bool lean) {
  if (lean && p.stamps) return launch<true, false, B200_ACT_NONE, B200_ACT_NONE, true, true>(ctx, s, ma, mb, mc, p, grid);
  if (lean) {
    if (p.ep.dact == B200_ACT_RELU) return launch<true, false, B200_ACT_NONE, B200_ACT_RELU, true, false>(ctx, s, ma, mb, mc, p, grid);
    return launch<true, false, B200_ACT_NONE, B200_ACT_NONE, true, false>(ctx, s, ma, mb, mc, p, grid);
  }
  if (p.stamps) return launch<true, false, B200_ACT_NONE, B200_ACT_NONE, false, true>(ctx, s, ma, mb, mc, p, grid);
  return launch<true, false, B200_ACT_NONE, B200_ACT_NONE, false, false>(ctx, s, ma, mb, mc, p, grid);
}

}  // namespace b200tc
