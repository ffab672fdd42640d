// tcgen05 contraction, operand layout "mnmn": the kernel variants of this layout (gemm_tc_kernel.cuh).
#include <cuda.h>
#include <math.h>

#include "gemm_tc_kernel.cuh"

namespace b200tc {

int tc_launch_mnmn(b200_ctx *ctx, TcState *s, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mc, const TcParams &p,
                 int grid, bool lean) {
  if (lean && p.stamps) return launch<false, false, B200_ACT_NONE, B200_ACT_NONE, true, true>(ctx, s, ma, mb, mc, p, grid);
  if (lean) {
    return launch<false, false, B200_ACT_NONE, B200_ACT_NONE, true, false>(ctx, s, ma, mb, mc, p, grid);
  }
  if (p.stamps) return launch<false, false, B200_ACT_NONE, B200_ACT_NONE, false, true>(ctx, s, ma, mb, mc, p, grid);
  return launch<false, false, B200_ACT_NONE, B200_ACT_NONE, false, false>(ctx, s, ma, mb, mc, p, grid);
}

}  // namespace b200tc
