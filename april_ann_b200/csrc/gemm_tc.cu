// tcgen05 / TMEM / TMA TF32 contraction (placeholder until the tensor path lands:
// reports "unsupported" so the dispatcher uses the FFMA kernel).
#include "common.cuh"

int gemm_tc(b200_ctx *, int, int, int, int, int, const float *, int, const float *, int, float *, int,
            const GemmEpilogue &) {
  return B200_ERR_UNSUPPORTED;
}
void gemm_tc_destroy(b200_ctx *) {}
