// BLAS seam and the fused dense-layer entry points of the C ABI.
// b200_sgemm & co. slot in where AprilMath::doGemm/doGemv/doGer sit
// (mathcore/c_src/cblas_headers.h:240-535); b200_linear_* are what
// DotProductANNComponent+BiasANNComponent+ActivationFunctionANNComponent call.
#include "common.cuh"

int gemm_dispatch(b200_ctx *ctx, int transA, int transB, int M, int N, int K, const float *A,
                  int lda, const float *B, int ldb, float *C, int ldc, const GemmEpilogue &ep) {
  if (ctx->math_mode == B200_MATH_TF32) {
    int st = gemm_tc(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, ep);
    if (st != B200_ERR_UNSUPPORTED) return st;
    // shape or alignment outside the tensor path (tiny / skinny / unaligned): FFMA kernel
  }
  return gemm_simt(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, ep);
}

extern "C" int b200_sgemm(b200_ctx *ctx, int transA, int transB, int M, int N, int K, float alpha,
                          const float *A, int lda, const float *B, int ldb, float beta, float *C,
                          int ldc) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && A && B && C, "NULL pointer");
  ARG_CHECK(M >= 0 && N >= 0 && K >= 0, "negative dimension");
  ARG_CHECK(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "leading dimension too small");
  GemmEpilogue ep;
  ep.alpha = alpha;
  ep.beta = beta;
  return gemm_dispatch(ctx, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, ep);
}

// y[M or N] = alpha * op(A) x + beta*y, A is MxN row-major (gemv.cu:44)
extern "C" int b200_sgemv(b200_ctx *ctx, int transA, int M, int N, float alpha, const float *A,
                          int lda, const float *x, int incx, float beta, float *y, int incy) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && A && x && y, "NULL pointer");
  ARG_CHECK(incx >= 1 && incy >= 1 && lda >= N, "bad stride");
  GemmEpilogue ep;
  ep.alpha = alpha;
  ep.beta = beta;
  const int rows = transA ? N : M, k = transA ? M : N;
  // C[rows,1] (ldc=incy) = op(A)[rows,k] . B[k,1]  with B[k*incx + 0]
  return gemm_simt(ctx, transA, 0, rows, 1, k, A, lda, x, incx, y, incy, ep);
}

// A[M,N] += alpha * x y^T (ger.cu:42)
extern "C" int b200_sger(b200_ctx *ctx, int M, int N, float alpha, const float *x, int incx,
                         const float *y, int incy, float *A, int lda) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && A && x && y, "NULL pointer");
  ARG_CHECK(incx >= 1 && incy >= 1 && lda >= N, "bad stride");
  GemmEpilogue ep;
  ep.alpha = alpha;
  ep.beta = 1.0f;
  // K = 1: op(A)[m,0] = x[m*incx], op(B)[0,n] = y[n*incy] (transB=1, ldb=incy)
  return gemm_simt(ctx, 0, 1, M, N, 1, x, incx, y, incy, A, lda, ep);
}

extern "C" int b200_linear_fwd(b200_ctx *ctx, int M, int N, int K, const float *X, int ldx,
                               const float *W, int ldw, const float *bias, int act, float *Y,
                               int ldy) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && X && W && Y, "NULL pointer");
  ARG_CHECK(act == B200_ACT_NONE || act == B200_ACT_LOGISTIC || act == B200_ACT_TANH ||
                act == B200_ACT_RELU || act == B200_ACT_LINEAR,
            "only element-wise activations fuse into the contraction");
  if (skinny_applicable(M, N, K))
    return skinny_fwd(ctx, M, N, K, X, ldx, W, ldw, bias, (act == B200_ACT_LINEAR) ? B200_ACT_NONE : act, Y, ldy);
  GemmEpilogue ep;
  ep.bias = bias;
  ep.act = (act == B200_ACT_LINEAR) ? B200_ACT_NONE : act;
  // Y = X . W^T : op(A)=X (no trans), op(B)=W^T (trans)
  return gemm_dispatch(ctx, 0, 1, M, N, K, X, ldx, W, ldw, Y, ldy, ep);
}

extern "C" int b200_linear_bwd_data(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy,
                                    const float *W, int ldw, int act_prev, const float *Yprev,
                                    int ldyp, float *dX, int lddx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dY && W && dX, "NULL pointer");
  const bool has_prev = (act_prev == B200_ACT_LOGISTIC || act_prev == B200_ACT_TANH || act_prev == B200_ACT_RELU);
  if (has_prev) ARG_CHECK(Yprev, "Yprev is required when act_prev is set");
  if (skinny_applicable(M, N, K))
    return skinny_bwd_data(ctx, M, N, K, dY, lddy, W, ldw, has_prev ? act_prev : B200_ACT_NONE, Yprev, ldyp, dX, lddx);
  GemmEpilogue ep;
  if (has_prev) {
    ep.dact = act_prev;
    ep.dsrc = Yprev;
    ep.ld_dsrc = ldyp;
  }
  // dX[M,K] = dY[M,N] . W[N,K] : contraction over N
  return gemm_dispatch(ctx, 0, 0, M, K, N, dY, lddy, W, ldw, dX, lddx, ep);
}

// dX += ( dY . W ) (.) act'(Yprev): the accumulate form.  With the tensor-core path and a ReLU (or no) derivative the
// tiles are ADDED into dX with TMA reduce-add stores, so split-K needs no exchange between CTAs; the caller
// provides a zeroed (or partially accumulated) dX.
extern "C" int b200_linear_bwd_data_acc(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy, const float *W,
                                        int ldw, int act_prev, const float *Yprev, int ldyp, float *dX, int lddx) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dY && W && dX, "NULL pointer");
  const bool has_prev = (act_prev == B200_ACT_LOGISTIC || act_prev == B200_ACT_TANH || act_prev == B200_ACT_RELU);
  if (has_prev) ARG_CHECK(Yprev, "Yprev is required when act_prev is set");
  GemmEpilogue ep;
  ep.beta = 1.0f;
  if (has_prev) {
    ep.dact = act_prev;
    ep.dsrc = Yprev;
    ep.ld_dsrc = ldyp;
  }
  return gemm_dispatch(ctx, 0, 0, M, K, N, dY, lddy, W, ldw, dX, lddx, ep);
}

extern "C" int b200_linear_bwd_weight(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy,
                                      const float *X, int ldx, float scale, float beta, float *dW,
                                      int lddw, float *db) {
  B200_ENTER(ctx);
  ARG_CHECK(ctx && dY && X && dW, "NULL pointer");
  if (skinny_applicable(M, N, K)) return skinny_bwd_weight(ctx, M, N, K, dY, lddy, X, ldx, scale, beta, dW, lddw, db);
  GemmEpilogue ep;
  ep.alpha = scale;
  ep.beta = beta;
  // dW[N,K] = dY^T[N,M] . X[M,K] : contraction over the bunch M
  int st = gemm_dispatch(ctx, 1, 0, N, K, M, dY, lddy, X, ldx, dW, lddw, ep);
  if (st) return st;
  if (db) return b200_bias_grad(ctx, M, N, dY, lddy, scale, beta, db);
  return B200_OK;
}
