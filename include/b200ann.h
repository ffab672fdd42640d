/*
 * b200ann.h -- C ABI of the B200-native training hot path behind the APRIL-ANN API.
 *
 * Plain C: raw device pointers, int dims / leading dimensions, an explicit
 * per-device context.  No C++ or torch types cross this boundary.  All matrices
 * are float32, row-major.  Every entry point returns 0 on success and a non-zero
 * status otherwise; b200_last_error_string() describes the last failure of the
 * calling thread (the reference aborts through ERROR_EXIT,
 * packages/basics/util/c_src/error_print.h:55-77; a host shim turns a non-zero
 * status into that).  Kernels are enqueued on the context's stream and are
 * asynchronous; b200_sync() waits.
 *
 * Each group cites the reference interface it replaces (paths relative to the
 * reference checkout).
 */
#ifndef B200ANN_H
#define B200ANN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;

/* status codes (the reference's exit codes 150-163 are CUDA init/alloc/copy:
 * packages/basics/mathcore/c_src/gpu_helper.h:57-68) */
enum {
  B200_OK = 0,
  B200_ERR_BAD_ARG = 128,      /* same code the reference uses for bad sizes/types */
  B200_ERR_NOT_BUILT = 129,
  B200_ERR_CUDA = 150,
  B200_ERR_ALLOC = 151,
  B200_ERR_NCCL = 160,
  B200_ERR_UNSUPPORTED = 161
};

/* activation kinds (packages/ann/ann/c_src/<kind>_actf_component.cc) */
enum {
  B200_ACT_NONE = 0,
  B200_ACT_LOGISTIC = 1,   /* 1/(1+e^-x);   derivative from output y(1-y), clamped */
  B200_ACT_TANH = 2,       /* 2/(1+e^-x)-1; derivative from output 0.5(1-y^2), clamped */
  B200_ACT_RELU = 3,       /* derivative (x>0), evaluated as (y>0) */
  B200_ACT_SOFTMAX = 4,
  B200_ACT_LOG_SOFTMAX = 5,
  B200_ACT_LINEAR = 6,
  /* the cheap activations of SURVEY.md 8(f)4 (activation_function_kernels.cu:52-182); element-wise launches
   * (b200_actf_fwd_ex / _bwd_ex), not fused into the contraction epilogue */
  B200_ACT_LOG_LOGISTIC = 7,  /* x<-10 ? x : -log1p(e^-x); derivative cancelled by cross-entropy (identity) */
  B200_ACT_SOFTPLUS = 8,      /* x>10 ? x : log1p(e^x);    derivative from the INPUT: logistic(x) */
  B200_ACT_SOFTSIGN = 9,      /* x/(1+|x|);                derivative from the OUTPUT, clamped: 1/(1+|y|)^2 */
  B200_ACT_LEAKY_RELU = 10,   /* p0 = leak;                derivative from the INPUT: x>0 ? 1 : leak */
  B200_ACT_HARDTANH = 11      /* clamp(x, p0, p1);         derivative from the INPUT: (x<p0 || x>p1) ? 0 : 1 */
};

/* math modes for the contractions */
enum {
  B200_MATH_FP32 = 0,      /* FFMA fp32 accumulate: the parity mode (tolerance 1e-5 rel-L2) */
  B200_MATH_TF32 = 1       /* tcgen05 kind::tf32, fp32 accumulate in TMEM (tolerance 2e-3 rel-L2) */
};

/* ------------------------------------------------------------------ runtime
 * replaces GPUHelper (packages/basics/mathcore/c_src/gpu_helper.h:42-148) and the
 * device half of GPUMirroredMemoryBlock
 * (packages/basics/mathcore/c_src/gpu_mirrored_memory_block.h:178-304,476,518). */
const char *b200_last_error_string(void);
int b200_device_count(int *count);
int b200_create(int device, b200_ctx **out);
int b200_destroy(b200_ctx *ctx);
int b200_set_math_mode(b200_ctx *ctx, int mode);
int b200_get_math_mode(b200_ctx *ctx, int *mode);
int b200_sm_count(b200_ctx *ctx, int *count);
void *b200_stream(b200_ctx *ctx);                 /* cudaStream_t of the compute stream */
int b200_sync(b200_ctx *ctx);
int b200_malloc(b200_ctx *ctx, void **dptr, size_t bytes);   /* stream-ordered caching pool */
int b200_free(b200_ctx *ctx, void *dptr);
int b200_pool_trim(b200_ctx *ctx);                            /* return cached blocks to the driver */
int b200_host_alloc(void **hptr, size_t bytes);               /* pinned host memory */
int b200_host_free(void *hptr);
int b200_memcpy_h2d(b200_ctx *ctx, void *dst, const void *src, size_t bytes);  /* async on the stream */
int b200_memcpy_d2h(b200_ctx *ctx, void *dst, const void *src, size_t bytes);  /* async on the stream */
int b200_memcpy_d2d(b200_ctx *ctx, void *dst, const void *src, size_t bytes);
int b200_memset_zero(b200_ctx *ctx, void *dst, size_t bytes);
/* Side branches of a training step.  The reference runs one kernel after another on stream 0
 * (packages/basics/mathcore/c_src/gpu_helper.cu:28-35); here the work that is not on the critical
 * path of a step (weight gradients, per-tensor SGD, loss statistics) is issued on side streams and
 * becomes parallel branches of the step's CUDA graph.
 *   b200_branch_begin(i)   until b200_branch_end, launches go to side stream i, ordered after what
 *                          the main stream holds now
 *   b200_branch_wait(i)    the current stream waits for what side stream i holds now
 *   b200_branch_join_all   the main stream waits for every open branch (call before the step ends)
 *   b200_set_sm_budget     SMs one persistent contraction may plan for when two of them are meant
 *                          to run side by side (0 = the whole device) */
int b200_branch_begin(b200_ctx *ctx, int branch);
int b200_branch_end(b200_ctx *ctx);
int b200_branch_wait(b200_ctx *ctx, int branch);
int b200_branch_join_all(b200_ctx *ctx);
int b200_set_sm_budget(b200_ctx *ctx, int sms);
/* point-to-point edge: record marks what the current stream (main or the open branch) holds now, wait makes the
 * current stream wait for exactly that; id in [0, 8) */
int b200_fence_record(b200_ctx *ctx, int id);
int b200_fence_wait(b200_ctx *ctx, int id);
/* CUDA-event timing on the compute stream (bench / roofline) */
int b200_event_create(void **ev);
int b200_event_destroy(void *ev);
int b200_event_record(b200_ctx *ctx, void *ev);
int b200_event_elapsed_ms(void *start, void *stop, float *ms);   /* synchronises on stop */
/* number of kernels this library launched on ctx since creation (bench "gpu_launches") */
int b200_launch_count(b200_ctx *ctx, uint64_t *count);

/* ------------------------------------------------------------------ BLAS seam
 * replaces AprilMath::doGemm/doGemv/doGer/doAxpy/doScal/doCopy
 * (packages/basics/mathcore/c_src/cblas_headers.h:240-535, gemm.cu:248-327,
 * gemv.cu:44, ger.cu:42, axpy.cu:42, scal.cu:38, copy.cu:43).
 * Row-major; trans = 0 (no transpose) or 1 (transpose), as CblasNoTrans/CblasTrans. */
int b200_sgemm(b200_ctx *ctx, int transA, int transB, int M, int N, int K,
               float alpha, const float *A, int lda, const float *B, int ldb,
               float beta, float *C, int ldc);
int b200_sgemv(b200_ctx *ctx, int transA, int M, int N, float alpha, const float *A, int lda,
               const float *x, int incx, float beta, float *y, int incy);
int b200_sger(b200_ctx *ctx, int M, int N, float alpha, const float *x, int incx,
              const float *y, int incy, float *A, int lda);
int b200_saxpy(b200_ctx *ctx, size_t n, float alpha, const float *x, float *y);
int b200_sscal(b200_ctx *ctx, size_t n, float alpha, float *x);
int b200_scopy(b200_ctx *ctx, size_t n, const float *x, float *y);
int b200_cmul(b200_ctx *ctx, size_t n, const float *x, float *y);          /* y *= x  (matCmul) */
int b200_sum(b200_ctx *ctx, size_t n, const float *x, float *out);         /* *out = sum x (device scalar) */
int b200_nrm2sq(b200_ctx *ctx, size_t n, const float *x, float *out);      /* *out += sum x^2 (device scalar) */

/* ------------------------------------------------------------------ fused dense layer
 * replaces DotProductANNComponent + BiasANNComponent + ActivationFunctionANNComponent
 * (packages/ann/ann/c_src/dot_product_component.cc:63-98,123-152,194-216;
 *  bias_component.cc:46-73,87-122; activation_function_component.cc:48-120).
 *   fwd : Y[M,N]  = act( X[M,K] . W[N,K]^T + b[N] )                    (bias may be NULL)
 *   bwd_data : dX[M,K] = ( dY[M,N] . W[N,K] ) (.) act'(Yprev[M,K])      (act_prev NONE => plain)
 *              where Yprev is the OUTPUT of the previous activation
 *   bwd_weight : dW[N,K] = beta*dW + scale * dY[M,N]^T . X[M,K]
 *                db[N]   = beta*db + scale * sum_m dY[m,:]              (db may be NULL)
 *   scale carries the trainer's 1/sqrt(shared_count*bunch)
 *   (packages/trainable/lua_src/supervised.lua:797-803). */
int b200_linear_fwd(b200_ctx *ctx, int M, int N, int K, const float *X, int ldx,
                    const float *W, int ldw, const float *bias, int act, float *Y, int ldy);
int b200_linear_bwd_data(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy,
                         const float *W, int ldw, int act_prev, const float *Yprev, int ldyp,
                         float *dX, int lddx);
/* dX += (dY . W) (.) act'(Yprev): accumulate form of bwd_data (the reference's beta = 1 GEMM into a zeroed matrix,
 * convolution_component.cc:265).  On the tensor-core path the tiles are added with TMA reduce-add stores, which
 * makes split-K exchange-free. */
int b200_linear_bwd_data_acc(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy,
                             const float *W, int ldw, int act_prev, const float *Yprev, int ldyp,
                             float *dX, int lddx);
int b200_linear_bwd_weight(b200_ctx *ctx, int M, int N, int K, const float *dY, int lddy,
                           const float *X, int ldx, float scale, float beta,
                           float *dW, int lddw, float *db);

/* ------------------------------------------------------------------ element-wise
 * replaces ANN::Kernels::apply* (packages/ann/ann/c_src/activation_function_kernels.cu:60-182)
 * and the bias axpy loops (mathcore/c_src/axpy.cu:111,275). */
int b200_actf_fwd(b200_ctx *ctx, int act, size_t n, const float *x, float *y);
/* dx = act'(.) * dy ; y = activation output (relu: y>0 <=> x>0) */
int b200_actf_bwd(b200_ctx *ctx, int act, size_t n, const float *y, const float *dy, float *dx);
/* parametrised / input-derivative activations (p0, p1: leak / inf, sup).  bwd: x = activation input, y = output;
 * either may be NULL when the kind does not need it */
int b200_actf_fwd_ex(b200_ctx *ctx, int act, float p0, float p1, size_t n, const float *x, float *y);
int b200_actf_bwd_ex(b200_ctx *ctx, int act, float p0, float p1, size_t n, const float *x, const float *y,
                     const float *dy, float *dx);
/* PReLU (prelu_actf_component.cc:55-114): y = x>0 ? x : a[n]*x (scalar: one a); da = beta*da + scale * sum (x<0)*x*dy;
 * tmp = M*N floats of workspace */
int b200_prelu_fwd(b200_ctx *ctx, int M, int N, const float *x, const float *a, int scalar, float *y);
int b200_prelu_bwd(b200_ctx *ctx, int M, int N, const float *x, const float *a, int scalar, const float *dy, float *dx);
int b200_prelu_grad(b200_ctx *ctx, int M, int N, const float *x, const float *dy, int scalar, float scale, float beta,
                    float *da, float *tmp);
/* dropout (dropout_component.cc:67-134): mask[i] = (i-th rand() of the component's MT19937) < prob ? 0 : 1, drawn
 * on the device in the reference's stream order.  mt_state_dev: b200_mt_state_bytes() bytes = 624 state words
 * (after MTRand's reload) + int32 position of the next unread word.  b200_mask_apply: y = mask<0.5 ? value : x */
size_t b200_mt_state_bytes(void);
int b200_dropout_mask(b200_ctx *ctx, void *mt_state_dev, size_t n, float prob, float *mask);
int b200_mask_apply(b200_ctx *ctx, size_t n, const float *x, const float *mask, float value, float *y);
int b200_bias_fwd(b200_ctx *ctx, int M, int N, const float *x, const float *b, float *y);
int b200_bias_grad(b200_ctx *ctx, int M, int N, const float *dy, int lddy, float scale,
                   float beta, float *db);

/* ------------------------------------------------------------------ row-wise
 * replaces applySoftmax/applyLogSoftmax/applySoftmaxDerivative
 * (activation_function_kernels.cu:184-355) and the loss kernels
 * (packages/ann/loss/c_src/loss_kernels.cu:38-266, multiclass_cross_entropy_loss_function.cc:48-71,
 *  mse_loss_function.cc:41-109, cross_entropy_loss_function.cc:41-118). */
int b200_softmax_fwd(b200_ctx *ctx, int M, int C, const float *x, float *y);
int b200_log_softmax_fwd(b200_ctx *ctx, int M, int C, const float *x, float *y);
int b200_softmax_bwd(b200_ctx *ctx, int M, int C, const float *y, const float *dy, float *dx);
/* loss_rows[M] and/or grad[M,C] (either may be NULL) */
int b200_mcce_loss_grad(b200_ctx *ctx, int M, int C, const float *logp, const float *target,
                        float *loss_rows, float *grad);
int b200_mse_loss_grad(b200_ctx *ctx, int M, int C, const float *out, const float *target,
                       float *loss_rows, float *grad);
int b200_ce_loss_grad(b200_ctx *ctx, int M, int C, const float *log_out, const float *target,
                      float *loss_rows, float *grad);
/* zero-one loss (zero_one_loss_function.cc:39-132): C == 1: (out > TH) != (target > 0.5); else arg-max of the row
 * against the arg-max of a dense target (target_cols == C) or a 1-based class label (target_cols == 1) */
int b200_zero_one_loss(b200_ctx *ctx, int M, int C, const float *out, const float *target, int target_cols, float TH,
                       float *loss_rows);
/* one pass: logits -> logp (may be NULL), loss_rows, grad = exp(clamp(logp)) - target */
int b200_log_softmax_mcce_fused(b200_ctx *ctx, int M, int C, const float *logits,
                                const float *target, float *logp, float *loss_rows, float *grad);
/* running statistics over loss rows on the device:
 * stats[0] += sum(rows), stats[1] += sum(rows^2), stats[2] += M   (float64 on the device) */
int b200_loss_accumulate(b200_ctx *ctx, int M, const float *loss_rows, double *stats);

/* Output layer of a classifier in one launch (N <= 16 classes, K % 4 == 0): the dot_product + bias forward
 * (dot_product_component.cc:63-98, bias_component.cc:46-73), log_softmax
 * (activation_function_kernels.cu:289-325), the multi-class cross-entropy rows and gradient
 * (loss_kernels.cu:171-185, multiclass_cross_entropy_loss_function.cc:61-71) and the data gradient of the
 * layer below, dX = (grad . W) (.) act'(X) with X that layer's activation output
 * (dot_product_component.cc:123-152 + the actf derivative).  logits / logp / loss_rows / grad / dX may be
 * NULL.  Returns B200_ERR_UNSUPPORTED for shapes it does not cover. */
int b200_output_layer_fused(b200_ctx *ctx, int M, int N, int K, const float *X, int ldx, const float *W, int ldw,
                            const float *bias, const float *target, float *logits, float *logp,
                            float *loss_rows, float *grad, int dact, float *dX, int lddx);

/* ------------------------------------------------------------------ SGD
 * replaces ann.optimizer.sgd:execute (packages/ann/optimizer/lua_src/optimizer_sgd.lua:50-100)
 * and its helpers (base_optimizer.lua:28-49).  One launch updates every tensor:
 *   g += l2*w ; u = mt*u (or 0) ; u += lrd*g ; w -= u ; [L1 truncate] ; [prune subnormals]
 * with lrd = lr / (1 + decay*count) computed on the device from *count (int64, device). */
typedef struct {
  float *w;              /* weights, updated in place            */
  float *g;              /* gradients (already scaled)           */
  float *u;              /* momentum/update buffer               */
  uint64_t n;            /* elements                             */
  int32_t rows, cols;    /* for max_norm_penalty (rows of w)     */
  float lr, momentum, weight_decay, l1_norm, max_norm_penalty;
  int32_t pad_;
} b200_sgd_tensor;
int b200_sgd_multi_tensor(b200_ctx *ctx, int ntensors, const b200_sgd_tensor *tensors_dev,
                          const b200_sgd_tensor *tensors_host, double decay,
                          const int64_t *count_dev, int write_back_grad);
/* same; flags bit 0 = write the regularised gradient back, bit 1 = this is the last update launch of
 * the step: its last CTA to finish adds 1 to count_dev[0] (count_dev[1] is the ticket word, zero
 * between launches), which replaces a separate counter launch */
#define B200_SGD_WRITE_BACK_GRAD 1
#define B200_SGD_INCREMENT_COUNT 2
int b200_sgd_multi_tensor_ex(b200_ctx *ctx, int ntensors, const b200_sgd_tensor *tensors_dev,
                             const b200_sgd_tensor *tensors_host, double decay, int64_t *count_dev,
                             int flags);
int b200_counter_increment(b200_ctx *ctx, int64_t *count_dev);
/* global gradient-norm clip of train_step (packages/trainable/lua_src/supervised.lua:805-811): one norm over
 * the whole (flat) gradient arena, then g *= max_norm/norm when norm > max_norm.  norm2sq_dev: device float. */
int b200_grad_clip(b200_ctx *ctx, size_t n, float *grads, float max_norm, float *norm2sq_dev);
/* adagrad / rmsprop / adadelta (packages/ann/optimizer/lua_src/optimizer_{adagrad,rmsprop,adadelta}.lua) as
 * one multi-tensor launch.  State per tensor: u (rmsprop Eupdate / adadelta lr*update), s1 (Egradient / Erms),
 * s2 (adadelta Eupdate).  b200_optimizer_lookahead is rmsprop's w -= momentum*Eupdate before the gradient is
 * evaluated (optimizer_rmsprop.lua:44-51). */
enum { B200_OPT_SGD = 0, B200_OPT_ADAGRAD = 1, B200_OPT_RMSPROP = 2, B200_OPT_ADADELTA = 3 };
typedef struct {
  float *w, *g, *u, *s1, *s2;
  uint64_t n;
  int32_t rows, cols;
  float lr, momentum, decay, epsilon, weight_decay, max_norm_penalty;
  int32_t write_back_grad, pad_;
} b200_opt_tensor;
int b200_optimizer_multi_tensor(b200_ctx *ctx, int algo, int ntensors, const b200_opt_tensor *tensors_dev,
                                const b200_opt_tensor *tensors_host, int64_t *count_dev, int increment_count);
int b200_optimizer_lookahead(b200_ctx *ctx, int ntensors, const b200_opt_tensor *tensors_dev,
                             const b200_opt_tensor *tensors_host);

/* ------------------------------------------------------------------ replica group over NVLink peer memory
 * New relative to the reference (single device, gpu_helper.h:65-68).  b200_dp_fused_update is the
 * data-parallel update as one kernel: gradient reduce-scatter (by default every rank PULLS the peers'
 * gradients of its own shard with P2P loads; B200_DP_PUSH=1 selects the push variant, P2P stores into the
 * owners' receive blocks), the SGD step above on the rank's shard, all-gather of the updated weights
 * (P2P stores into every replica).  The arenas of every rank are mapped into each process with CUDA IPC
 * (b200_ipc_export / _import: the 64-byte handle travels over the host-side rendezvous).  Ordering
 * between GPUs: per (bucket, rank) step tags in `flags` (b200_dp_flags_bytes() bytes per rank, zeroed
 * once); b200_dp_wait, at the end of a step, returns once every rank's shard of every bucket has landed. */
#define B200_DP_MAX_RANKS 8
typedef struct {
  float *grads[B200_DP_MAX_RANKS];      /* base of every rank's gradient arena (own entry = local pointer) */
  float *weights[B200_DP_MAX_RANKS];    /* base of every rank's weight arena */
  float *recv[B200_DP_MAX_RANKS];       /* every rank's receive block: (nranks-1) x arena_elems floats, one arena-shaped
                                           slot per source rank (sources in rank order, the owner skipped) */
  long long *flags[B200_DP_MAX_RANKS];  /* every rank's flag block */
  uint64_t arena_elems;                 /* floats of one arena (same on every rank) */
  int32_t nranks, rank;
  /* NVLS: multicast addresses of the gradient / weight arenas when every rank's arenas are bound to one multicast
   * object (NULL otherwise).  multimem.ld_reduce on mc_grads sums the replicas inside the NVSwitch, multimem.st on
   * mc_weights writes every replica. */
  float *mc_grads, *mc_weights;
} b200_dp_group;
int b200_ipc_export(b200_ctx *ctx, void *dptr, void *handle64);
int b200_ipc_import(b200_ctx *ctx, const void *handle64, void **dptr);
int b200_ipc_close(b200_ctx *ctx, void *dptr);
size_t b200_dp_flags_bytes(void);
size_t b200_dp_debug_offset(void);   /* bring-up: 4 globaltimer stamps per bucket live behind the tags */
int b200_dp_fused_update(b200_ctx *ctx, const b200_dp_group *grp, int ntensors, const b200_sgd_tensor *tensors_dev,
                         const b200_sgd_tensor *tensors_host, double decay, int64_t *count_dev, int bucket);
/* count_dev: FOUR int64 on the device: [0] optimizer step count (lr decay), [1] ticket word of the update kernels,
 * [2] replica-group epoch (the cross-GPU tags are epoch + 1; advanced by b200_dp_wait, never set back), [3] reserved */
int b200_dp_wait(b200_ctx *ctx, const b200_dp_group *grp, int nbuckets, int64_t *count_dev);

/* ------------------------------------------------------------------ convolution / pooling
 * replaces ConvolutionANNComponent, ConvolutionBiasANNComponent, MaxPoolingANNComponent
 * (packages/ann/ann/c_src/convolution_component.cc:135-354, convolution_bias_component.cc:120-222,
 *  maxpooling_component.cc:129-260).  NCHW, valid convolution, W[n, C*kh*kw] flattened in
 *  (plane,row,col) order. */
int b200_conv2d_fwd(b200_ctx *ctx, int B, int C, int H, int W_, int n, int kh, int kw, int sh, int sw,
                    const float *x, const float *w, const float *bias, int act, float *y);
int b200_conv2d_bwd_data(b200_ctx *ctx, int B, int C, int H, int W_, int n, int kh, int kw, int sh,
                         int sw, const float *dy, const float *w, float *dx);
int b200_conv2d_bwd_weight(b200_ctx *ctx, int B, int C, int H, int W_, int n, int kh, int kw, int sh,
                           int sw, const float *dy, const float *x, float scale, float beta,
                           float *dw, float *db);
/* y[b,p,:] = x[b,p,:] + bias[p] ; db[p] = beta*db[p] + scale * sum_{b,pixels} dy[b,p,:] */
int b200_conv_bias_fwd(b200_ctx *ctx, int B, int n, int HW, const float *x, const float *bias, float *y);
int b200_conv_bias_grad(b200_ctx *ctx, int B, int n, int HW, const float *dy, float scale, float beta,
                        float *db);
int b200_maxpool_fwd(b200_ctx *ctx, int B, int C, int H, int W_, int kh, int kw, int sh, int sw,
                     const float *x, float *y, int32_t *argmax);
int b200_maxpool_bwd(b200_ctx *ctx, int B, int C, int H, int W_, int kh, int kw, int sh, int sw,
                     const float *dy, const int32_t *argmax, float *dx);

/* ------------------------------------------------------------------ input staging
 * replaces DataSetToken::getPatternBunch (packages/basics/dataset/c_src/datasetToken.h:182-209):
 * out[i,:] = data[idx[i],:] on the device. */
int b200_gather_rows(b200_ctx *ctx, int nrows, int cols, const float *data, const int32_t *idx,
                     float *out);

/* ------------------------------------------------------------------ data parallel
 * new (the reference has no multi-GPU path: gpu_helper.h:65-68 hard-codes device 0).
 * NCCL communicator per context; id is ncclUniqueId bytes from rank 0. */
int b200_comm_unique_id(void *id128);                             /* 128 bytes out */
int b200_comm_init(b200_ctx *ctx, int nranks, int rank, const void *id128);
int b200_comm_destroy(b200_ctx *ctx);
int b200_allreduce_sum(b200_ctx *ctx, float *buf, size_t n);      /* in place, on the comm stream,
                                                                     ordered after/before the compute stream */
int b200_allreduce_sum_f64(b200_ctx *ctx, double *buf, size_t n);
/* bucketed overlap: the all-reduce runs on the communication stream after everything enqueued so far
 * on the compute stream; b200_comm_wait(slot) makes the compute stream wait for that bucket
 * (slot in [0, 16)).  Both are capturable in a CUDA graph. */
int b200_allreduce_sum_async(b200_ctx *ctx, float *buf, size_t n, int slot);
int b200_comm_wait(b200_ctx *ctx, int slot);
int b200_broadcast(b200_ctx *ctx, float *buf, size_t n, int root);

#ifdef __cplusplus
}
#endif
#endif /* B200ANN_H */
