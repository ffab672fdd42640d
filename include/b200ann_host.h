/*
 * b200ann_host.h -- C ABI of the host-side mirror (components, losses, SGD, trainer,
 * random) that language bindings attach to.  These are the calls the reference's Lua
 * binding layer makes on its C++ objects:
 *   ann.components.*              packages/ann/ann/binding/bind_ann_base.lua.cc:287-2187
 *   ann.loss.*                    packages/ann/loss/binding/bind_loss_functions.lua.cc:50-140
 *   ann.mlp.all_all.generate      packages/ann/ann/lua_src/annbase.lua:509-660
 *   trainable.supervised_trainer  packages/trainable/lua_src/supervised.lua
 *   ann.optimizer.sgd options     packages/ann/optimizer/lua_src/optimizer_sgd.lua:22-48
 *   random                        packages/basics/random/binding/bind_mtrand.lua.cc
 * Opaque handles, plain C types, int status (0 = ok; b200_last_error_string() explains).
 * Host float buffers are row-major float32.
 */
#ifndef B200ANN_HOST_H
#define B200ANN_HOST_H

#include "b200ann.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_component b200_component;
typedef struct b200_trainer b200_trainer;
typedef struct b200_random b200_random;

enum { B200_LOSS_MSE = 0, B200_LOSS_CROSS_ENTROPY = 1, B200_LOSS_MULTI_CLASS_CROSS_ENTROPY = 2,
       B200_LOSS_ZERO_ONE = 3 /* ann.loss.zero_one: validation / use only, not differentiable */ };
enum { B200_TOKEN_INPUT = 0, B200_TOKEN_OUTPUT = 1, B200_TOKEN_ERROR_INPUT = 2, B200_TOKEN_ERROR_OUTPUT = 3 };

int b200_add_launches(b200_ctx *ctx, uint64_t n);   /* graph replays: kernels launched per replay */

/* random(seed) -- bind_mtrand.lua.cc:45-75,85-166,168-215 */
b200_random *b200h_random_new(uint32_t seed);
void b200h_random_free(b200_random *r);
double b200h_random_rand(b200_random *r, double n);
uint32_t b200h_random_randint(b200_random *r, uint32_t n);     /* [0, n] */
int b200h_random_shuffle(b200_random *r, int size, int *out);   /* 0-based permutation */
/* the generator's 624 state words + the index of the next unread one (624: block exhausted): the layout
 * b200_dropout_mask continues from (b200_mt_state_bytes) */
int b200h_random_export_state(b200_random *r, uint32_t *words624, int32_t *next);

/* component constructors (ownership: the caller frees its handle; stacks share ownership) */
b200_component *b200h_stack_new(const char *name);
int b200h_stack_push(b200_component *stack, b200_component *child);
b200_component *b200h_hyperplane_new(const char *name, unsigned in, unsigned out, const char *dot_name,
                                     const char *bias_name, const char *dot_weights, const char *bias_weights);
b200_component *b200h_dot_product_new(const char *name, const char *weights, unsigned in, unsigned out);
b200_component *b200h_bias_new(const char *name, const char *weights, unsigned size);
/* kind: logistic tanh relu softmax log_softmax linear log_logistic softplus softsign leaky_relu hardtanh */
b200_component *b200h_actf_new(const char *kind, const char *name);
/* leaky_relu: p0 = leak; hardtanh: p0 = inf, p1 = sup (leaky_relu_actf_component.cc, hardtanh_actf_component.cc) */
b200_component *b200h_actf_new_ex(const char *kind, const char *name, float p0, float p1);
/* ann.components.actf.prelu{ size=, scalar=, name=, weights= }  (prelu_actf_component.cc) */
b200_component *b200h_prelu_new(const char *name, const char *weights, unsigned size, int scalar);
/* ann.components.dropout{ name=, size=, prob=, value=, random=, norm= }  (bind_ann_base.lua.cc:1643-1670);
 * the component copies the generator: its mask stream continues from the state `random` has now */
b200_component *b200h_dropout_new(const char *name, b200_random *random, float prob, float value, int norm, unsigned size);
b200_component *b200h_rewrap_new(const char *name, const int *size, int ndims);
b200_component *b200h_flatten_new(const char *name);
b200_component *b200h_convolution_new(const char *name, const char *weights, const int *kernel, const int *step,
                                      int ndims, int n);
b200_component *b200h_convolution_bias_new(const char *name, const char *weights, int n);
b200_component *b200h_max_pooling_new(const char *name, const int *kernel, const int *step, int ndims);
b200_component *b200h_mlp_generate(const char *topology);
void b200h_component_free(b200_component *c);

/* trainable.supervised_trainer(component, loss, bunch_size) */
b200_trainer *b200h_trainer_new(b200_ctx *ctx, b200_component *net, int loss_kind, int bunch_size);
void b200h_trainer_free(b200_trainer *t);
int b200h_trainer_build(b200_trainer *t, unsigned input, unsigned output);
int b200h_trainer_set_option(b200_trainer *t, const char *name, double value);
int b200h_trainer_get_option(b200_trainer *t, const char *name, double *value);
int b200h_trainer_set_layerwise_option(b200_trainer *t, const char *pattern, const char *name, double value);
int b200h_trainer_randomize_weights(b200_trainer *t, b200_random *rnd, double inf, double sup, int use_fanin,
                                    int use_fanout, const char *name_match /* may be NULL */);
/* flags: "fuse", "cuda_graph", "keep_gradients", "smooth_gradients" */
int b200h_trainer_set_flag(b200_trainer *t, const char *flag, int value);
int b200h_trainer_num_weights(b200_trainer *t, int *n);
int b200h_trainer_weight_name(b200_trainer *t, int i, char *buf, int buflen);
int b200h_trainer_weight_dims(b200_trainer *t, const char *name, int *dims2);
/* which: 0 weights, 1 gradients (of the last step), 2 momentum/update buffer (sgd update, rmsprop Eupdates,
 * adadelta update), 3 first optimizer state (adagrad/adadelta Egradients, rmsprop Erms), 4 second (adadelta Eupdates) */
int b200h_trainer_tensor_get(b200_trainer *t, const char *name, int which, float *host);
int b200h_trainer_tensor_set(b200_trainer *t, const char *name, int which, const float *host);
int b200h_trainer_num_parameters(b200_trainer *t, uint64_t *n);
int b200h_trainer_input_size(b200_trainer *t, int *n);
int b200h_trainer_output_size(b200_trainer *t, int *n);

/* train_step / validate_step (supervised.lua:725-862): host bunch in, bunch-mean loss out */
int b200h_trainer_train_step(b200_trainer *t, const float *x, const float *target, int bunch, float *loss,
                             float *loss_rows /* may be NULL */);
/* train_step(input, target, loss, optimizer, bunch_size, smooth, mask, max_gradients_norm) of
 * supervised.lua:725-821: smoothing_bunch = the bunch_size argument (0: the trainer's bunch_size, as the reference
 * does -- NOT the row count), max_gradients_norm = global gradient-norm clip (0: off) */
int b200h_trainer_train_step_ex(b200_trainer *t, const float *x, const float *target, int bunch, int smoothing_bunch,
                                double max_gradients_norm, float *loss, float *loss_rows);
int b200h_trainer_validate_step(b200_trainer *t, const float *x, const float *target, int bunch, float *loss,
                                float *loss_rows);
/* use_dataset (supervised.lua:1291-1430): forward only over n patterns; y holds n*output_size floats */
int b200h_trainer_use_dataset(b200_trainer *t, const float *x, int n, float *y);
/* ann.optimizer.{sgd,adagrad,rmsprop,adadelta}: selects the optimizer (options reset to its defaults, state
 * cleared); its state is readable / writable with b200h_trainer_tensor_get/set (which 2, 3, 4) */
int b200h_trainer_set_optimizer(b200_trainer *t, const char *name);
/* number of optimizer:execute calls so far -- with the weights and the state tensors this is everything
 * a checkpoint needs (optimizer_sgd.lua:102-119 exports options + count + update) */
int b200h_trainer_get_count(b200_trainer *t, int64_t *count);
int b200h_trainer_set_count(b200_trainer *t, int64_t count);
int b200h_trainer_set_loss_threshold(b200_trainer *t, float th);   /* zero_one: TH of the two-class case */
/* train_dataset / validate_dataset (supervised.lua:1149-1226,1291-1360): order = shuffled pattern
 * indices (0-based) or NULL for sequential; returns loss:get_accum_loss() */
int b200h_trainer_train_dataset(b200_trainer *t, const float *x, const float *target, int n, const int *order,
                                float *mean, float *variance);
int b200h_trainer_validate_dataset(b200_trainer *t, const float *x, const float *target, int n, float *mean,
                                   float *variance);
/* calculate (forward only): y must hold bunch*output_size floats */
int b200h_trainer_calculate(b200_trainer *t, const float *x, int bunch, float *y);
/* token of a named component after the last step: dims[4], data may be NULL to query the shape;
 * returns B200_ERR_BAD_ARG if the token was not materialised (inside a fused run) */
int b200h_trainer_component_token(b200_trainer *t, const char *component, int which, float *data, int *dims,
                                  int *ndims);
/* pipelined stepping: enqueue H2D of a bunch into the staging buffers / run a step on them
 * without any device->host traffic / read the accumulated loss statistics */
int b200h_trainer_stage(b200_trainer *t, const float *x_pinned, const float *target_pinned, int bunch);
int b200h_trainer_step_staged(b200_trainer *t, int bunch);
int b200h_trainer_loss_reset(b200_trainer *t);
int b200h_trainer_loss_get(b200_trainer *t, float *mean, float *variance);
int b200h_trainer_last_loss_async(b200_trainer *t, double *pinned_out);  /* enqueue D2H of the last bunch's loss sum */
/* data parallel replica group */
int b200h_trainer_set_data_parallel(b200_trainer *t, int nranks, int rank);
int b200h_trainer_broadcast_weights(b200_trainer *t);
/* replica group over NVLink peer memory: export this rank's 4 CUDA IPC handles (weights arena, gradient
 * arena, flag block, receive block: 256 bytes), exchange them over the host-side rendezvous, connect with
 * everybody's (nranks x 256 bytes, rank order).  Afterwards train steps use b200_dp_fused_update instead
 * of NCCL. */
int b200h_trainer_dp_export(b200_trainer *t, int nranks, void *handles256);
/* measurement hook (tools/dp_bench.py): `reps` fused updates over ALL parameters back to back, every rank
 * in step; returns the mean time of one (CUDA events).  Weights are advanced with whatever the gradient
 * arenas hold. */
int b200h_trainer_dp_bench(b200_trainer *t, int reps, float *us_per_update);
int b200h_trainer_dp_debug(b200_trainer *t, long long *stamps64);   /* 4 globaltimer stamps per bucket of the last step */
int b200h_trainer_dp_connect(b200_trainer *t, int nranks, int rank, const void *all_handles);
/* The bucket plan the fused update uses, as pure host logic (no device needed): tensors in arena order with
 * their gradient bytes; buckets [lo[i], hi[i]) of >= bucket_bytes or 32 tensors, at most 16 of them (the
 * threshold doubles until they fit).  Returns the number of buckets, or -1 when no plan exists (more than
 * 512 tensors: the trainer then takes the NCCL all-reduce path). */
int b200h_dp_bucket_plan(const size_t *tensor_bytes, int ntensors, size_t bucket_bytes, int *lo, int *hi, int cap);
/* replica group over symmetric memory (NVLS): every rank allocates b200h_trainer_dp_symmetric_bytes() bytes that
 * are mapped into every process (bases[r] = rank r's buffer as THIS process addresses it) and, where the fabric
 * supports it, bound to one multicast object (mc_base; NULL: P2P loads / stores on the same buffers).  The trainer
 * moves its weight and gradient arenas into its own buffer; afterwards the update of a step is one kernel per
 * bucket in which the NVSwitch sums the gradients (multimem.ld_reduce) and broadcasts the new weights (multimem.st). */
size_t b200h_trainer_dp_symmetric_bytes(b200_trainer *t);
int b200h_trainer_dp_connect_symmetric(b200_trainer *t, int nranks, int rank, void *const *bases, void *mc_base, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif
