// extern "C" surface of the host mirror (include/b200ann_host.h): opaque handles over the
// C++ classes of ann.h, exceptions turned into status codes + b200_last_error_string().
#include <string.h>

#include "../../../include/b200ann_host.h"
#include "ann.h"

void b200_set_error(const char *fmt, ...);

using namespace b200;

struct b200_component { ComponentPtr c; };
struct b200_trainer { SupervisedTrainer *t; std::shared_ptr<StackANNComponent> net; };
struct b200_random { MTRand r; explicit b200_random(uint32_t s) : r(s) {} };

#define API_TRY(body)                                  \
  try {                                                \
    body;                                              \
    return B200_OK;                                    \
  } catch (const Error &e) {                           \
    b200_set_error("%s", e.what());                    \
    return e.code ? e.code : B200_ERR_BAD_ARG;         \
  } catch (const std::exception &e) {                  \
    b200_set_error("%s", e.what());                    \
    return B200_ERR_BAD_ARG;                           \
  }

static b200_component *wrap(const ComponentPtr &c) { return new b200_component{c}; }
static std::string str(const char *s) { return s ? std::string(s) : std::string(); }
template <class F>
static b200_component *make(F f) {
  try {
    return wrap(f());
  } catch (const std::exception &e) {
    b200_set_error("%s", e.what());
    return nullptr;
  }
}

extern "C" {

b200_random *b200h_random_new(uint32_t seed) { return new b200_random(seed); }
void b200h_random_free(b200_random *r) { delete r; }
double b200h_random_rand(b200_random *r, double n) { return r->r.rand(n); }
uint32_t b200h_random_randint(b200_random *r, uint32_t n) { return r->r.randInt(n); }
int b200h_random_export_state(b200_random *r, uint32_t *words624, int32_t *next) {
  if (!r || !words624 || !next) { b200_set_error("random export_state: NULL"); return B200_ERR_BAD_ARG; }
  r->r.exportState(words624, next);
  return B200_OK;
}
int b200h_random_shuffle(b200_random *r, int size, int *out) {
  if (!r || !out || size <= 0) { b200_set_error("random shuffle: size must be >= 0"); return B200_ERR_BAD_ARG; }
  r->r.shuffle(size, out);
  return B200_OK;
}

b200_component *b200h_stack_new(const char *name) {
  return make([&] { return std::make_shared<StackANNComponent>(str(name)); });
}
int b200h_stack_push(b200_component *stack, b200_component *child) {
  API_TRY({
    auto *s = stack ? dynamic_cast<StackANNComponent *>(stack->c.get()) : nullptr;
    if (!s || !child) throw Error(B200_ERR_BAD_ARG, "stack push: not a stack / NULL child");
    s->pushComponent(child->c);
  })
}
b200_component *b200h_hyperplane_new(const char *name, unsigned in, unsigned out, const char *dot_name,
                                     const char *bias_name, const char *dot_weights, const char *bias_weights) {
  return make([&] { return makeHyperplane(str(name), in, out, str(dot_name), str(bias_name), str(dot_weights), str(bias_weights)); });
}
b200_component *b200h_dot_product_new(const char *name, const char *weights, unsigned in, unsigned out) {
  return make([&] { return std::make_shared<DotProductANNComponent>(str(name), str(weights), in, out); });
}
b200_component *b200h_bias_new(const char *name, const char *weights, unsigned size) {
  return make([&] { return std::make_shared<BiasANNComponent>(str(name), str(weights), size); });
}
b200_component *b200h_actf_new(const char *kind, const char *name) {
  return make([&] {
    const int act = actfFromName(str(kind));
    // defaults of the bindings: leaky_relu leak 0.01, hardtanh [-1, 1] (bind_ann_base.lua.cc)
    const float p0 = act == B200_ACT_LEAKY_RELU ? 0.01f : (act == B200_ACT_HARDTANH ? -1.0f : 0.0f);
    const float p1 = act == B200_ACT_HARDTANH ? 1.0f : 0.0f;
    return std::make_shared<ActivationFunctionANNComponent>(str(name), act, p0, p1);
  });
}
b200_component *b200h_actf_new_ex(const char *kind, const char *name, float p0, float p1) {
  return make([&] { return std::make_shared<ActivationFunctionANNComponent>(str(name), actfFromName(str(kind)), p0, p1); });
}
b200_component *b200h_prelu_new(const char *name, const char *weights, unsigned size, int scalar) {
  return make([&] { return std::make_shared<PReLUActfANNComponent>(str(name), str(weights), size, scalar != 0); });
}
b200_component *b200h_dropout_new(const char *name, b200_random *random, float prob, float value, int norm, unsigned size) {
  return make([&] {
    if (!random) throw Error(B200_ERR_BAD_ARG, "dropout: a random object is mandatory");
    return std::make_shared<DropoutANNComponent>(str(name), random->r, prob, value, norm != 0, size);
  });
}
b200_component *b200h_rewrap_new(const char *name, const int *size, int ndims) {
  return make([&] { return std::make_shared<RewrapANNComponent>(str(name), std::vector<int>(size, size + ndims)); });
}
b200_component *b200h_flatten_new(const char *name) {
  return make([&] { return std::make_shared<FlattenANNComponent>(str(name)); });
}
b200_component *b200h_convolution_new(const char *name, const char *weights, const int *kernel, const int *step,
                                      int ndims, int n) {
  return make([&] {
    std::vector<int> k(kernel, kernel + ndims), s;
    if (step) s.assign(step, step + ndims);
    return std::make_shared<ConvolutionANNComponent>(str(name), str(weights), k, s, n);
  });
}
b200_component *b200h_convolution_bias_new(const char *name, const char *weights, int n) {
  return make([&] { return std::make_shared<ConvolutionBiasANNComponent>(str(name), str(weights), n); });
}
b200_component *b200h_max_pooling_new(const char *name, const int *kernel, const int *step, int ndims) {
  return make([&] {
    std::vector<int> k(kernel, kernel + ndims), s;
    if (step) s.assign(step, step + ndims);
    return std::make_shared<MaxPoolingANNComponent>(str(name), k, s);
  });
}
b200_component *b200h_mlp_generate(const char *topology) {
  return make([&] { return mlpAllAllGenerate(str(topology)); });
}
void b200h_component_free(b200_component *c) { delete c; }

b200_trainer *b200h_trainer_new(b200_ctx *ctx, b200_component *net, int loss_kind, int bunch_size) {
  try {
    auto s = net ? std::dynamic_pointer_cast<StackANNComponent>(net->c) : nullptr;
    if (!s) {
      // a single component is wrapped in a stack, as the trainer only drives stacks
      if (!net) throw Error(B200_ERR_BAD_ARG, "trainer: NULL component");
      s = std::make_shared<StackANNComponent>("stack");
      s->pushComponent(net->c);
    }
    auto *t = new b200_trainer{nullptr, s};
    t->t = new SupervisedTrainer(ctx, s, loss_kind, bunch_size);
    return t;
  } catch (const std::exception &e) {
    b200_set_error("%s", e.what());
    return nullptr;
  }
}
void b200h_trainer_free(b200_trainer *t) {
  if (!t) return;
  delete t->t;
  delete t;
}
int b200h_trainer_build(b200_trainer *t, unsigned input, unsigned output) { API_TRY(t->t->build(input, output)) }
int b200h_trainer_set_option(b200_trainer *t, const char *name, double v) { API_TRY(t->t->setOption(str(name), v)) }
int b200h_trainer_get_option(b200_trainer *t, const char *name, double *v) { API_TRY(*v = t->t->getOption(str(name))) }
int b200h_trainer_set_layerwise_option(b200_trainer *t, const char *pattern, const char *name, double v) {
  API_TRY(t->t->setLayerwiseOption(str(pattern), str(name), v))
}
int b200h_trainer_randomize_weights(b200_trainer *t, b200_random *rnd, double inf, double sup, int use_fanin,
                                    int use_fanout, const char *name_match) {
  API_TRY({
    if (!rnd) throw Error(B200_ERR_BAD_ARG, "randomize_weights: random object is mandatory");
    t->t->randomizeWeights(rnd->r, inf, sup, use_fanin != 0, use_fanout != 0, str(name_match));
  })
}
int b200h_trainer_set_flag(b200_trainer *t, const char *flag, int value) {
  API_TRY({
    std::string f = str(flag);
    // every flag changes kernel arguments or the topology of the captured step graphs
    t->t->invalidateGraphs();
    if (f == "keep_gradients") t->t->sgd_dirty_public();
    if (f == "fuse") t->net->fuse = value != 0;
    else if (f == "cuda_graph") t->t->use_cuda_graph = value != 0;
    else if (f == "branches") t->t->use_branches = value != 0;
    else if (f == "zero_accumulate") t->t->zero_accumulate = value != 0;
    else if (f == "fuse_output_layer") t->t->fuse_output_layer = value != 0;
    else if (f == "dp_fused") {   // 0: back to the NCCL all-reduce path (1 needs a connected replica group)
      if (value != 0 && t->t->dp_group.nranks < 2) throw Error(B200_ERR_BAD_ARG, "dp_fused needs b200h_trainer_dp_connect");
      t->t->dp_fused = value != 0;
    }
    else if (f == "keep_gradients") t->t->keep_gradients = value != 0;
    else if (f == "smooth_gradients") t->t->smooth_gradients = value != 0;
    else throw Error(B200_ERR_BAD_ARG, "unknown flag " + f);
  })
}
int b200h_trainer_num_weights(b200_trainer *t, int *n) { API_TRY(*n = (int)t->t->weights_order.size()) }
int b200h_trainer_weight_name(b200_trainer *t, int i, char *buf, int buflen) {
  API_TRY({
    const std::string &s = t->t->weights_order.at(i);
    if ((int)s.size() + 1 > buflen) throw Error(B200_ERR_BAD_ARG, "buffer too small");
    memcpy(buf, s.c_str(), s.size() + 1);
  })
}
int b200h_trainer_weight_dims(b200_trainer *t, const char *name, int *dims2) {
  API_TRY({
    MatrixPtr w = t->t->weights_table.at(str(name));
    dims2[0] = w->dim(0);
    dims2[1] = w->cols();
  })
}
static MatrixPtr pick(b200_trainer *t, const char *name, int which) {
  if (which < 0 || which > 4) throw Error(B200_ERR_BAD_ARG, "which: 0 weights, 1 gradients, 2 update, 3/4 optimizer state");
  if (which >= 3) t->t->ensureOptimizerState();
  MatrixDict &d = which == 0 ? t->t->weights_table
                  : which == 1 ? t->t->grads
                  : which == 2 ? t->t->updates
                  : which == 3 ? t->t->state1 : t->t->state2;
  auto it = d.find(str(name));
  if (it == d.end()) throw Error(B200_ERR_BAD_ARG, "unknown weights name " + str(name));
  return it->second;
}
int b200h_trainer_tensor_get(b200_trainer *t, const char *name, int which, float *host) {
  API_TRY(pick(t, name, which)->toHost(host))
}
int b200h_trainer_tensor_set(b200_trainer *t, const char *name, int which, const float *host) {
  API_TRY({
    MatrixPtr m = pick(t, name, which);
    m->fromHost(host);
    check(b200_sync(t->t->ctx));
  })
}
int b200h_trainer_num_parameters(b200_trainer *t, uint64_t *n) { API_TRY(*n = t->t->numParameters()) }
int b200h_trainer_input_size(b200_trainer *t, int *n) { API_TRY(*n = (int)t->net->getInputSize()) }
int b200h_trainer_output_size(b200_trainer *t, int *n) { API_TRY(*n = (int)t->net->getOutputSize()) }

int b200h_trainer_train_step(b200_trainer *t, const float *x, const float *target, int bunch, float *loss,
                             float *loss_rows) {
  API_TRY({
    float l = t->t->trainStep(x, target, bunch, loss_rows);
    if (loss) *loss = l;
  })
}
int b200h_trainer_train_step_ex(b200_trainer *t, const float *x, const float *target, int bunch, int smoothing_bunch,
                                double max_gradients_norm, float *loss, float *loss_rows) {
  API_TRY({
    if (max_gradients_norm < 0) throw Error(B200_ERR_BAD_ARG, "max_gradients_norm must be >= 0");
    float l = t->t->trainStep(x, target, bunch, loss_rows, smoothing_bunch, max_gradients_norm);
    if (loss) *loss = l;
  })
}
int b200h_trainer_use_dataset(b200_trainer *t, const float *x, int n, float *y) { API_TRY(t->t->useDataset(x, n, y)) }
int b200h_trainer_set_optimizer(b200_trainer *t, const char *name) {
  API_TRY({
    const std::string n = str(name);
    const int k = n == "sgd" ? B200_OPT_SGD : n == "adagrad" ? B200_OPT_ADAGRAD : n == "rmsprop" ? B200_OPT_RMSPROP
                  : n == "adadelta" ? B200_OPT_ADADELTA : -1;
    if (k < 0) throw Error(B200_ERR_BAD_ARG, "unknown optimizer " + n + " (sgd, adagrad, rmsprop, adadelta)");
    t->t->setOptimizer(k);
  })
}
int b200h_trainer_get_count(b200_trainer *t, int64_t *count) { API_TRY(*count = t->t->getCount()) }
int b200h_trainer_set_count(b200_trainer *t, int64_t count) { API_TRY(t->t->setCount(count)) }
int b200h_trainer_set_loss_threshold(b200_trainer *t, float th) { API_TRY(t->t->loss.TH = th) }
int b200h_trainer_validate_step(b200_trainer *t, const float *x, const float *target, int bunch, float *loss,
                                float *loss_rows) {
  API_TRY({
    float l = t->t->validateStep(x, target, bunch, loss_rows);
    if (loss) *loss = l;
  })
}
int b200h_trainer_train_dataset(b200_trainer *t, const float *x, const float *target, int n, const int *order,
                                float *mean, float *variance) {
  API_TRY(t->t->trainDataset(x, target, n, order, mean, variance))
}
int b200h_trainer_validate_dataset(b200_trainer *t, const float *x, const float *target, int n, float *mean,
                                   float *variance) {
  API_TRY(t->t->validateDataset(x, target, n, mean, variance))
}
int b200h_trainer_calculate(b200_trainer *t, const float *x, int bunch, float *y) {
  API_TRY({
    const int in = (int)t->net->getInputSize();
    MatrixPtr dx = Matrix::create(t->t->ctx, std::vector<int>{bunch, in});
    dx->fromHost(x);
    MatrixPtr out = t->t->calculate(dx);
    out->toHost(y);
  })
}
int b200h_trainer_component_token(b200_trainer *t, const char *component, int which, float *data, int *dims,
                                  int *ndims) {
  API_TRY({
    ComponentDict comps;
    MatrixDict w = t->t->weights_table;
    ANNComponent *found = nullptr;
    std::vector<ANNComponent *> todo{t->net.get()};
    while (!todo.empty() && !found) {
      ANNComponent *c = todo.back();
      todo.pop_back();
      if (c->getName() == str(component)) found = c;
      if (auto *s = dynamic_cast<StackANNComponent *>(c))
        for (auto &k : s->components) todo.push_back(k.get());
    }
    if (!found) throw Error(B200_ERR_BAD_ARG, "unknown component " + str(component));
    MatrixPtr m = which == B200_TOKEN_INPUT ? found->getInput()
                  : which == B200_TOKEN_OUTPUT ? found->getOutput()
                  : which == B200_TOKEN_ERROR_INPUT ? found->getErrorInput()
                                                    : found->getErrorOutput();
    if (!m) throw Error(B200_ERR_BAD_ARG, "token not materialised (inside a fused run; set flag fuse=0)");
    *ndims = (int)m->dims.size();
    for (int i = 0; i < *ndims && i < 4; ++i) dims[i] = m->dims[i];
    if (data) m->toHost(data);
  })
}
int b200h_trainer_stage(b200_trainer *t, const float *x, const float *target, int bunch) {
  API_TRY(t->t->stage(x, target, bunch))
}
int b200h_trainer_step_staged(b200_trainer *t, int bunch) { API_TRY(t->t->stepStaged(bunch)) }
int b200h_trainer_loss_reset(b200_trainer *t) { API_TRY(t->t->loss.reset()) }
int b200h_trainer_loss_get(b200_trainer *t, float *mean, float *variance) {
  API_TRY(t->t->loss.getAccumLoss(mean, variance))
}
int b200h_trainer_last_loss_async(b200_trainer *t, double *pinned_out) {
  API_TRY(check(b200_memcpy_d2h(t->t->ctx, pinned_out, t->t->loss.stats_dev + 3, sizeof(double))))
}
int b200h_trainer_set_data_parallel(b200_trainer *t, int nranks, int rank) {
  API_TRY(t->t->setDataParallel(nranks, rank))
}
int b200h_trainer_broadcast_weights(b200_trainer *t) { API_TRY(t->t->broadcastWeights()) }
int b200h_trainer_dp_bench(b200_trainer *t, int reps, float *us_per_update) { API_TRY(*us_per_update = t->t->dpBench(reps)) }
int b200h_trainer_dp_debug(b200_trainer *t, long long *stamps64) {
  API_TRY(
      if (!t->t->dp_flags) throw b200::Error(B200_ERR_BAD_ARG, "no replica group");
      b200::check(b200_sync(t->t->ctx));
      b200::check(b200_memcpy_d2h(t->t->ctx, stamps64, (char *)t->t->dp_flags + b200_dp_debug_offset(), 64 * sizeof(long long)));
      b200::check(b200_sync(t->t->ctx)))
}
int b200h_trainer_dp_export(b200_trainer *t, int nranks, void *handles256) {
  API_TRY(t->t->dpExport(nranks, (unsigned char *)handles256))
}
int b200h_dp_bucket_plan(const size_t *tensor_bytes, int ntensors, size_t bucket_bytes, int *lo, int *hi, int cap) {
  if (ntensors < 0 || (ntensors > 0 && !tensor_bytes)) return -1;
  std::vector<std::pair<int, int>> plan;
  if (!dp_bucket_plan(std::vector<size_t>(tensor_bytes, tensor_bytes + ntensors), bucket_bytes, &plan)) return -1;
  for (int i = 0; i < (int)plan.size() && i < cap; ++i) {
    if (lo) lo[i] = plan[i].first;
    if (hi) hi[i] = plan[i].second;
  }
  return (int)plan.size();
}
size_t b200h_trainer_dp_symmetric_bytes(b200_trainer *t) {
  try {
    if (!t->t->weights_arena) return 0;
    return SupervisedTrainer::dpSymmetricBytes(t->t->weights_arena->size());
  } catch (...) {
    return 0;
  }
}
int b200h_trainer_dp_connect_symmetric(b200_trainer *t, int nranks, int rank, void *const *bases, void *mc_base, size_t bytes) {
  API_TRY(t->t->dpConnectSymmetric(nranks, rank, bases, mc_base, bytes))
}
int b200h_trainer_dp_connect(b200_trainer *t, int nranks, int rank, const void *all_handles) {
  API_TRY(t->t->dpConnect(nranks, rank, (const unsigned char *)all_handles))
}

}  // extern "C"
