// Loss functions, SGD options and the supervised trainer of the host mirror.
//   ann.loss.*                   packages/ann/loss/c_src/*_loss_function.cc, loss_function.h:34-120
//   ann.optimizer.sgd            packages/ann/optimizer/lua_src/optimizer_sgd.lua:22-100
//   trainable.supervised_trainer packages/trainable/lua_src/supervised.lua:575-862,1149-1430
// The whole train step (forward, loss, backward, weight gradients, [all-reduce], SGD, loss
// statistics, step counter) is enqueued on one stream and, once warm, replayed as a CUDA graph.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "../../../include/b200ann_host.h"
#include "ann.h"

int b200_make_current(b200_ctx *ctx);   // runtime.cu: makes the context's device the current one

namespace b200 {

static void cudaCheck(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw Error(B200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// matrices created while a step is being captured must outlive the graph
static thread_local std::vector<MatrixPtr> *g_capture_registry = nullptr;
void registerCapturedMatrix(const MatrixPtr &m) {
  if (g_capture_registry) g_capture_registry->push_back(m);
}

// ------------------------------------------------------------------ loss
LossFunction::LossFunction(b200_ctx *ctx, int kind, unsigned size) : kind(kind), size(size), ctx(ctx) {
  if (kind == LOSS_MULTI_CLASS_CROSS_ENTROPY && size > 0 && size < 3)
    throw Error(128, "Multi class cross entropy is only allowed for multi-class problems (three or more output "
                     "log softmax neurons). Use cross entropy instead.");
  void *p;
  check(b200_malloc(ctx, &p, 4 * sizeof(double)));
  stats_dev = (double *)p;
  reset();
}
LossFunction::~LossFunction() {
  if (stats_dev) b200_free(ctx, stats_dev);
}
void LossFunction::reset() { check(b200_memset_zero(ctx, stats_dev, 4 * sizeof(double))); }

static void checkLossArgs(const LossFunction &l, const MatrixPtr &in, const MatrixPtr &tg) {
  if (!in || !tg) throw Error(128, "Incorrect input token type, expected token matrix");
  if (l.kind == LOSS_ZERO_ONE) {
    // a dense target or a [bunch,1] vector of 1-based class labels (zero_one_loss_function.cc:82-118)
    if (in->dims.size() != 2 || tg->rows() != in->rows() || (tg->cols() != in->cols() && tg->cols() != 1))
      throw Error(128, "Incorrect target matrix bunch_size");
    return;
  }
  if (in->size() != tg->size()) throw Error(128, "Different token sizes found: input vs target");
  if (in->dims.size() != 2) throw Error(128, "loss input must be a 2-dimensional matrix");
  if (l.size != 0 && (unsigned)in->cols() != l.size) throw Error(128, "loss input size mismatch");
}
MatrixPtr LossFunction::computeLoss(const MatrixPtr &in, const MatrixPtr &tg) {
  checkLossArgs(*this, in, tg);
  MatrixPtr rows = Matrix::create(ctx, std::vector<int>{in->rows()});
  const int M = in->rows(), C = in->cols();
  if (kind == LOSS_ZERO_ONE) check(b200_zero_one_loss(ctx, M, C, in->data, tg->data, tg->cols(), TH, rows->data));
  else if (kind == LOSS_MSE) check(b200_mse_loss_grad(ctx, M, C, in->data, tg->data, rows->data, nullptr));
  else if (kind == LOSS_CROSS_ENTROPY) check(b200_ce_loss_grad(ctx, M, C, in->data, tg->data, rows->data, nullptr));
  else check(b200_mcce_loss_grad(ctx, M, C, in->data, tg->data, rows->data, nullptr));
  return rows;
}
MatrixPtr LossFunction::computeGradient(const MatrixPtr &in, const MatrixPtr &tg) {
  if (kind == LOSS_ZERO_ONE) throw Error(128, "NON DIFERENTIABLE LOSS FUNCTION");   // zero_one_loss_function.cc:124-129
  checkLossArgs(*this, in, tg);
  MatrixPtr g = Matrix::create(ctx, in->dims);
  const int M = in->rows(), C = in->cols();
  if (kind == LOSS_MSE) check(b200_mse_loss_grad(ctx, M, C, in->data, tg->data, nullptr, g->data));
  else if (kind == LOSS_CROSS_ENTROPY) check(b200_ce_loss_grad(ctx, M, C, in->data, tg->data, nullptr, g->data));
  else check(b200_mcce_loss_grad(ctx, M, C, in->data, tg->data, nullptr, g->data));
  return g;
}
void LossFunction::fusedLogSoftmaxMCCE(const MatrixPtr &logits, const MatrixPtr &tg, MatrixPtr &logp,
                                       MatrixPtr &rows, MatrixPtr &grad) {
  checkLossArgs(*this, logits, tg);
  const int M = logits->rows(), C = logits->cols();
  if (!logp) logp = Matrix::create(ctx, logits->dims);
  rows = Matrix::create(ctx, std::vector<int>{M});
  check(b200_log_softmax_mcce_fused(ctx, M, C, logits->data, tg->data, logp->data, rows->data,
                                    grad ? grad->data : nullptr));
}
void LossFunction::accumLoss(const MatrixPtr &rows) {
  check(b200_loss_accumulate(ctx, (int)rows->size(), rows->data, stats_dev));
}
void LossFunction::getAccumLoss(float *mean, float *variance) {
  double h[4];
  check(b200_memcpy_d2h(ctx, h, stats_dev, 3 * sizeof(double)));
  check(b200_sync(ctx));
  const double n = h[2];
  const double m = n > 0 ? h[0] / n : 0.0;
  double v = n > 1 ? (h[1] - h[0] * h[0] / n) / (n - 1) : 0.0;
  if (v < 0) v = 0;
  if (mean) *mean = (float)m;
  if (variance) *variance = (float)v;
}

// ------------------------------------------------------------------ optimizer options
SGDOptimizer::SGDOptimizer() { setKind(B200_OPT_SGD); }
void SGDOptimizer::setKind(int k) {
  kind = k;
  layerwise_options.clear();
  count = 0;
  if (k == B200_OPT_SGD)   // optimizer_sgd.lua:39-47
    global_options = {{"learning_rate", 0.01}, {"momentum", 0.0}, {"decay", 1e-05},
                      {"weight_decay", 0.0},   {"L1_norm", 0.0},  {"max_norm_penalty", 0.0}};
  else if (k == B200_OPT_ADAGRAD)   // optimizer_adagrad.lua:33-39
    global_options = {{"learning_rate", 1.0}, {"decay", 0.95}, {"epsilon", 1e-06}, {"weight_decay", 0.0}, {"max_norm_penalty", 0.0}};
  else if (k == B200_OPT_RMSPROP)   // optimizer_rmsprop.lua:35-42
    global_options = {{"learning_rate", 0.01}, {"momentum", 0.0}, {"decay", 0.99}, {"epsilon", 1e-06},
                      {"weight_decay", 0.0},   {"max_norm_penalty", 0.0}};
  else if (k == B200_OPT_ADADELTA)  // optimizer_adadelta.lua:36-43
    global_options = {{"learning_rate", 1.0}, {"momentum", 0.0}, {"decay", 0.95}, {"epsilon", 1e-06},
                      {"weight_decay", 0.0},  {"max_norm_penalty", 0.0}};
  else
    throw Error(B200_ERR_BAD_ARG, "unknown optimizer");
}
bool SGDOptimizer::validOption(const std::string &n) const { return global_options.count(n) != 0; }
void SGDOptimizer::setOption(const std::string &n, double v) {
  if (!validOption(n)) throw Error(B200_ERR_BAD_ARG, "Not recognized option " + n);
  global_options[n] = v;
}
double SGDOptimizer::getOption(const std::string &n) const {
  if (!validOption(n)) throw Error(B200_ERR_BAD_ARG, "Not recognized option " + n);
  return global_options.at(n);
}
void SGDOptimizer::setLayerwiseOption(const std::string &layer, const std::string &n, double v) {
  if (!validOption(n)) throw Error(B200_ERR_BAD_ARG, "Not recognized option " + n);
  if (n == "decay" && kind == B200_OPT_SGD)
    throw Error(B200_ERR_BAD_ARG, "decay option cannot be defined layerwise, only globally");
  layerwise_options[layer][n] = v;
}
double SGDOptimizer::getOptionOf(const std::string &layer, const std::string &n) const {
  auto it = layerwise_options.find(layer);
  if (it != layerwise_options.end()) {
    auto jt = it->second.find(n);
    if (jt != it->second.end()) return jt->second;
  }
  return getOption(n);
}

// ------------------------------------------------------------------ trainer
struct SupervisedTrainer::Graph {
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
  std::vector<MatrixPtr> keep;
  MatrixPtr rows, out;
  const float *x_ptr = nullptr, *t_ptr = nullptr;
  uint64_t launches = 0;
  int warm = 0;
  ~Graph() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
  }
};

SupervisedTrainer::SupervisedTrainer(b200_ctx *ctx, const std::shared_ptr<StackANNComponent> &net, int loss_kind,
                                     int bunch_size)
    : ctx(ctx), net(net), loss(ctx, loss_kind, 0), bunch_size(bunch_size) {
  if (!ctx) throw Error(B200_ERR_CUDA, "trainer needs a device context: this build has no CPU path");
  if (const char *e = getenv("B200_DP_BUCKET_MB")) dp_bucket_bytes = (size_t)(atof(e) * 1048576.0);
  // tuning switches (A/B measurements): side branches of the step, and dgrad || wgrad of a layer
  if (const char *e = getenv("B200_BRANCHES")) use_branches = atoi(e) != 0;
  if (const char *e = getenv("B200_FUSE_OUTPUT")) fuse_output_layer = atoi(e) != 0;
  if (const char *e = getenv("B200_SGD_AS_READY")) sgd_as_ready = atoi(e) != 0;
  if (const char *e = getenv("B200_CONCURRENT_BWD")) net->contraction_mode = atoi(e);
  if (const char *e = getenv("B200_ZERO_ACC")) zero_accumulate = atoi(e) != 0;
  void *p;
  // [0] optimizer step count, [1] ticket, [2] replica-group epoch, [3] reserved (include/b200ann.h, b200_dp_wait)
  check(b200_malloc(ctx, &p, 4 * sizeof(int64_t)));
  count_dev = (int64_t *)p;
  check(b200_memset_zero(ctx, count_dev, 4 * sizeof(int64_t)));
}
SupervisedTrainer::~SupervisedTrainer() {
  b200_sync(ctx);
  for (auto &kv : graphs) delete kv.second;
  if (copy_stream) {
    cudaStreamSynchronize((cudaStream_t)copy_stream);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy((cudaEvent_t)ev_copied[i]);
      cudaEventDestroy((cudaEvent_t)ev_trained[i]);
    }
    cudaStreamDestroy((cudaStream_t)copy_stream);
  }
  if (count_dev) b200_free(ctx, count_dev);
  if (sgd_dev) b200_free(ctx, sgd_dev);
  if (sgd_light_dev) b200_free(ctx, sgd_light_dev);
  if (opt_dev) b200_free(ctx, opt_dev);
  if (norm_dev) b200_free(ctx, norm_dev);
  for (void *p : dp_imported) b200_ipc_close(ctx, p);
  if (dp_flags && dp_flags_owned) b200_free(ctx, dp_flags);
  if (dp_recv) b200_free(ctx, dp_recv);
}

void SupervisedTrainer::build(unsigned input, unsigned output) {
  build_in = input;
  build_out = output;
  net->setContext(ctx);
  // pass 1: discover weight names and shapes
  MatrixDict discovered;
  ComponentDict comps;
  net->build(input, output, discovered, comps);
  weights_order.clear();
  for (auto &kv : discovered) weights_order.push_back(kv.first);
  std::sort(weights_order.begin(), weights_order.end());  // supervised.lua:680-681
  // pass 2: re-home every tensor in flat arenas (weights / gradients / momentum), 512-byte aligned,
  // laid out in REVERSE layer order = the order in which the backward pass finishes gradients, so
  // that a bucket of finished gradients is one contiguous range for the all-reduce
  arena_order.clear();
  {
    const auto &flat = net->flatComponents();
    for (auto it = flat.rbegin(); it != flat.rend(); ++it) {
      if (!(*it)->hasWeightsName()) continue;
      const std::string &wn = (*it)->getWeightsName();
      if (discovered.count(wn) && std::find(arena_order.begin(), arena_order.end(), wn) == arena_order.end())
        arena_order.push_back(wn);
    }
    for (auto &n : weights_order)
      if (std::find(arena_order.begin(), arena_order.end(), n) == arena_order.end()) arena_order.push_back(n);
  }
  std::map<std::string, size_t> offs;
  size_t total = 0;
  for (auto &n : arena_order) {
    offs[n] = total;
    total += (discovered[n]->size() + 127) & ~size_t(127);
  }
  total_params = 0;
  if (total > (size_t)INT32_MAX)
    throw Error(B200_ERR_UNSUPPORTED, "models with more than 2^31-1 (padded) parameters are not supported by the flat arenas");
  state1_arena.reset();
  state2_arena.reset();
  state1.clear();
  state2.clear();
  weights_arena = Matrix::create(ctx, std::vector<int>{(int)total});
  grads_arena = Matrix::create(ctx, std::vector<int>{(int)total});
  updates_arena = Matrix::create(ctx, std::vector<int>{(int)total});
  weights_arena->zeros();
  grads_arena->zeros();
  updates_arena->zeros();
  weights_table.clear();
  grads.clear();
  updates.clear();
  for (size_t i = 0; i < weights_order.size(); ++i) {
    const std::string &n = weights_order[i];
    const std::vector<int> &d = discovered[n]->dims;
    weights_table[n] = Matrix::view(weights_arena, offs[n], d);
    grads[n] = Matrix::view(grads_arena, offs[n], d);
    updates[n] = Matrix::view(updates_arena, offs[n], d);
    total_params += discovered[n]->size();
  }
  ComponentDict comps2;
  net->build(input, output, weights_table, comps2);
  loss.size = 0;
  sgd_dirty = true;
  invalidateGraphs();
}

// Captured step graphs bake in everything that is a kernel ARGUMENT rather than device data: the
// learning-rate decay, the gradient scale, the write-back flag, the graph topology (branches, fusion).
// Whoever changes one of those drops the graphs; the next steps re-capture.
void SupervisedTrainer::invalidateGraphs() {
  if (graphs.empty()) return;
  b200_sync(ctx);
  for (auto &kv : graphs) delete kv.second;
  graphs.clear();
}

void SupervisedTrainer::setOption(const std::string &name, double v) {
  optimizer.setOption(name, v);
  sgd_dirty = true;
  // the per-tensor hyper-parameters live in the device tables (re-uploaded before the next step), but
  // SGD's global `decay` is passed to the update kernels by value
  if (name == "decay" && optimizer.kind == B200_OPT_SGD) invalidateGraphs();
}
void SupervisedTrainer::setOptimizer(int kind) {
  optimizer.setKind(kind);
  sgd_dirty = true;
  invalidateGraphs();
  check(b200_memset_zero(ctx, count_dev, 2 * sizeof(int64_t)));
  if (updates_arena) updates_arena->zeros();
  state1_arena.reset();
  state2_arena.reset();
  state1.clear();
  state2.clear();
}
int64_t SupervisedTrainer::getCount() {
  int64_t c = 0;
  check(b200_memcpy_d2h(ctx, &c, count_dev, sizeof(int64_t)));
  check(b200_sync(ctx));
  return c;
}
void SupervisedTrainer::setCount(int64_t c) {
  int64_t v[2] = {c, 0};
  check(b200_sync(ctx));
  check(b200_memcpy_h2d(ctx, count_dev, v, 2 * sizeof(int64_t)));
  check(b200_sync(ctx));
  optimizer.count = c;
}
void SupervisedTrainer::setLayerwiseOption(const std::string &pattern, const std::string &name, double v) {
  // supervised.lua:262-269: the Lua pattern is expanded over the weight names
  for (auto &n : weights_order)
    if (luaPatternMatch(pattern, n)) optimizer.setLayerwiseOption(n, name, v);
  sgd_dirty = true;
}

void SupervisedTrainer::randomizeWeights(MTRand &rnd, double inf, double sup, bool use_fanin, bool use_fanout,
                                         const std::string &name_match) {
  // supervised.lua:575-635 + connection.cc:77-91 (rnd_weight macro :37-46)
  if (weights_order.empty()) throw Error(B200_ERR_NOT_BUILT, "Execute build method before randomize_weights");
  for (auto &n : weights_order) {
    if (!name_match.empty() && !luaPatternMatch(name_match, n)) continue;
    MatrixPtr w = weights_table[n];
    double constant = 0;
    if (use_fanin) constant += w->dim(1);
    if (use_fanout) constant += w->dim(0);
    double cinf = inf, csup = sup;
    if (constant > 0) { cinf = inf / sqrt(constant); csup = sup / sqrt(constant); }
    double dinf = (double)(float)cinf, dsup = (double)(float)csup;
    const double nearzero = 1e-7;
    if (fabs(dinf) < nearzero) dinf = nearzero;
    if (fabs(dsup) < nearzero) dsup = -nearzero;
    const double range = dsup - dinf;
    std::vector<float> host(w->size());
    for (auto &v : host) {
      unsigned it = 0;
      do {
        v = (float)(rnd.rand(range) + dinf);
        ++it;
      } while (it < 1000 && fabs(v) < nearzero);
    }
    w->fromHost(host.data());
    check(b200_sync(ctx));
  }
}

void SupervisedTrainer::uploadSgdTable() {
  const int nt = (int)arena_order.size();
  sgd_host.resize(nt);
  for (int i = 0; i < nt; ++i) {
    const std::string &n = arena_order[i];
    b200_sgd_tensor &t = sgd_host[i];
    memset(&t, 0, sizeof(t));
    t.w = weights_table[n]->data;
    t.g = grads[n]->data;
    t.u = updates[n]->data;
    t.n = weights_table[n]->size();
    t.rows = weights_table[n]->dim(0);
    t.cols = weights_table[n]->cols();
    auto opt = [&](const char *o) { return optimizer.validOption(o) ? (float)optimizer.getOptionOf(n, o) : 0.0f; };
    t.lr = opt("learning_rate");
    t.momentum = opt("momentum");
    t.weight_decay = opt("weight_decay");
    t.l1_norm = opt("L1_norm");
    t.max_norm_penalty = opt("max_norm_penalty");
  }
  if (optimizer.kind != B200_OPT_SGD) {
    // adagrad / rmsprop / adadelta: arena-shaped state blocks, one generic table
    const size_t total = weights_arena->size();
    if (!state1_arena) {
      state1_arena = Matrix::create(ctx, std::vector<int>{(int)total});
      state1_arena->zeros();
      state2_arena = Matrix::create(ctx, std::vector<int>{(int)total});
      state2_arena->zeros();
      for (auto &n : weights_order) {
        const size_t off = (size_t)(weights_table[n]->data - weights_arena->data);
        state1[n] = Matrix::view(state1_arena, off, weights_table[n]->dims);
        state2[n] = Matrix::view(state2_arena, off, weights_table[n]->dims);
      }
    }
    opt_host.resize(nt);
    for (int i = 0; i < nt; ++i) {
      const std::string &n = arena_order[i];
      b200_opt_tensor &o = opt_host[i];
      memset(&o, 0, sizeof(o));
      o.w = sgd_host[i].w; o.g = sgd_host[i].g; o.u = sgd_host[i].u;
      o.s1 = state1[n]->data; o.s2 = state2[n]->data;
      o.n = sgd_host[i].n; o.rows = sgd_host[i].rows; o.cols = sgd_host[i].cols;
      o.lr = sgd_host[i].lr; o.momentum = sgd_host[i].momentum;
      o.decay = (float)optimizer.getOptionOf(n, "decay");
      o.epsilon = (float)optimizer.getOptionOf(n, "epsilon");
      o.weight_decay = sgd_host[i].weight_decay; o.max_norm_penalty = sgd_host[i].max_norm_penalty;
      o.write_back_grad = keep_gradients ? 1 : 0;
    }
    if (!opt_dev) {
      void *p;
      check(b200_malloc(ctx, &p, sizeof(b200_opt_tensor) * (size_t)std::max(nt, 1)));
      opt_dev = (b200_opt_tensor *)p;
    }
    check(b200_memcpy_h2d(ctx, opt_dev, opt_host.data(), sizeof(b200_opt_tensor) * (size_t)nt));
  }
  if (!sgd_dev) {
    void *p;
    check(b200_malloc(ctx, &p, sizeof(b200_sgd_tensor) * (size_t)std::max(nt, 1)));
    sgd_dev = (b200_sgd_tensor *)p;
  }
  check(b200_memcpy_h2d(ctx, sgd_dev, sgd_host.data(), sizeof(b200_sgd_tensor) * (size_t)nt));
  // heavy = written by a big contraction (dot_product with more than 16 outputs, convolution)
  tensor_heavy.assign(nt, 0);
  for (auto *c : net->flatComponents()) {
    if (!c->hasWeightsName()) continue;
    auto *d = dynamic_cast<DotProductANNComponent *>(c);
    const bool heavy = (d && d->getOutputSize() > 16) || dynamic_cast<ConvolutionANNComponent *>(c);
    if (!heavy) continue;
    for (int i = 0; i < nt; ++i)
      if (arena_order[i] == c->getWeightsName()) tensor_heavy[i] = 1;
  }
  sgd_light_host.clear();
  for (int i = 0; i < nt; ++i)
    if (!tensor_heavy[i]) sgd_light_host.push_back(sgd_host[i]);
  if (!sgd_light_dev) {
    void *p;
    check(b200_malloc(ctx, &p, sizeof(b200_sgd_tensor) * (size_t)std::max(nt, 1)));
    sgd_light_dev = (b200_sgd_tensor *)p;
  }
  if (!sgd_light_host.empty())
    check(b200_memcpy_h2d(ctx, sgd_light_dev, sgd_light_host.data(), sizeof(b200_sgd_tensor) * sgd_light_host.size()));
  check(b200_sync(ctx));  // the host tables may be rewritten before the copies would otherwise run
  sgd_dirty = false;
  // captured graphs read the hyper-parameters from the device table, but the max-norm launches
  // are part of the graph structure: re-capture when that set changes
  std::string sig;
  for (auto &t : sgd_host) sig += (t.max_norm_penalty > 0.0f) ? '1' : '0';
  for (auto &t : sgd_host) sig += (t.momentum > 0.0f) ? 'm' : '-';   // rmsprop's look-ahead launch exists only with momentum
  if (sig != sgd_signature) {
    invalidateGraphs();
    sgd_signature = sig;
  }
}

// The whole output layer in one launch (b200_output_layer_fused): h = the input of net->deferred_dot.
// Returns the logits; fills logp / loss rows / gradient and, when training, hands the data gradient of
// the layer below to the stack (precomputed_dx) with the same fusion decision doBackprop would take.
MatrixPtr SupervisedTrainer::outputLayerFused(const MatrixPtr &h, const MatrixPtr &t, bool training, MatrixPtr &logp,
                                              MatrixPtr &rows, MatrixPtr &grad) {
  DotProductANNComponent *dot = net->deferred_dot;
  BiasANNComponent *bias = net->deferred_bias;
  if (h->dims.size() < 2 || (unsigned)h->cols() != dot->getInputSize())
    throw Error(B200_ERR_BAD_ARG, "Incorrect input size [" + dot->getName() + "]");
  const int bunch = h->rows(), K = (int)dot->getInputSize(), N = (int)dot->getOutputSize();
  if (!t || t->size() != (size_t)bunch * N) throw Error(128, "Different token sizes found: input vs target");
  MatrixPtr logits = Matrix::create(ctx, std::vector<int>{bunch, N});
  if (!logp) logp = Matrix::create(ctx, logits->dims);
  rows = Matrix::create(ctx, std::vector<int>{bunch});
  if (training) grad = Matrix::create(ctx, logits->dims);
  // the layer below: an element-wise activation whose output is h -> its derivative is applied here
  const auto &flat = net->flatComponents();
  size_t idx = 0;
  while (idx < flat.size() && flat[idx] != dot) ++idx;
  ActivationFunctionANNComponent *pa = idx >= 1 ? dynamic_cast<ActivationFunctionANNComponent *>(flat[idx - 1]) : nullptr;
  if (pa && (!pa->elementwise() || !pa->output)) pa = nullptr;
  MatrixPtr dx;
  const bool want_dx = training && !(idx == 0 && net->skip_input_gradient) && (!pa || pa->output->data == h->data);
  if (want_dx) dx = Matrix::create(ctx, std::vector<int>{bunch, K});
  check(b200_output_layer_fused(ctx, bunch, N, K, h->data, K, dot->weights_matrix->data, K,
                                bias ? bias->bias_vector->data : nullptr, t->data, logits->data, logp->data, rows->data,
                                grad ? grad->data : nullptr, pa ? pa->act : B200_ACT_NONE, dx ? dx->data : nullptr, K));
  dot->input = h;
  if (bias) { bias->input.reset(); bias->output.reset(); }
  (bias ? (ANNComponent *)bias : (ANNComponent *)dot)->output = logits;
  if (dx) {
    net->precomputed_for = dot;
    net->precomputed_dx = dx;
  }
  return logits;
}

// One training step enqueued on the stream: supervised.lua:769-819 + optimizer_sgd.lua:50-100.
// Bucket plan of the fused replica-group update, in arena order (= the order in which the backward pass
// finishes gradients).  b200_dp_fused_update takes at most 32 tensors per launch and the tag block has 16
// bucket slots; when a net of very many small tensors cannot be cut that way the caller stays on the
// NCCL all-reduce path, which has no such limits.
// Buckets of the fused replica-group update over tensors in arena order: each closes once it holds `bucket_bytes`
// of gradients or 32 tensors (the kernel's shared tables); the threshold doubles until at most 16 buckets remain
// (the tag block).  False when no such plan exists (more than 16 * 32 tensors): the caller uses the NCCL path.
bool dp_bucket_plan(const std::vector<size_t> &tensor_bytes, size_t bucket_bytes, std::vector<std::pair<int, int>> *out) {
  const int nt = (int)tensor_bytes.size();
  size_t threshold = bucket_bytes ? bucket_bytes : 1;
  for (int attempt = 0; attempt < 32; ++attempt, threshold *= 2) {
    std::vector<std::pair<int, int>> plan;
    int lo = 0;
    size_t bytes = 0;
    for (int i = 0; i < nt; ++i) {
      bytes += tensor_bytes[i];
      if (bytes >= threshold || i + 1 - lo == 32 || i == nt - 1) {
        plan.emplace_back(lo, i + 1);
        lo = i + 1;
        bytes = 0;
      }
    }
    if (plan.size() <= 16) {
      if (out) *out = plan;
      return true;
    }
    if (nt > 16 * 32) break;
  }
  return false;
}

bool SupervisedTrainer::dpBucketPlan(std::vector<std::pair<int, int>> *out) {
  std::vector<size_t> bytes;
  for (auto &n : arena_order) bytes.push_back(grads[n]->size() * sizeof(float));
  return dp_bucket_plan(bytes, dp_bucket_bytes, out);
}

// The update of a step as ONE launch over every tensor once every gradient exists: the path of the
// optimizers other than SGD, and of SGD when the global gradient norm is clipped (the clip needs every
// gradient before the first update).  Replica groups all-reduce the whole gradient arena first.
void SupervisedTrainer::runUpdateSimple(double max_gradients_norm) {
  const int nt = (int)arena_order.size();
  check(b200_branch_join_all(ctx));
  if (dp_nranks > 1) {
    check(b200_allreduce_sum_async(ctx, grads_arena->data, grads_arena->size(), 0));
    check(b200_comm_wait(ctx, 0));
  }
  if (max_gradients_norm > 0.0) {
    if (!norm_dev) {
      void *p;
      check(b200_malloc(ctx, &p, 512));
      norm_dev = (float *)p;
    }
    check(b200_grad_clip(ctx, grads_arena->size(), grads_arena->data, (float)max_gradients_norm, norm_dev));
  }
  if (optimizer.kind == B200_OPT_SGD) {
    const int wb = keep_gradients ? B200_SGD_WRITE_BACK_GRAD : 0;
    check(b200_sgd_multi_tensor_ex(ctx, nt, sgd_dev, sgd_host.data(), optimizer.getOption("decay"), count_dev,
                                   wb | B200_SGD_INCREMENT_COUNT));
  } else {
    check(b200_optimizer_multi_tensor(ctx, optimizer.kind, nt, opt_dev, opt_host.data(), count_dev, 1));
  }
}

void SupervisedTrainer::runStep(const MatrixPtr &x, const MatrixPtr &t, int global_bunch, double max_gradients_norm) {
  if (loss.kind == LOSS_ZERO_ONE) throw Error(128, "NON DIFERENTIABLE LOSS FUNCTION");
  // rmsprop evaluates the gradient at the look-ahead point w - momentum*Eupdate (optimizer_rmsprop.lua:44-51)
  if (optimizer.kind == B200_OPT_RMSPROP)
    check(b200_optimizer_lookahead(ctx, (int)opt_host.size(), opt_dev, opt_host.data()));
  net->reset();
  for (auto &kv : grads) kv.second->fresh = true;
  {
    // Tensor-core mode: the gradients of the big contractions are zeroed now, on the weight-gradient branch and
    // under the forward pass, and the contractions ACCUMULATE into them (the reference's own order:
    // md.zeros(grads) then beta = 1 GEMMs, supervised.lua:792-794) -- with beta == 1 the tensor-core kernel adds
    // its tiles with TMA reduce-add stores and may split K without any exchange between CTAs.
    int mode = B200_MATH_FP32;
    check(b200_get_math_mode(ctx, &mode));
    net->zero_accumulate = zero_accumulate;
    if (zero_accumulate && mode == B200_MATH_TF32 && use_branches && net->fuse) {
      bool any = false;
      for (size_t i = 0; i < arena_order.size(); ++i) {
        if (!tensor_heavy[i]) continue;
        MatrixPtr g = grads[arena_order[i]];
        if (!any) check(b200_branch_begin(ctx, 2));
        any = true;
        g->zeros();
        g->fresh = false;
      }
      if (any) check(b200_branch_end(ctx));
    }
  }
  net->use_branches = use_branches;   // (the forward pass zeroes the data-gradient buffers on a branch as well)
  auto *last = dynamic_cast<ActivationFunctionANNComponent *>(net->lastComponent());
  const bool fused_loss =
      net->fuse && last && last->act == B200_ACT_LOG_SOFTMAX && loss.kind == LOSS_MULTI_CLASS_CROSS_ENTROPY;
  MatrixPtr out, rows, grad;
  net->skip_input_gradient = true;
  net->precomputed_for = nullptr;
  net->precomputed_dx.reset();
  if (fused_loss) {
    net->defer_last_actf = true;
    net->defer_output_layer = fuse_output_layer;
    MatrixPtr logits = net->doForward(x, true);
    net->defer_last_actf = false;
    net->defer_output_layer = false;
    if (net->deferred_dot) {
      logits = outputLayerFused(logits, t, true, out, rows, grad);
    } else {
      grad = Matrix::create(ctx, logits->dims);
      loss.fusedLogSoftmaxMCCE(logits, t, out, rows, grad);
    }
    last->input = logits;
    last->output = out;
    net->output = out;
    net->last_actf_backprop_is_identity = true;
  } else {
    out = net->doForward(x, true);
    rows = loss.computeLoss(out, t);
    grad = loss.computeGradient(out, t);
    net->last_actf_backprop_is_identity = false;
  }
  // backward pass with the weight gradients interleaved (reverse layer order).
  net->grad_bunch = smooth_gradients ? (float)global_bunch : 0.0f;
  net->prepareGradScales();
  const int nt = (int)arena_order.size();
  const double decay = optimizer.getOption("decay");
  const int wb = keep_gradients ? B200_SGD_WRITE_BACK_GRAD : 0;
  auto tensorIndex = [&](const std::string &wn) {
    for (int i = 0; i < nt; ++i)
      if (arena_order[i] == wn) return i;
    return -1;
  };
  // a tensor is final once the earliest (in forward order) component that shares it has contributed
  auto isFinalContribution = [&](ANNComponent *c) {
    for (auto *o : net->flatComponents()) {
      if (o == c) return true;
      if (o->hasWeightsName() && o->getWeightsName() == c->getWeightsName()) return false;
    }
    return true;
  };
  net->interleave_grads = &grads;
  std::vector<char> done(nt, 0);
  auto cleanup = [&]() {
    net->interleave_grads = nullptr;
    net->on_gradients_ready = nullptr;
    net->on_backprop_issued = nullptr;
    net->use_branches = false;
    net->last_actf_backprop_is_identity = false;
  };
  bool any_max_norm = false;
  for (auto &t : sgd_host) any_max_norm = any_max_norm || t.max_norm_penalty > 0.0f;
  if (optimizer.kind != B200_OPT_SGD || max_gradients_norm > 0.0) {
    net->use_branches = use_branches;
    if (use_branches) check(b200_branch_begin(ctx, 0));
    loss.accumLoss(rows);
    if (use_branches) check(b200_branch_end(ctx));
    try {
      net->doBackprop(grad);
      cleanup();
      runUpdateSimple(max_gradients_norm);
    } catch (...) {
      cleanup();
      b200_branch_join_all(ctx);
      throw;
    }
  } else if (dp_nranks > 1 && dp_fused && !keep_gradients && !any_max_norm && dpBucketPlan(nullptr)) {
    // Replica group over NVLink peer memory: one kernel per bucket reduces the peers' gradient shards, updates
    // this rank's shard and writes the new weights into every replica (b200_dp_fused_update).  A bucket is
    // launched once its gradients are final AND the data gradients that read its weights have been issued
    // (the peers will overwrite them), on branch 0 beside the contractions still to come; the last bucket
    // on the main stream after the join.  b200_dp_wait then holds the step until every shard has landed.
    // buckets are planned up front (dpBucketPlan): at least dp_bucket_bytes of gradients each, at most 32
    // tensors (the kernel's shared tables) and at most 16 buckets (the tag block)
    std::vector<std::pair<int, int>> plan;
    dpBucketPlan(&plan);
    int nextb = 0, nbuckets = 0;
    auto launchBucket = [&](int lo, int hi, bool last) {
      if (last) {
        // The last bucket waits for EVERY branch, the earlier buckets on branch 0 included.  (Letting it run beside
        // them was tried at N = 8 and deadlocked the replica group: the kernels spin on cross-GPU tags and rely on
        // every CTA of a launch being resident; two such launches plus a contraction on one device break that
        // guarantee, and a rank whose CTAs cannot all start stalls every peer.)
        check(b200_branch_join_all(ctx));
      } else if (use_branches) {
        check(b200_branch_begin(ctx, 0));
        check(b200_branch_wait(ctx, 1));
        check(b200_branch_wait(ctx, 2));
        int sms = 0;
        check(b200_sm_count(ctx, &sms));
        check(b200_set_sm_budget(ctx, sms - sms / 2));
      }
      check(b200_dp_fused_update(ctx, &dp_group, hi - lo, sgd_dev + lo, sgd_host.data() + lo, decay, count_dev, nbuckets));
      check(b200_set_sm_budget(ctx, 0));
      if (!last && use_branches) check(b200_branch_end(ctx));
      ++nbuckets;
    };
    auto flush = [&](bool force) {
      while (nextb < (int)plan.size()) {
        bool ready = true;
        for (int i = plan[nextb].first; i < plan[nextb].second; ++i) ready = ready && done[i];
        if (!ready) break;
        const bool last = nextb + 1 == (int)plan.size();
        if (last && !force) break;   // the last bucket goes on the main stream, after the join
        launchBucket(plan[nextb].first, plan[nextb].second, last);
        ++nextb;
      }
    };
    net->use_branches = use_branches;
    net->on_backprop_issued = [&](ANNComponent *c, int) {
      if (!isFinalContribution(c)) return;
      const int i = tensorIndex(c->getWeightsName());
      if (i >= 0) done[i] = 1;
      flush(false);
    };
    if (use_branches) check(b200_branch_begin(ctx, 0));
    loss.accumLoss(rows);
    if (use_branches) check(b200_branch_end(ctx));
    try {
      net->doBackprop(grad);
      cleanup();
      for (int i = 0; i < nt; ++i) done[i] = 1;   // tensors no component touched this step: zero gradient on every rank
      flush(true);
      check(b200_branch_join_all(ctx));
      check(b200_dp_wait(ctx, &dp_group, nbuckets, count_dev));
      check(b200_counter_increment(ctx, count_dev));
    } catch (...) {
      cleanup();
      b200_branch_join_all(ctx);
      throw;
    }
  } else if (dp_nranks > 1) {
    // Replica group: finished gradients are all-reduced bucket by bucket on the communication stream
    // while the rest of the backward pass runs, and each bucket is updated as soon as it has arrived.
    std::vector<std::pair<int, int>> buckets;   // [first, last) tensor indices in arena order
    int next = 0;
    auto flush = [&](bool force) {
      int hi = next;
      size_t bytes = 0;
      while (hi < nt && done[hi]) { bytes += grads[arena_order[hi]]->size() * sizeof(float); ++hi; }
      if (hi == next) return;
      if (!force && bytes < dp_bucket_bytes && hi < nt) return;
      if ((int)buckets.size() >= 15 && hi < nt) return;   // keep one slot for the tail
      float *base = grads[arena_order[next]]->data;
      const MatrixPtr &last = grads[arena_order[hi - 1]];
      const size_t count = (size_t)(last->data - base) + last->size();
      check(b200_allreduce_sum_async(ctx, base, count, (int)buckets.size()));
      buckets.emplace_back(next, hi);
      next = hi;
    };
    net->on_gradients_ready = [&](ANNComponent *c) {
      if (!isFinalContribution(c)) return;
      const int i = tensorIndex(c->getWeightsName());
      if (i >= 0) done[i] = 1;
      flush(false);
    };
    // same branch structure as the single-replica step: weight gradients on side branches (the all-reduce
    // of a bucket waits for every branch that may hold part of it), statistics on branch 0
    net->use_branches = use_branches;
    if (use_branches) check(b200_branch_begin(ctx, 0));
    loss.accumLoss(rows);
    if (use_branches) check(b200_branch_end(ctx));
    try {
      net->doBackprop(grad);
      cleanup();
      for (int i = 0; i < nt; ++i) done[i] = 1;   // tensors no component touched this step stay zero: reduce them too
      flush(true);
      // every bucket but the last is updated on branch 0 as soon as it has arrived, beside the contractions
      // still running; the last one on the main stream after the join, and it bumps the step counter
      for (size_t b = 0; b < buckets.size(); ++b) {
        const bool last = (b + 1 == buckets.size());
        if (last) check(b200_branch_join_all(ctx));
        else if (use_branches) check(b200_branch_begin(ctx, 0));
        check(b200_comm_wait(ctx, (int)b));
        check(b200_sgd_multi_tensor_ex(ctx, buckets[b].second - buckets[b].first, sgd_dev + buckets[b].first,
                                       sgd_host.data() + buckets[b].first, decay, count_dev,
                                       wb | (last ? B200_SGD_INCREMENT_COUNT : 0)));
        if (!last && use_branches) check(b200_branch_end(ctx));
      }
      if (buckets.empty()) check(b200_branch_join_all(ctx));
    } catch (...) {
      cleanup();
      b200_branch_join_all(ctx);
      throw;
    }
  } else {
    // Single replica: the loss statistics, the weight gradients and the updates run on side branches;
    // only the data gradients stay on the critical path.  The tensor of a big contraction is updated by
    // its own launch as soon as its gradient is final AND the data gradient that reads the weights has
    // been issued (on_backprop_issued): on branch 0, beside the next contraction -- except the last one of
    // the step, which stays on the branch of its gradient kernel.  All the small tensors (biases, output
    // layer) take one launch at the end of branch 1, behind their gradient kernels.
    if (use_branches) check(b200_branch_begin(ctx, 0));
    loss.accumLoss(rows);
    if (use_branches) check(b200_branch_end(ctx));
    int last_heavy = -1;
    ANNComponent *first_heavy = nullptr;   // its tensor is the last one of the backward pass
    for (auto *c : net->flatComponents()) {
      if (!c->hasWeightsName()) continue;
      const int i = tensorIndex(c->getWeightsName());
      if (i >= 0 && tensor_heavy[i]) { first_heavy = c; break; }
    }
    net->use_branches = use_branches;
    net->on_backprop_issued = [&](ANNComponent *c, int branch) {
      if (!isFinalContribution(c)) return;
      const int i = tensorIndex(c->getWeightsName());
      if (i < 0 || done[i] || !tensor_heavy[i] || !sgd_as_ready) return;
      if (c == first_heavy) {
        // the last tensor of the backward pass: updated on the main stream after the join (below), where its
        // launch can also bump the step counter
        last_heavy = i;
        return;
      }
      if (use_branches) {
        const int b = (c == first_heavy && branch > 0) ? branch : 0;
        check(b200_branch_begin(ctx, b));
        if (branch > 0 && branch != b) check(b200_branch_wait(ctx, branch));
      }
      // an update that will run beside the next weight-gradient contraction (which plans for half of the
      // SMs, see StackANNComponent::doBackprop) takes the other half; the last one has the device to itself
      if (use_branches && c != first_heavy && net->contraction_mode == 1) {
        int sms = 0;
        check(b200_sm_count(ctx, &sms));
        const long tiles = (long)((sgd_host[i].rows + 127) / 128) * (long)((sgd_host[i].cols + 255) / 256);
        if (tiles <= sms) check(b200_set_sm_budget(ctx, sms - sms / 2));   // same small-layer regime as the contractions
      }
      check(b200_sgd_multi_tensor_ex(ctx, 1, sgd_dev + i, sgd_host.data() + i, decay, count_dev, wb));
      check(b200_set_sm_budget(ctx, 0));
      if (use_branches) check(b200_branch_end(ctx));
      done[i] = 1;
    };
    try {
      net->doBackprop(grad);
      // heavy tensors no component touched this step (zero gradient) still take their momentum / decay step
      if (!sgd_as_ready) {
        // everything in one launch once every gradient exists
        check(b200_branch_join_all(ctx));
        check(b200_sgd_multi_tensor_ex(ctx, nt, sgd_dev, sgd_host.data(), decay, count_dev, wb | B200_SGD_INCREMENT_COUNT));
      }
      for (int i = 0; i < nt && sgd_as_ready; ++i) {
        if (done[i] || !tensor_heavy[i] || i == last_heavy) continue;
        check(b200_sgd_multi_tensor_ex(ctx, 1, sgd_dev + i, sgd_host.data() + i, decay, count_dev, wb));
        done[i] = 1;
      }
      if (sgd_as_ready && !sgd_light_host.empty()) {
        if (use_branches) check(b200_branch_begin(ctx, 1));
        check(b200_sgd_multi_tensor_ex(ctx, (int)sgd_light_host.size(), sgd_light_dev, sgd_light_host.data(), decay,
                                       count_dev, wb));
        if (use_branches) check(b200_branch_end(ctx));
      }
      check(b200_branch_join_all(ctx));
      // the updates ran on several streams: the counter is bumped once they are joined -- by the last big
      // tensor's update (its last CTA to finish does it) when there is one, else by a launch of its own
      if (sgd_as_ready) {
        if (last_heavy >= 0) {
          check(b200_sgd_multi_tensor_ex(ctx, 1, sgd_dev + last_heavy, sgd_host.data() + last_heavy, decay, count_dev,
                                         wb | B200_SGD_INCREMENT_COUNT));
          done[last_heavy] = 1;
        } else {
          check(b200_counter_increment(ctx, count_dev));
        }
      }
    } catch (...) {
      cleanup();
      b200_branch_join_all(ctx);
      throw;
    }
    cleanup();
  }
  last_loss_rows = rows;
  last_output = out;
}

void SupervisedTrainer::trainStepDevice(const MatrixPtr &x, const MatrixPtr &t, int smoothing_bunch,
                                        double max_gradients_norm) {
  if (weights_order.empty()) throw Error(B200_ERR_NOT_BUILT, "Execute build method before call this method");
  check(b200_make_current(ctx));
  if (sgd_dirty) {
    // hyper-parameters are baked into the SGD table; captured graphs read it from the device
    uploadSgdTable();
  }
  const int bunch = x->rows();
  // supervised.lua:757,800: the smoothing factor uses `bunch_size or self.bunch_size or 1`, NOT the row count
  // of the bunch (train_dataset passes the actual length of every bunch, supervised.lua:1203)
  if (smoothing_bunch <= 0) smoothing_bunch = bunch_size > 0 ? bunch_size : 1;
  const int global_bunch = smoothing_bunch * dp_nranks;
  if (max_gradients_norm != graph_max_norm) {   // the clip threshold is a kernel argument of the captured step
    invalidateGraphs();
    graph_max_norm = max_gradients_norm;
  }
  cudaStream_t stream = (cudaStream_t)b200_stream(ctx);
  Graph *g = nullptr;
  if (use_cuda_graph) {
    const auto key = std::make_pair(std::make_pair(bunch, smoothing_bunch), (const float *)x->data);
    auto it = graphs.find(key);
    if (it == graphs.end()) {
      if (graphs.size() >= 16) {   // callers that feed ever-changing buffers: do not hoard graphs
        b200_sync(ctx);
        for (auto &kv : graphs) delete kv.second;
        graphs.clear();
      }
      g = new Graph();
      graphs[key] = g;
    } else {
      g = it->second;
    }
  }
  if (g && g->exec && g->x_ptr == x->data && g->t_ptr == t->data) {
    cudaCheck(cudaGraphLaunch(g->exec, stream), "cudaGraphLaunch");
    b200_add_launches(ctx, g->launches);
    last_loss_rows = g->rows;
    last_output = g->out;
    optimizer.count++;
    return;
  }
  if (g && g->warm >= 1 && !g->exec) {
    // second step of this bunch size: the pool and scratch are warm, capture it
    uint64_t before = 0, after = 0;
    b200_launch_count(ctx, &before);
    g_capture_registry = &g->keep;
    cudaCheck(cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed), "cudaStreamBeginCapture");
    try {
      runStep(x, t, global_bunch, max_gradients_norm);
    } catch (...) {
      g_capture_registry = nullptr;
      cudaGraph_t dead = nullptr;
      cudaStreamEndCapture(stream, &dead);
      if (dead) cudaGraphDestroy(dead);
      throw;
    }
    g_capture_registry = nullptr;
    cudaCheck(cudaStreamEndCapture(stream, &g->graph), "cudaStreamEndCapture");
    cudaCheck(cudaGraphInstantiate(&g->exec, g->graph, 0), "cudaGraphInstantiate");
    b200_launch_count(ctx, &after);
    g->launches = after - before;
    b200_add_launches(ctx, (uint64_t)0 - g->launches);  // capture enqueued nothing yet
    g->rows = last_loss_rows;
    g->out = last_output;
    g->x_ptr = x->data;
    g->t_ptr = t->data;
    cudaCheck(cudaGraphLaunch(g->exec, stream), "cudaGraphLaunch");
    b200_add_launches(ctx, g->launches);
    optimizer.count++;
    return;
  }
  runStep(x, t, global_bunch, max_gradients_norm);
  if (g) g->warm++;
  optimizer.count++;
}

void SupervisedTrainer::validateStepDevice(const MatrixPtr &x, const MatrixPtr &t) {
  // supervised.lua:825-862
  net->reset();
  auto *last = dynamic_cast<ActivationFunctionANNComponent *>(net->lastComponent());
  const bool fused_loss =
      net->fuse && last && last->act == B200_ACT_LOG_SOFTMAX && loss.kind == LOSS_MULTI_CLASS_CROSS_ENTROPY;
  MatrixPtr out, rows, nograd;
  if (fused_loss) {
    net->defer_last_actf = true;
    net->defer_output_layer = fuse_output_layer;
    MatrixPtr logits = net->doForward(x, false);
    net->defer_last_actf = false;
    net->defer_output_layer = false;
    if (net->deferred_dot) logits = outputLayerFused(logits, t, false, out, rows, nograd);
    else loss.fusedLogSoftmaxMCCE(logits, t, out, rows, nograd);
    last->input = logits;
    last->output = out;
    net->output = out;
  } else {
    out = net->doForward(x, false);
    rows = loss.computeLoss(out, t);
  }
  loss.accumLoss(rows);
  last_loss_rows = rows;
  last_output = out;
}

MatrixPtr SupervisedTrainer::calculate(const MatrixPtr &x) {
  net->reset();
  return net->doForward(x, false);
}

static MatrixPtr stageView(b200_ctx *ctx, MatrixPtr &stage, int rows, int cols, int min_rows) {
  const size_t need = (size_t)std::max(rows, min_rows) * cols;
  if (need > (size_t)INT32_MAX) throw Error(B200_ERR_UNSUPPORTED, "bunch of more than 2^31-1 values");
  if (!stage || stage->size() < need) stage = Matrix::create(ctx, std::vector<int>{(int)need});
  return Matrix::view(stage, 0, std::vector<int>{rows, cols});
}

float SupervisedTrainer::trainStep(const float *x, const float *t, int bunch, float *loss_rows_out, int smoothing_bunch,
                                   double max_gradients_norm) {
  const int in = (int)net->getInputSize(), out = (int)net->getOutputSize();
  if (in <= 0 || out <= 0) throw Error(B200_ERR_NOT_BUILT, "Execute build method before call this method");
  MatrixPtr sx = stageView(ctx, stage_x, bunch, in, bunch_size);
  MatrixPtr st = stageView(ctx, stage_t, bunch, out, bunch_size);
  sx->fromHost(x);
  st->fromHost(t);
  trainStepDevice(sx, st, smoothing_bunch, max_gradients_norm);
  std::vector<float> rows(bunch);
  last_loss_rows->toHost(rows.data());
  if (loss_rows_out) memcpy(loss_rows_out, rows.data(), sizeof(float) * bunch);
  float s = 0.0f;  // matSum(loss)/dim  (bind_loss_functions.lua.cc:71)
  for (float v : rows) s += v;
  return s / (float)bunch;
}
float SupervisedTrainer::validateStep(const float *x, const float *t, int bunch, float *loss_rows_out) {
  const int in = (int)net->getInputSize(), out = (int)net->getOutputSize();
  MatrixPtr sx = stageView(ctx, stage_x, bunch, in, bunch_size);
  MatrixPtr st = stageView(ctx, stage_t, bunch, out, bunch_size);
  sx->fromHost(x);
  st->fromHost(t);
  validateStepDevice(sx, st);
  std::vector<float> rows(bunch);
  last_loss_rows->toHost(rows.data());
  if (loss_rows_out) memcpy(loss_rows_out, rows.data(), sizeof(float) * bunch);
  float s = 0.0f;
  for (float v : rows) s += v;
  return s / (float)bunch;
}

void SupervisedTrainer::trainDataset(const float *x, const float *t, int n, const int *order, float *mean,
                                     float *var) {
  // supervised.lua:1149-1226: loss:reset(), iterate bunches in the given order (the shuffle is
  // drawn by the caller's random object, trainable.lua:217-220), return loss:get_accum_loss()
  const int in = (int)net->getInputSize(), out = (int)net->getOutputSize();
  if (in <= 0 || out <= 0) throw Error(B200_ERR_NOT_BUILT, "Execute build method before call this method");
  MatrixPtr dx = Matrix::create(ctx, std::vector<int>{n, in});
  MatrixPtr dt = Matrix::create(ctx, std::vector<int>{n, out});
  dx->fromHost(x);
  dt->fromHost(t);
  std::vector<int32_t> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = order ? order[i] : i;
  void *p;
  check(b200_malloc(ctx, &p, sizeof(int32_t) * (size_t)n));
  int32_t *idx_dev = (int32_t *)p;
  check(b200_memcpy_h2d(ctx, idx_dev, idx.data(), sizeof(int32_t) * (size_t)n));
  loss.reset();
  for (int k = 0; k < n; k += bunch_size) {
    const int b = std::min(bunch_size, n - k);
    MatrixPtr sx = stageView(ctx, stage_x, b, in, bunch_size);
    MatrixPtr st = stageView(ctx, stage_t, b, out, bunch_size);
    check(b200_gather_rows(ctx, b, in, dx->data, idx_dev + k, sx->data));
    check(b200_gather_rows(ctx, b, out, dt->data, idx_dev + k, st->data));
    trainStepDevice(sx, st, b);   // #bunch_indexes, supervised.lua:1203
  }
  loss.getAccumLoss(mean, var);  // also synchronises: idx / dataset buffers are idle afterwards
  check(b200_free(ctx, idx_dev));
}

void SupervisedTrainer::validateDataset(const float *x, const float *t, int n, float *mean, float *var) {
  const int in = (int)net->getInputSize(), out = (int)net->getOutputSize();
  MatrixPtr dx = Matrix::create(ctx, std::vector<int>{n, in});
  MatrixPtr dt = Matrix::create(ctx, std::vector<int>{n, out});
  dx->fromHost(x);
  dt->fromHost(t);
  loss.reset();
  for (int k = 0; k < n; k += bunch_size) {
    const int b = std::min(bunch_size, n - k);
    validateStepDevice(Matrix::view(dx, (size_t)k * in, std::vector<int>{b, in}),
                       Matrix::view(dt, (size_t)k * out, std::vector<int>{b, out}));
  }
  loss.getAccumLoss(mean, var);
}

void SupervisedTrainer::useDataset(const float *x, int n, float *y) {
  // supervised.lua:1291-1430: forward (not training) bunch by bunch, outputs collected on the host
  const int in = (int)net->getInputSize(), out = (int)net->getOutputSize();
  if (in <= 0 || out <= 0) throw Error(B200_ERR_NOT_BUILT, "Execute build method before call this method");
  MatrixPtr dx = Matrix::create(ctx, std::vector<int>{n, in});
  dx->fromHost(x);
  MatrixPtr dy = Matrix::create(ctx, std::vector<int>{n, out});
  for (int k = 0; k < n; k += bunch_size) {
    const int b = std::min(bunch_size, n - k);
    MatrixPtr o = calculate(Matrix::view(dx, (size_t)k * in, std::vector<int>{b, in}));
    if ((int)o->size() != b * out) throw Error(B200_ERR_BAD_ARG, "use_dataset: unexpected output size");
    check(b200_memcpy_d2d(ctx, dy->data + (size_t)k * out, o->data, sizeof(float) * (size_t)b * out));
  }
  dy->toHost(y);
}

double SupervisedTrainer::norm2(const std::string &pattern) {
  // supervised.lua:1556-1576: max over matching matrices of the largest row 2-norm
  double best = 0.0;
  for (auto &n : weights_order) {
    if (!luaPatternMatch(pattern, n)) continue;
    MatrixPtr w = weights_table[n];
    std::vector<float> h(w->size());
    w->toHost(h.data());
    const int rows = w->dim(0), cols = w->cols();
    for (int r = 0; r < rows; ++r) {
      double s = 0;
      for (int c = 0; c < cols; ++c) s += (double)h[(size_t)r * cols + c] * h[(size_t)r * cols + c];
      best = std::max(best, sqrt(s));
    }
  }
  return best;
}

void SupervisedTrainer::dpExport(int nranks, unsigned char *handles) {
  if (weights_order.empty()) throw Error(B200_ERR_NOT_BUILT, "Execute build method before joining a replica group");
  if (nranks < 2 || nranks > B200_DP_MAX_RANKS) throw Error(B200_ERR_BAD_ARG, "replica group of 2..8 ranks");
  if (!dp_flags) {
    void *p;
    check(b200_malloc(ctx, &p, b200_dp_flags_bytes()));
    dp_flags = (long long *)p;
    check(b200_memset_zero(ctx, dp_flags, b200_dp_flags_bytes()));
    // receive block: one arena-shaped slot per peer (only this rank's shards of it are ever written)
    const size_t bytes = (size_t)(nranks - 1) * weights_arena->size() * sizeof(float);
    check(b200_malloc(ctx, &p, bytes));
    dp_recv = (float *)p;
    check(b200_memset_zero(ctx, dp_recv, bytes));
    check(b200_sync(ctx));
  }
  check(b200_ipc_export(ctx, weights_arena->data, handles));
  check(b200_ipc_export(ctx, grads_arena->data, handles + 64));
  check(b200_ipc_export(ctx, dp_flags, handles + 128));
  check(b200_ipc_export(ctx, dp_recv, handles + 192));
}
void SupervisedTrainer::dpConnect(int nranks, int rank, const unsigned char *all) {
  if (nranks < 2 || nranks > B200_DP_MAX_RANKS) throw Error(B200_ERR_BAD_ARG, "replica group of 2..8 ranks");
  if (!dp_flags) throw Error(B200_ERR_BAD_ARG, "dpExport first");
  dp_group = b200_dp_group{};
  dp_group.nranks = nranks;
  dp_group.rank = rank;
  dp_group.arena_elems = weights_arena->size();
  for (int p = 0; p < nranks; ++p) {
    if (p == rank) {
      dp_group.weights[p] = weights_arena->data;
      dp_group.grads[p] = grads_arena->data;
      dp_group.flags[p] = dp_flags;
      dp_group.recv[p] = dp_recv;
      continue;
    }
    void *ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; ++k) {
      check(b200_ipc_import(ctx, all + (size_t)p * 256 + 64 * k, &ptr[k]));
      dp_imported.push_back(ptr[k]);
    }
    dp_group.weights[p] = (float *)ptr[0];
    dp_group.grads[p] = (float *)ptr[1];
    dp_group.flags[p] = (long long *)ptr[2];
    dp_group.recv[p] = (float *)ptr[3];
  }
  dp_fused = true;
  if (const char *e = getenv("B200_DP_FUSED")) dp_fused = atoi(e) != 0;
  invalidateGraphs();
}

// Moves the weight and gradient arenas into caller-provided device memory (same layout), e.g. a buffer that is
// mapped into every process of a replica group.  Every view and every component is re-bound.
void SupervisedTrainer::rehomeArenas(float *w, float *g) {
  if (weights_order.empty()) throw Error(B200_ERR_NOT_BUILT, "Execute build method before joining a replica group");
  check(b200_sync(ctx));
  invalidateGraphs();
  const int total = (int)weights_arena->size();
  MatrixPtr nw = Matrix::wrap(ctx, w, std::vector<int>{total}), ng = Matrix::wrap(ctx, g, std::vector<int>{total});
  nw->copyFrom(*weights_arena);
  ng->copyFrom(*grads_arena);
  check(b200_sync(ctx));
  std::map<std::string, size_t> offs;
  for (auto &n : weights_order) offs[n] = (size_t)(weights_table[n]->data - weights_arena->data);
  weights_arena = nw;
  grads_arena = ng;
  for (auto &n : weights_order) {
    const std::vector<int> d = weights_table[n]->dims;
    weights_table[n] = Matrix::view(weights_arena, offs[n], d);
    grads[n] = Matrix::view(grads_arena, offs[n], d);
  }
  ComponentDict comps;
  net->build(build_in, build_out, weights_table, comps);
  sgd_dirty = true;
}

size_t SupervisedTrainer::dpSymmetricBytes(size_t arena_floats) {
  const size_t a = (arena_floats * sizeof(float) + 1023) & ~size_t(1023);
  return 2 * a + ((b200_dp_flags_bytes() + 1023) & ~size_t(1023));
}

void SupervisedTrainer::dpConnectSymmetric(int nranks, int rank, void *const *bases, void *mc_base, size_t bytes) {
  if (nranks < 2 || nranks > B200_DP_MAX_RANKS) throw Error(B200_ERR_BAD_ARG, "replica group of 2..8 ranks");
  const size_t floats = weights_arena->size();
  const size_t a = (floats * sizeof(float) + 1023) & ~size_t(1023);
  if (bytes < dpSymmetricBytes(floats)) throw Error(B200_ERR_BAD_ARG, "symmetric buffer too small (b200h_trainer_dp_symmetric_bytes)");
  for (int p = 0; p < nranks; ++p)
    if (!bases[p]) throw Error(B200_ERR_BAD_ARG, "dpConnectSymmetric: NULL base pointer");
  char *mine = (char *)bases[rank];
  rehomeArenas((float *)mine, (float *)(mine + a));
  check(b200_memset_zero(ctx, mine + 2 * a, b200_dp_flags_bytes()));
  check(b200_sync(ctx));
  if (dp_flags && dp_flags_owned) b200_free(ctx, dp_flags);
  dp_flags = (long long *)(mine + 2 * a);
  dp_flags_owned = false;
  dp_group = b200_dp_group{};
  dp_group.nranks = nranks;
  dp_group.rank = rank;
  dp_group.arena_elems = floats;
  for (int p = 0; p < nranks; ++p) {
    char *b = (char *)bases[p];
    dp_group.weights[p] = (float *)b;
    dp_group.grads[p] = (float *)(b + a);
    dp_group.flags[p] = (long long *)(b + 2 * a);
    dp_group.recv[p] = nullptr;    // (no receive blocks: the copy-engine / push variants are not available here)
  }
  dp_group.mc_weights = mc_base ? (float *)mc_base : nullptr;
  dp_group.mc_grads = mc_base ? (float *)((char *)mc_base + a) : nullptr;
  if (const char *e = getenv("B200_DP_MULTICAST"))
    if (atoi(e) == 0) dp_group.mc_weights = dp_group.mc_grads = nullptr;   // same memory, P2P pull kernel
  dp_fused = true;
  if (const char *e = getenv("B200_DP_FUSED")) dp_fused = atoi(e) != 0;
  invalidateGraphs();
}

float SupervisedTrainer::dpBench(int reps) {
  if (!dp_fused) throw Error(B200_ERR_BAD_ARG, "dpConnect first");
  if (sgd_dirty) uploadSgdTable();
  const double decay = optimizer.getOption("decay");
  cudaEvent_t e0, e1;
  cudaCheck(cudaEventCreate(&e0), "cudaEventCreate");
  cudaCheck(cudaEventCreate(&e1), "cudaEventCreate");
  cudaStream_t stream = (cudaStream_t)b200_stream(ctx);
  auto once = [&]() {
    check(b200_dp_fused_update(ctx, &dp_group, (int)sgd_host.size(), sgd_dev, sgd_host.data(), decay, count_dev, 0));
    check(b200_dp_wait(ctx, &dp_group, 1, count_dev));
    check(b200_counter_increment(ctx, count_dev));
  };
  for (int i = 0; i < 3; ++i) once();
  cudaCheck(cudaEventRecord(e0, stream), "cudaEventRecord");
  for (int i = 0; i < reps; ++i) once();
  cudaCheck(cudaEventRecord(e1, stream), "cudaEventRecord");
  cudaCheck(cudaEventSynchronize(e1), "cudaEventSynchronize");
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return ms * 1e3f / (float)reps;
}

void SupervisedTrainer::broadcastWeights() {
  check(b200_broadcast(ctx, weights_arena->data, weights_arena->size(), 0));
  check(b200_sync(ctx));
}

}  // namespace b200

namespace b200 {
// pipelined stepping used by train loops that keep the loss on the device: stage() copies the bunch into
// one of two staging slots on a copy stream (it only waits for the step that last trained on that slot),
// stepStaged() makes the compute stream wait for that copy and runs the step.  With the host running
// ahead, the copy of bunch k+1 overlaps the training of bunch k.
void SupervisedTrainer::stage(const float *x, const float *t, int bunch) {
  const int in = (int)net->getInputSize(), out = (int)net->getOutputSize();
  if (in <= 0 || out <= 0) throw Error(B200_ERR_NOT_BUILT, "Execute build method before call this method");
  check(b200_make_current(ctx));
  if (!copy_stream) {
    cudaStream_t cs;
    cudaCheck(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking), "cudaStreamCreate");
    copy_stream = cs;
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t a, b;
      cudaCheck(cudaEventCreateWithFlags(&a, cudaEventDisableTiming), "cudaEventCreate");
      cudaCheck(cudaEventCreateWithFlags(&b, cudaEventDisableTiming), "cudaEventCreate");
      ev_copied[i] = a;
      ev_trained[i] = b;
    }
  }
  const int s = next_slot;
  next_slot ^= 1;
  cudaStream_t cs = (cudaStream_t)copy_stream;
  MatrixPtr sx = stageView(ctx, pipe_x[s], bunch, in, bunch_size);
  MatrixPtr st = stageView(ctx, pipe_t[s], bunch, out, bunch_size);
  if (slot_trained[s]) cudaCheck(cudaStreamWaitEvent(cs, (cudaEvent_t)ev_trained[s], 0), "cudaStreamWaitEvent");
  cudaCheck(cudaMemcpyAsync(sx->data, x, sizeof(float) * (size_t)bunch * in, cudaMemcpyHostToDevice, cs), "H2D");
  cudaCheck(cudaMemcpyAsync(st->data, t, sizeof(float) * (size_t)bunch * out, cudaMemcpyHostToDevice, cs), "H2D");
  cudaCheck(cudaEventRecord((cudaEvent_t)ev_copied[s], cs), "cudaEventRecord");
  staged_slot = s;
}
void SupervisedTrainer::stepStaged(int bunch) {
  const int in = (int)net->getInputSize(), out = (int)net->getOutputSize();
  if (staged_slot < 0) throw Error(B200_ERR_BAD_ARG, "step_staged before stage");
  const int s = staged_slot;
  cudaStream_t stream = (cudaStream_t)b200_stream(ctx);
  cudaCheck(cudaStreamWaitEvent(stream, (cudaEvent_t)ev_copied[s], 0), "cudaStreamWaitEvent");
  trainStepDevice(stageView(ctx, pipe_x[s], bunch, in, bunch_size), stageView(ctx, pipe_t[s], bunch, out, bunch_size), bunch);
  cudaCheck(cudaEventRecord((cudaEvent_t)ev_trained[s], stream), "cudaEventRecord");
  slot_trained[s] = true;
}
}  // namespace b200
