// Host-side C++ mirror of the reference's plugin interfaces for the training hot path:
//   ANN::ANNComponent            packages/ann/ann/c_src/ann_component.h:95-606
//   ANN::LossFunction            packages/ann/loss/c_src/loss_function.h:34-120
//   ann.optimizer.sgd            packages/ann/optimizer/lua_src/optimizer_sgd.lua
//   trainable.supervised_trainer packages/trainable/lua_src/supervised.lua
//   Basics::MTRand               packages/basics/random/c_src/MersenneTwister.h
// Same method names, argument meaning and error behaviour; everything below a Matrix is
// a call into the C ABI of include/b200ann.h (no CPU fallback: without a device,
// construction of a Context fails).
#pragma once
#include <stdint.h>

#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/b200ann.h"

namespace b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
void check(int status);  // throws Error(status, b200_last_error_string())
// bucket plan of the fused replica-group update (trainer.cc); pure host logic
bool dp_bucket_plan(const std::vector<size_t> &tensor_bytes, size_t bucket_bytes, std::vector<std::pair<int, int>> *out);

// ---------------------------------------------------------------- MTRand
class MTRand {
 public:
  explicit MTRand(uint32_t seed) { this->seed(seed); }
  void seed(uint32_t s);
  uint32_t randInt();
  uint32_t randInt(uint32_t n);          // [0, n]
  double rand(double n = 1.0);           // [0, n]
  void shuffle(int size, int *out);      // 0-based permutation
  // 624 state words (after the last reload) + index of the next unread word: what the device-side
  // generator of the dropout mask continues from
  void exportState(uint32_t *words624, int32_t *next) const {
    for (int i = 0; i < 624; ++i) words624[i] = state[i];
    *next = left == 0 ? 624 : pos;
  }
 private:
  void reload();
  uint32_t state[624];
  int left = 0, pos = 0;
};

// ---------------------------------------------------------------- Matrix
// Device-resident float32 tensor, contiguous row-major.  Mirrors the part of
// Basics::Matrix<float> (packages/basics/matrix/c_src/matrix.h:77) the hot path uses.
class Matrix;
using MatrixPtr = std::shared_ptr<Matrix>;
class Matrix : public std::enable_shared_from_this<Matrix> {
 public:
  b200_ctx *ctx;
  std::vector<int> dims;
  float *data = nullptr;
  bool owns = false;        // allocated from the pool (freed in the destructor)
  MatrixPtr parent;         // keeps a viewed block alive
  int shared_count = 0;     // matrix.h:637-643
  bool fresh = true;        // gradient accumulators: nothing written yet this step

  static MatrixPtr create(b200_ctx *ctx, const std::vector<int> &dims);
  static MatrixPtr view(const MatrixPtr &parent, size_t offset, const std::vector<int> &dims);
  static MatrixPtr wrap(b200_ctx *ctx, float *ptr, const std::vector<int> &dims);
  ~Matrix();
  size_t size() const;
  int dim(int i) const { return dims.at(i); }
  int rows() const { return dims.at(0); }
  int cols() const { return (int)(size() / (size_t)dims.at(0)); }
  MatrixPtr rewrap(const std::vector<int> &new_dims);   // metadata-only reshape
  void zeros();
  void fromHost(const float *src);      // async H2D on the context stream
  void toHost(float *dst) const;        // D2H + sync
  void copyFrom(const Matrix &o);
};

using MatrixDict = std::map<std::string, MatrixPtr>;
void registerCapturedMatrix(const MatrixPtr &m);   // keeps step activations alive for captured graphs

// ---------------------------------------------------------------- components
class ANNComponent;
using ComponentPtr = std::shared_ptr<ANNComponent>;
using ComponentDict = std::map<std::string, ANNComponent *>;

class ANNComponent {
 public:
  ANNComponent(const std::string &name, const std::string &weights_name, unsigned in, unsigned out)
      : name(name), weights_name(weights_name), input_size(in), output_size(out) {}
  virtual ~ANNComponent() {}
  // ann_component.h: doForward / doBackprop / computeAllGradients / reset / build
  virtual MatrixPtr doForward(const MatrixPtr &input, bool during_training) = 0;
  virtual MatrixPtr doBackprop(const MatrixPtr &error_input) = 0;
  virtual void computeAllGradients(MatrixDict &grads) {}
  virtual void reset(unsigned it = 0);
  virtual void build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components);
  virtual const char *kind() const = 0;
  virtual bool hasWeightsName() const { return !weights_name.empty(); }
  virtual void setContext(b200_ctx *c) { ctx = c; }

  const std::string &getName() const { return name; }
  const std::string &getWeightsName() const { return weights_name; }
  unsigned getInputSize() const { return input_size; }
  unsigned getOutputSize() const { return output_size; }
  MatrixPtr getInput() const { return input; }
  MatrixPtr getOutput() const { return output; }
  MatrixPtr getErrorInput() const { return error_input; }
  MatrixPtr getErrorOutput() const { return error_output; }

  // 1/sqrt(shared_count*bunch) is folded into the gradient kernels; the trainer sets
  // the bunch factor before computeAllGradients (supervised.lua:797-803).
  float grad_bunch = 0.0f;   // 0 => unscaled sums (raw component API)
  float grad_scale = 1.0f;   // set by the enclosing stack: 1/sqrt(total shared count * grad_bunch)
  // uses of the weights this component adds per step (dot_product_component.cc:177 adds 1,
  // convolution_component.cc:302 adds the number of output pixels)
  virtual int sharedCountContribution() const { return 0; }

  std::string name, weights_name;
  unsigned input_size, output_size;
  b200_ctx *ctx = nullptr;
  MatrixPtr input, output, error_input, error_output;
};

class DotProductANNComponent : public ANNComponent {
 public:
  DotProductANNComponent(const std::string &name, const std::string &wname, unsigned in, unsigned out);
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  void computeAllGradients(MatrixDict &grads) override;
  void reset(unsigned it = 0) override;
  void build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) override;
  const char *kind() const override { return "dot_product"; }
  int sharedCountContribution() const override { return 1; }
  MatrixPtr weights_matrix;
  // tensor-core mode: the data gradient of this layer is ADDED into a buffer that was allocated and zeroed during
  // the forward pass (on a side branch), so that the contraction can split K without an exchange
  MatrixPtr dx_zeroed;
};

class BiasANNComponent : public ANNComponent {
 public:
  BiasANNComponent(const std::string &name, const std::string &wname, unsigned size);
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  void computeAllGradients(MatrixDict &grads) override;
  void reset(unsigned it = 0) override;
  void build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) override;
  const char *kind() const override { return "bias"; }
  int sharedCountContribution() const override { return 1; }
  MatrixPtr bias_vector;
};

class ActivationFunctionANNComponent : public ANNComponent {
 public:
  ActivationFunctionANNComponent(const std::string &name, int act, float p0 = 0.0f, float p1 = 0.0f);
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  const char *kind() const override { return "actf"; }
  void build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) override;
  // fusable into the epilogue of the contraction that feeds it
  bool elementwise() const {
    return act == B200_ACT_LOGISTIC || act == B200_ACT_TANH || act == B200_ACT_RELU || act == B200_ACT_LINEAR;
  }
  bool rowwise() const { return act == B200_ACT_SOFTMAX || act == B200_ACT_LOG_SOFTMAX; }
  int act;
  float p0, p1;   // leaky_relu: leak ; hardtanh: inf, sup
};

// prelu_actf_component.cc: y = x>0 ? x : a*x with a learnable a[size,1] (scalar: [1,1])
class PReLUActfANNComponent : public ANNComponent {
 public:
  PReLUActfANNComponent(const std::string &name, const std::string &wname, unsigned size, bool scalar);
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  void computeAllGradients(MatrixDict &grads) override;
  void reset(unsigned it = 0) override;
  void build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) override;
  const char *kind() const override { return "prelu"; }
  int sharedCountContribution() const override { return 1; }
  bool scalar;
  MatrixPtr weights_matrix;
};

// dropout_component.cc:67-134.  The mask is drawn on the device from the component's own MT19937 stream
// (the state of the `random` object given at construction is copied at the first training forward).
class DropoutANNComponent : public ANNComponent {
 public:
  DropoutANNComponent(const std::string &name, const MTRand &random, float prob, float value, bool norm, unsigned size);
  ~DropoutANNComponent();
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  void build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) override;
  const char *kind() const override { return "dropout"; }
  MTRand random;
  float prob, value;
  bool normalize_after_training;
  void *mt_dev = nullptr;
  MatrixPtr mask;
};

class RewrapANNComponent : public ANNComponent {
 public:
  RewrapANNComponent(const std::string &name, const std::vector<int> &size);
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  void build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) override;
  const char *kind() const override { return "rewrap"; }
  std::vector<int> size;
};

class FlattenANNComponent : public ANNComponent {
 public:
  explicit FlattenANNComponent(const std::string &name) : ANNComponent(name, "", 0, 0) {}
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  const char *kind() const override { return "flatten"; }
};

class ConvolutionANNComponent : public ANNComponent {
 public:
  ConvolutionANNComponent(const std::string &name, const std::string &wname, const std::vector<int> &kernel,
                          const std::vector<int> &step, int n);
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  void computeAllGradients(MatrixDict &grads) override;
  void reset(unsigned it = 0) override;
  void build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) override;
  const char *kind() const override { return "convolution"; }
  int sharedCountContribution() const override { return number_input_windows; }
  std::vector<int> kernel, step;
  int n;
  int number_input_windows = 0;
  MatrixPtr weights_matrix;
};

class ConvolutionBiasANNComponent : public ANNComponent {
 public:
  ConvolutionBiasANNComponent(const std::string &name, const std::string &wname, int n);
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  void computeAllGradients(MatrixDict &grads) override;
  void reset(unsigned it = 0) override;
  void build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) override;
  const char *kind() const override { return "convolution_bias"; }
  int sharedCountContribution() const override { return number_input_windows; }
  int n;
  int number_input_windows = 0;
  MatrixPtr bias_vector;
};

class MaxPoolingANNComponent : public ANNComponent {
 public:
  MaxPoolingANNComponent(const std::string &name, const std::vector<int> &kernel, const std::vector<int> &step);
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  void reset(unsigned it = 0) override;
  const char *kind() const override { return "max_pooling"; }
  std::vector<int> kernel, step;
  // int32 positions of the selected inputs (maxpooling_component.cc:166-189), one block per forward pass,
  // allocated like an activation: a captured step graph keeps its own block alive (Graph::keep)
  MatrixPtr argmax;
};

// stack_component.cc:81-104, with hyperplane_component.cc:77-97 flattened into it.
// Recognised runs (dot_product [+ bias] [+ element-wise actf], convolution [+ convolution_bias]
// [+ element-wise actf]) execute as one fused launch; anything else runs component by component.
class StackANNComponent : public ANNComponent {
 public:
  explicit StackANNComponent(const std::string &name) : ANNComponent(name, "", 0, 0) {}
  void pushComponent(const ComponentPtr &c) { components.push_back(c); }
  MatrixPtr doForward(const MatrixPtr &input, bool during_training) override;
  MatrixPtr doBackprop(const MatrixPtr &error_input) override;
  void computeAllGradients(MatrixDict &grads) override;
  void reset(unsigned it = 0) override;
  void build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) override;
  const char *kind() const override { return "stack"; }
  void setContext(b200_ctx *c) override;
  std::vector<ComponentPtr> components;   // as pushed (hyperplanes are nested stacks)
  bool fuse = true;                       // false: run every component separately (all tokens observable)
  bool zero_accumulate = false;           // trainer: data gradients accumulate into buffers zeroed during the forward pass
  bool skip_input_gradient = false;       // trainer: the network-input gradient is never used
  ANNComponent *lastComponent();
  // set by the trainer when the loss kernel already produced d(loss)/d(pre-activation)
  bool last_actf_backprop_is_identity = false;
  // set by the trainer: stop before a trailing row-wise activation so that the loss kernel
  // can fuse log_softmax + loss + gradient
  bool defer_last_actf = false;
  // with defer_last_actf: also stop before a trailing dot_product [+ bias] with at most 16 outputs, so that
  // the trainer can run the whole output layer (forward, log_softmax, loss, gradient, data gradient of the
  // layer below) as one launch; doForward then returns that layer's input and fills deferred_dot / _bias
  bool defer_output_layer = false;
  DotProductANNComponent *deferred_dot = nullptr;
  BiasANNComponent *deferred_bias = nullptr;
  // set by the trainer after that launch: the data gradient of deferred_dot already exists
  DotProductANNComponent *precomputed_for = nullptr;
  MatrixPtr precomputed_dx;
  // trainer: compute every component's weight gradients inside doBackprop, as soon as its error
  // input exists (reverse layer order), so that finished gradients can be all-reduced while the
  // rest of the backward pass runs.  on_gradients_ready(component) fires after each one.
  MatrixDict *interleave_grads = nullptr;
  std::function<void(ANNComponent *)> on_gradients_ready;
  // trainer, single replica: weight gradients are issued on side branches of the step (branch 1: bias
  // gradients and other light work, branch 2: the big contractions) while the data gradients stay on
  // the main stream; on_backprop_issued(component) fires once the component's own data gradient has
  // been issued, i.e. when nothing later in the step reads its weights any more.
  bool use_branches = false;
  // dgrad / wgrad of a layer: 1 = side by side, half the SMs each; 2 = one after the other, full width
  // (the weight gradient still on its branch); 0 = both full width at once
  int contraction_mode = 1;
  std::function<void(ANNComponent *, int branch)> on_backprop_issued;
  void prepareGradScales();   // sets grad_scale of every weight-bearing component from grad_bunch
  const std::vector<ANNComponent *> &flatComponents() const { return flat; }
 private:
  void flatten(std::vector<ANNComponent *> &out);
  std::vector<ANNComponent *> flat;
  friend class SupervisedTrainer;
};

ComponentPtr makeHyperplane(const std::string &name, unsigned in, unsigned out, const std::string &dot_name,
                            const std::string &bias_name, const std::string &dot_weights,
                            const std::string &bias_weights);
// ann.mlp.all_all.generate  (packages/ann/ann/lua_src/annbase.lua:509-660)
std::shared_ptr<StackANNComponent> mlpAllAllGenerate(const std::string &topology);
int actfFromName(const std::string &kind);   // throws on unknown names

// ---------------------------------------------------------------- loss
enum LossKind { LOSS_MSE = 0, LOSS_CROSS_ENTROPY = 1, LOSS_MULTI_CLASS_CROSS_ENTROPY = 2, LOSS_ZERO_ONE = 3 };

class LossFunction {
 public:
  LossFunction(b200_ctx *ctx, int kind, unsigned size);
  ~LossFunction();
  // computeLoss: per-pattern loss vector [bunch] on the device (loss_function.h:109-113)
  MatrixPtr computeLoss(const MatrixPtr &input, const MatrixPtr &target);
  MatrixPtr computeGradient(const MatrixPtr &input, const MatrixPtr &target);
  // fused log_softmax + MCCE: logits -> (logp, loss rows, gradient) in one pass
  void fusedLogSoftmaxMCCE(const MatrixPtr &logits, const MatrixPtr &target, MatrixPtr &logp,
                           MatrixPtr &loss_rows, MatrixPtr &grad);
  void accumLoss(const MatrixPtr &loss_rows);       // device-side running statistics
  void getAccumLoss(float *mean, float *variance);  // one D2H of 3 doubles (loss_function.h:90-95)
  void reset();
  int kind;
  unsigned size;
  float TH = 0.5f;               // zero_one: decision threshold of the two-class case
  b200_ctx *ctx;
  double *stats_dev = nullptr;   // sum, sum of squares, count
};

// ---------------------------------------------------------------- optimizer
// ann.optimizer.{sgd,adagrad,rmsprop,adadelta}: option tables with the reference's defaults
// (optimizer_sgd.lua:39-47, optimizer_adagrad.lua:20-40, optimizer_rmsprop.lua:20-42, optimizer_adadelta.lua:20-44)
class SGDOptimizer {
 public:
  SGDOptimizer();
  int kind = B200_OPT_SGD;
  void setKind(int kind);               // resets the options to that optimizer's defaults
  bool validOption(const std::string &name) const;
  void setOption(const std::string &name, double v);
  double getOption(const std::string &name) const;
  void setLayerwiseOption(const std::string &layer, const std::string &name, double v);
  double getOptionOf(const std::string &layer, const std::string &name) const;
  std::map<std::string, double> global_options;
  std::map<std::string, std::map<std::string, double>> layerwise_options;
  int64_t count = 0;     // host mirror of the device counter
};

// ---------------------------------------------------------------- trainer
class SupervisedTrainer {
 public:
  SupervisedTrainer(b200_ctx *ctx, const std::shared_ptr<StackANNComponent> &net, int loss_kind, int bunch_size);
  ~SupervisedTrainer();
  void build(unsigned input = 0, unsigned output = 0);
  void setOption(const std::string &name, double v);
  double getOption(const std::string &name) const { return optimizer.getOption(name); }
  void setLayerwiseOption(const std::string &pattern, const std::string &name, double v);
  void randomizeWeights(MTRand &rnd, double inf, double sup, bool use_fanin, bool use_fanout,
                        const std::string &name_match);
  // one training step on a device-resident bunch; returns nothing on the host (the per-row
  // losses stay on the device in last_loss_rows).  supervised.lua:725-821
  // smoothing_bunch: the bunch_size of supervised.lua:757,800 (0 = the trainer's bunch_size);
  // max_gradients_norm: the global-norm clip of supervised.lua:805-811 (0 = off)
  void trainStepDevice(const MatrixPtr &x, const MatrixPtr &t, int smoothing_bunch = 0, double max_gradients_norm = 0.0);
  void validateStepDevice(const MatrixPtr &x, const MatrixPtr &t);
  // host-pointer API: H2D of the bunch, step, D2H of the bunch-mean loss
  float trainStep(const float *x, const float *t, int bunch, float *loss_rows_out, int smoothing_bunch = 0,
                  double max_gradients_norm = 0.0);
  // use_dataset (supervised.lua:1291-1430): forward only, bunch by bunch, outputs to the host
  void useDataset(const float *x, int n, float *y);
  // checkpoint / resume of the optimizer state (optimizer_sgd.lua:102-119 exports options + count + update)
  int64_t getCount();
  void setCount(int64_t c);
  void setOptimizer(int kind);
  void invalidateGraphs();
  void sgd_dirty_public() { sgd_dirty = true; }
  void ensureOptimizerState() { if (sgd_dirty || (optimizer.kind != B200_OPT_SGD && !state1_arena)) uploadSgdTable(); }
  MatrixDict state1, state2;            // adagrad Egradients / rmsprop Erms / adadelta Egradients ; adadelta Eupdates
  MatrixPtr state1_arena, state2_arena;
  float validateStep(const float *x, const float *t, int bunch, float *loss_rows_out);
  // dataset loops (supervised.lua:1149-1226, trainable.lua:95-330): the dataset is uploaded
  // once, bunches are gathered on the device, the epoch mean is read back once.
  void trainDataset(const float *x, const float *t, int n, const int *order, float *mean, float *var);
  void validateDataset(const float *x, const float *t, int n, float *mean, float *var);
  MatrixPtr calculate(const MatrixPtr &x);
  void stage(const float *x, const float *t, int bunch);   // async H2D into the staging buffers
  void stepStaged(int bunch);                              // train step on the staged bunch
  std::vector<std::string> weightNames() const { return weights_order; }
  size_t dp_bucket_bytes = 4u << 20;   // gradients are all-reduced in buckets of at least this size
  MatrixPtr weight(const std::string &n) { return weights_table.at(n); }
  MatrixPtr gradient(const std::string &n) { return grads.at(n); }
  double norm2(const std::string &pattern);

  // data-parallel replica group (weights identical on every rank; gradients summed)
  void setDataParallel(int nranks, int rank) { dp_nranks = nranks; dp_rank = rank; }
  void broadcastWeights();
  // replica group over NVLink peer memory (b200_dp_fused_update): dpExport writes the CUDA IPC handles of
  // this rank's weight arena, gradient arena and flag block (3 x 64 bytes); dpConnect maps every peer's
  void dpExport(int nranks, unsigned char *handles256);
  void dpConnect(int nranks, int rank, const unsigned char *all_handles);
  // replica group over SYMMETRIC memory: every rank's [weights | gradients | flags] live at the same offsets of a
  // buffer that is mapped into every process (bases[r]) and bound to a multicast object (mc_base, may be NULL).
  // The trainer moves its weight and gradient arenas into its own buffer (bases[rank]).
  void dpConnectSymmetric(int nranks, int rank, void *const *bases, void *mc_base, size_t bytes);
  static size_t dpSymmetricBytes(size_t arena_floats);
  void rehomeArenas(float *w, float *g);
  unsigned build_in = 0, build_out = 0;
  float dpBench(int reps);
  bool dp_fused = false;
  b200_dp_group dp_group = {};
  long long *dp_flags = nullptr;
  bool dp_flags_owned = true;            // false: the flag block lives in the caller's symmetric buffer
  float *dp_recv = nullptr;
  std::vector<void *> dp_imported;

  b200_ctx *ctx;
  std::shared_ptr<StackANNComponent> net;
  LossFunction loss;
  SGDOptimizer optimizer;
  int bunch_size;
  bool smooth_gradients = true;
  bool keep_gradients = false;    // write the regularised gradient back (observable grads; +4 B/param)
  bool use_cuda_graph = true;
  bool use_branches = true;       // single replica: weight gradients / updates / statistics on side branches
  bool sgd_as_ready = true;       // single replica: update each big tensor as soon as it is ready (false: one launch at the end)
  bool fuse_output_layer = true;  // <=16-class output layer + log_softmax + MCCE + data gradient in one launch
  // tensor-core mode: zero the big gradient tensors under the forward pass and let the contractions ACCUMULATE
  // (TMA reduce-add stores, exchange-free split-K).  Measured on C2: the contraction alone gains 8 % (18.5 vs
  // 20.0 us), but the 30 MB of memsets under the forward pass cost more than that (step 118 -> 125 us): off.
  bool zero_accumulate = false;
  MatrixDict weights_table, grads, updates;
  std::vector<std::string> weights_order;   // sorted names (initialisation / API order)
  std::vector<std::string> arena_order;     // layout of the flat arenas: reverse layer order
  MatrixPtr weights_arena, grads_arena, updates_arena;
  MatrixPtr last_loss_rows, last_output;
  int dp_nranks = 1, dp_rank = 0;
  int64_t *count_dev = nullptr;
  size_t numParameters() const { return total_params; }

 private:
  void runStep(const MatrixPtr &x, const MatrixPtr &t, int global_bunch, double max_gradients_norm);
  void runUpdateSimple(double max_gradients_norm);
  bool dpBucketPlan(std::vector<std::pair<int, int>> *out);   // dp_bucket_plan over this trainer's gradients
  b200_opt_tensor *opt_dev = nullptr;
  std::vector<b200_opt_tensor> opt_host;
  float *norm_dev = nullptr;
  MatrixPtr outputLayerFused(const MatrixPtr &h, const MatrixPtr &t, bool training, MatrixPtr &logp, MatrixPtr &rows,
                             MatrixPtr &grad);
  void uploadSgdTable();
  size_t total_params = 0;
  b200_sgd_tensor *sgd_dev = nullptr;
  std::vector<b200_sgd_tensor> sgd_host;
  // single replica: the tensors of the big contractions are updated one launch each, as soon as they are
  // ready; all the others (biases, the output layer) in one launch over this second table
  b200_sgd_tensor *sgd_light_dev = nullptr;
  std::vector<b200_sgd_tensor> sgd_light_host;
  std::vector<char> tensor_heavy;   // by arena index
  bool sgd_dirty = true;
  std::string sgd_signature;
  // staging + device-resident dataset
  MatrixPtr stage_x, stage_t;
  // pipelined host feeding (stage / stepStaged): two staging slots filled on a copy stream, so that the
  // host-to-device copy of bunch k+1 runs while bunch k is being trained
  MatrixPtr pipe_x[2], pipe_t[2];
  void *copy_stream = nullptr, *ev_copied[2] = {nullptr, nullptr}, *ev_trained[2] = {nullptr, nullptr};
  bool slot_trained[2] = {false, false};
  int next_slot = 0, staged_slot = -1;
  struct Graph;
  // (bunch rows, smoothing bunch, input buffer) -> captured step
  std::map<std::pair<std::pair<int, int>, const float *>, Graph *> graphs;
  double graph_max_norm = 0.0;          // the clip threshold baked into the captured graphs
};

bool luaPatternMatch(const std::string &pattern, const std::string &s);

}  // namespace b200
