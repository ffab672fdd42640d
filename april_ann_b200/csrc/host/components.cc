// ANN components of the hot path over the C ABI.  Each class follows the reference
// component of the same name (packages/ann/ann/c_src/<name>_component.cc); the stack fuses
// recognised runs into single launches.
#include <math.h>
#include <string.h>

#include <exception>
#include <sstream>

#include "ann.h"

namespace b200 {

static std::vector<int> dims2(int a, int b) { return std::vector<int>{a, b}; }

static MatrixPtr gradFor(MatrixDict &grads, const std::string &name, const MatrixPtr &like, float *beta) {
  auto it = grads.find(name);
  MatrixPtr g;
  if (it == grads.end()) {
    g = Matrix::create(like->ctx, like->dims);  // cloneOnlyDims + zeros (dot_product_component.cc:178-183)
    g->fresh = true;
    grads[name] = g;
  } else {
    g = it->second;
    if (g->dims != like->dims) throw Error(B200_ERR_BAD_ARG, "Incorrect weights matrix dimensions");
  }
  *beta = g->fresh ? 0.0f : 1.0f;  // the first writer of a step overwrites (== zeros + accumulate)
  g->fresh = false;
  return g;
}

// ------------------------------------------------------------------ base
void ANNComponent::reset(unsigned) {
  input.reset();
  output.reset();
  error_input.reset();
  error_output.reset();
}
void ANNComponent::build(unsigned in, unsigned out, MatrixDict &, ComponentDict &components) {
  if (input_size == 0) input_size = in;
  if (output_size == 0) output_size = out;
  if (in != 0 && input_size != in)
    throw Error(B200_ERR_BAD_ARG, "Incorrect input size at component " + name);
  if (out != 0 && output_size != out)
    throw Error(B200_ERR_BAD_ARG, "Incorrect output size at component " + name);
  if (!name.empty()) {
    if (components.count(name) && components[name] != this)
      throw Error(B200_ERR_BAD_ARG, "Non unique component name found: " + name);
    components[name] = this;
  }
}

// ------------------------------------------------------------------ dot_product
DotProductANNComponent::DotProductANNComponent(const std::string &name, const std::string &wname, unsigned in,
                                               unsigned out)
    : ANNComponent(name, wname.empty() ? name : wname, in, out) {}

MatrixPtr DotProductANNComponent::doForward(const MatrixPtr &in, bool) {
  if (!weights_matrix) throw Error(B200_ERR_NOT_BUILT, "Not built component " + name);
  if (in->dims.size() < 2) throw Error(B200_ERR_BAD_ARG, "At 2-dimensional matrix is expected [" + name + "]");
  if ((unsigned)in->cols() != input_size) throw Error(B200_ERR_BAD_ARG, "Incorrect input size [" + name + "]");
  input = in;
  const int bunch = in->rows();
  output = Matrix::create(ctx, dims2(bunch, (int)output_size));
  check(b200_linear_fwd(ctx, bunch, (int)output_size, (int)input_size, in->data, (int)input_size,
                        weights_matrix->data, (int)input_size, nullptr, B200_ACT_NONE, output->data,
                        (int)output_size));
  return output;
}
MatrixPtr DotProductANNComponent::doBackprop(const MatrixPtr &err) {
  error_input = err;
  const int bunch = err->rows();
  error_output = Matrix::create(ctx, dims2(bunch, (int)input_size));
  check(b200_linear_bwd_data(ctx, bunch, (int)output_size, (int)input_size, err->data, (int)output_size,
                             weights_matrix->data, (int)input_size, B200_ACT_NONE, nullptr, 0,
                             error_output->data, (int)input_size));
  return error_output;
}
void DotProductANNComponent::computeAllGradients(MatrixDict &grads) {
  if (!input || !error_input) throw Error(B200_ERR_BAD_ARG, "computeGradients before forward/backprop [" + name + "]");
  weights_matrix->shared_count += 1;
  float beta;
  MatrixPtr g = gradFor(grads, weights_name, weights_matrix, &beta);
  const int bunch = error_input->rows();
  check(b200_linear_bwd_weight(ctx, bunch, (int)output_size, (int)input_size, error_input->data, (int)output_size,
                               input->data, (int)input_size, grad_scale, beta, g->data, (int)input_size, nullptr));
}
void DotProductANNComponent::reset(unsigned it) {
  ANNComponent::reset(it);
  if (weights_matrix) weights_matrix->shared_count = 0;
}
void DotProductANNComponent::build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) {
  ANNComponent::build(in, out, weights, components);
  if (input_size == 0 || output_size == 0)
    throw Error(B200_ERR_BAD_ARG, "Impossible to compute input/output sizes for this component [" + name + "]");
  auto it = weights.find(weights_name);
  if (it != weights.end()) {
    weights_matrix = it->second;
    if (weights_matrix->dims != dims2((int)output_size, (int)input_size))
      throw Error(B200_ERR_BAD_ARG, "The weights matrix input/output sizes are not correct [" + name + "]");
  } else {
    if (!weights_matrix) weights_matrix = Matrix::create(ctx, dims2((int)output_size, (int)input_size));
    weights[weights_name] = weights_matrix;
  }
}

// ------------------------------------------------------------------ bias
BiasANNComponent::BiasANNComponent(const std::string &name, const std::string &wname, unsigned size)
    : ANNComponent(name, wname.empty() ? name : wname, size, size) {}

MatrixPtr BiasANNComponent::doForward(const MatrixPtr &in, bool) {
  if (!bias_vector) throw Error(B200_ERR_NOT_BUILT, "Not built component " + name);
  input = in;
  output = Matrix::create(ctx, in->dims);
  check(b200_bias_fwd(ctx, in->rows(), in->cols(), in->data, bias_vector->data, output->data));
  return output;
}
MatrixPtr BiasANNComponent::doBackprop(const MatrixPtr &err) {
  error_input = err;
  error_output = err;  // bias_component.cc:76-79: by-pass
  return err;
}
void BiasANNComponent::computeAllGradients(MatrixDict &grads) {
  if (!error_input) throw Error(B200_ERR_BAD_ARG, "computeGradients before backprop [" + name + "]");
  bias_vector->shared_count += 1;
  float beta;
  MatrixPtr g = gradFor(grads, weights_name, bias_vector, &beta);
  check(b200_bias_grad(ctx, error_input->rows(), error_input->cols(), error_input->data, error_input->cols(),
                       grad_scale, beta, g->data));
}
void BiasANNComponent::reset(unsigned it) {
  ANNComponent::reset(it);
  if (bias_vector) bias_vector->shared_count = 0;
}
void BiasANNComponent::build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) {
  ANNComponent::build(in, out, weights, components);
  if (input_size == 0 && output_size == 0)
    throw Error(B200_ERR_BAD_ARG, "Impossible to compute input/output sizes for this component [" + name + "]");
  if (input_size == 0) input_size = output_size;
  if (output_size == 0) output_size = input_size;
  if (input_size != output_size)
    throw Error(B200_ERR_BAD_ARG, "BiasANNComponent input/output sizes must be equal [" + name + "]");
  auto it = weights.find(weights_name);
  if (it != weights.end()) {
    bias_vector = it->second;
    if (bias_vector->dims != dims2((int)output_size, 1))
      throw Error(B200_ERR_BAD_ARG, "The weights matrix input/output sizes are not correct [" + name + "]");
  } else {
    if (!bias_vector) bias_vector = Matrix::create(ctx, dims2((int)output_size, 1));
    weights[weights_name] = bias_vector;
  }
}

// ------------------------------------------------------------------ activation functions
ActivationFunctionANNComponent::ActivationFunctionANNComponent(const std::string &name, int act, float p0, float p1)
    : ANNComponent(name, "", 0, 0), act(act), p0(p0), p1(p1) {
  if (act == B200_ACT_HARDTANH && !(p0 < p1)) throw Error(B200_ERR_BAD_ARG, "hardtanh: inf must be < sup [" + name + "]");
}

void ActivationFunctionANNComponent::build(unsigned in, unsigned out, MatrixDict &w, ComponentDict &c) {
  // activation_function_component.cc:150-166: input and output sizes are the same
  if (input_size == 0) input_size = in ? in : out;
  if (output_size == 0) output_size = input_size;
  ANNComponent::build(in, out, w, c);
}

MatrixPtr ActivationFunctionANNComponent::doForward(const MatrixPtr &in, bool) {
  input = in;
  output = Matrix::create(ctx, in->dims);
  if (act == B200_ACT_SOFTMAX)
    check(b200_softmax_fwd(ctx, in->rows(), in->cols(), in->data, output->data));
  else if (act == B200_ACT_LOG_SOFTMAX)
    check(b200_log_softmax_fwd(ctx, in->rows(), in->cols(), in->data, output->data));
  else if (act >= B200_ACT_LOG_LOGISTIC)
    check(b200_actf_fwd_ex(ctx, act, p0, p1, in->size(), in->data, output->data));
  else
    check(b200_actf_fwd(ctx, act, in->size(), in->data, output->data));
  return output;
}
MatrixPtr ActivationFunctionANNComponent::doBackprop(const MatrixPtr &err) {
  if (!output) throw Error(B200_ERR_BAD_ARG, "backprop before forward [" + name + "]");
  if (err->size() != output->size())
    throw Error(129, "Different bunches found at doForward and doBackprop [" + name + "]");
  error_input = err;
  if (act == B200_ACT_LOG_SOFTMAX || act == B200_ACT_LOG_LOGISTIC) {
    // log_softmax_actf_component.cc:44-52, log_logistic_actf_component.cc:44-53: the derivative is
    // cancelled by the cross-entropy derivative; the reference copies, the copy is not needed on an
    // immutable token.
    error_output = err;
    return err;
  }
  error_output = Matrix::create(ctx, err->dims);
  if (act == B200_ACT_SOFTMAX)
    check(b200_softmax_bwd(ctx, err->rows(), err->cols(), output->data, err->data, error_output->data));
  else if (act >= B200_ACT_LOG_LOGISTIC)
    check(b200_actf_bwd_ex(ctx, act, p0, p1, err->size(), input ? input->data : nullptr, output->data, err->data,
                           error_output->data));
  else
    check(b200_actf_bwd(ctx, act, err->size(), output->data, err->data, error_output->data));
  return error_output;
}

int actfFromName(const std::string &k) {
  if (k == "logistic") return B200_ACT_LOGISTIC;
  if (k == "tanh") return B200_ACT_TANH;
  if (k == "relu") return B200_ACT_RELU;
  if (k == "softmax") return B200_ACT_SOFTMAX;
  if (k == "log_softmax") return B200_ACT_LOG_SOFTMAX;
  if (k == "linear") return B200_ACT_LINEAR;
  if (k == "log_logistic") return B200_ACT_LOG_LOGISTIC;
  if (k == "softplus") return B200_ACT_SOFTPLUS;
  if (k == "softsign") return B200_ACT_SOFTSIGN;
  if (k == "leaky_relu") return B200_ACT_LEAKY_RELU;
  if (k == "hardtanh") return B200_ACT_HARDTANH;
  throw Error(B200_ERR_BAD_ARG, "Incorrect component class: " + k);
}

// ------------------------------------------------------------------ prelu
PReLUActfANNComponent::PReLUActfANNComponent(const std::string &name, const std::string &wname, unsigned size, bool scalar)
    : ANNComponent(name, wname.empty() ? name : wname, size, size), scalar(scalar) {}
void PReLUActfANNComponent::build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) {
  if (input_size == 0) input_size = in ? in : out;
  if (output_size == 0) output_size = input_size;
  ANNComponent::build(in, out, weights, components);
  if (input_size == 0) throw Error(256, "Unable to allocate prelu weights [" + name + "]");
  const int wsize = scalar ? 1 : (int)input_size;
  auto it = weights.find(weights_name);
  if (it != weights.end()) {
    weights_matrix = it->second;
  } else {
    if (!weights_matrix) weights_matrix = Matrix::create(ctx, std::vector<int>{wsize, 1});
    weights[weights_name] = weights_matrix;
  }
  if ((int)weights_matrix->size() != wsize) throw Error(257, "Unexpected matrix size [" + name + "]");
}
MatrixPtr PReLUActfANNComponent::doForward(const MatrixPtr &in, bool) {
  if (!weights_matrix) throw Error(B200_ERR_NOT_BUILT, "Not built component " + name);
  if ((unsigned)in->cols() != input_size) throw Error(B200_ERR_BAD_ARG, "Incorrect input size [" + name + "]");
  input = in;
  output = Matrix::create(ctx, in->dims);
  check(b200_prelu_fwd(ctx, in->rows(), in->cols(), in->data, weights_matrix->data, scalar ? 1 : 0, output->data));
  return output;
}
MatrixPtr PReLUActfANNComponent::doBackprop(const MatrixPtr &err) {
  if (!input) throw Error(B200_ERR_BAD_ARG, "backprop before forward [" + name + "]");
  error_input = err;
  error_output = Matrix::create(ctx, err->dims);
  check(b200_prelu_bwd(ctx, input->rows(), input->cols(), input->data, weights_matrix->data, scalar ? 1 : 0, err->data,
                       error_output->data));
  return error_output;
}
void PReLUActfANNComponent::computeAllGradients(MatrixDict &grads) {
  if (!input || !error_input) throw Error(B200_ERR_BAD_ARG, "computeGradients before forward/backprop [" + name + "]");
  weights_matrix->shared_count += 1;
  float beta;
  MatrixPtr g = gradFor(grads, weights_name, weights_matrix, &beta);
  MatrixPtr tmp = Matrix::create(ctx, input->dims);
  check(b200_prelu_grad(ctx, input->rows(), input->cols(), input->data, error_input->data, scalar ? 1 : 0, grad_scale, beta,
                        g->data, tmp->data));
}
void PReLUActfANNComponent::reset(unsigned it) {
  ANNComponent::reset(it);
  if (weights_matrix) weights_matrix->shared_count = 0;
}

// ------------------------------------------------------------------ dropout
DropoutANNComponent::DropoutANNComponent(const std::string &name, const MTRand &random, float prob, float value, bool norm,
                                         unsigned size)
    : ANNComponent(name, "", size, size), random(random), prob(prob), value(value), normalize_after_training(norm) {}
DropoutANNComponent::~DropoutANNComponent() {
  if (mt_dev) b200_free(ctx, mt_dev);
}
void DropoutANNComponent::build(unsigned in, unsigned out, MatrixDict &w, ComponentDict &c) {
  if (input_size == 0) input_size = in ? in : out;
  if (output_size == 0) output_size = input_size;
  ANNComponent::build(in, out, w, c);
  if (input_size != output_size) throw Error(128, "Incorrect input/output sizes [" + name + "]");
}
MatrixPtr DropoutANNComponent::doForward(const MatrixPtr &in, bool during_training) {
  input = in;
  if (!(prob > 0.0f && (during_training || normalize_after_training))) {
    output = in;   // dropout_component.cc:107-109
    return output;
  }
  output = Matrix::create(ctx, in->dims);
  if (during_training) {
    if (!mt_dev) {
      struct { uint32_t words[624]; int32_t next; int32_t pad[3]; } host;
      memset(&host, 0, sizeof(host));
      random.exportState(host.words, &host.next);
      if (sizeof(host) != b200_mt_state_bytes()) throw Error(B200_ERR_BAD_ARG, "MT19937 state layout mismatch");
      check(b200_malloc(ctx, &mt_dev, sizeof(host)));
      check(b200_memcpy_h2d(ctx, mt_dev, &host, sizeof(host)));
      check(b200_sync(ctx));
    }
    mask = Matrix::create(ctx, in->dims);
    check(b200_dropout_mask(ctx, mt_dev, in->size(), prob, mask->data));
    check(b200_mask_apply(ctx, in->size(), in->data, mask->data, value, output->data));
  } else {
    // matScal(output, 1 - prob): y = (1-prob) * x
    check(b200_memcpy_d2d(ctx, output->data, in->data, in->size() * sizeof(float)));
    check(b200_sscal(ctx, in->size(), 1.0f - prob, output->data));
  }
  return output;
}
MatrixPtr DropoutANNComponent::doBackprop(const MatrixPtr &err) {
  error_input = err;
  if (mask && prob > 0.0f) {
    if (err->size() != mask->size()) throw Error(129, "Different bunches found at doForward and doBackprop [" + name + "]");
    error_output = Matrix::create(ctx, err->dims);
    check(b200_mask_apply(ctx, err->size(), err->data, mask->data, 0.0f, error_output->data));
  } else {
    error_output = err;
  }
  return error_output;
}

// ------------------------------------------------------------------ rewrap / flatten
RewrapANNComponent::RewrapANNComponent(const std::string &name, const std::vector<int> &size)
    : ANNComponent(name, "", 0, 0), size(size) {}
void RewrapANNComponent::build(unsigned in, unsigned out, MatrixDict &w, ComponentDict &c) {
  unsigned total = 1;
  for (int d : size) total *= (unsigned)d;
  input_size = output_size = total;
  ANNComponent::build(in, out, w, c);
}
MatrixPtr RewrapANNComponent::doForward(const MatrixPtr &in, bool) {
  input = in;
  std::vector<int> d{in->rows()};
  d.insert(d.end(), size.begin(), size.end());
  output = in->rewrap(d);
  return output;
}
MatrixPtr RewrapANNComponent::doBackprop(const MatrixPtr &err) {
  error_input = err;
  error_output = err->rewrap(input->dims);
  return error_output;
}
MatrixPtr FlattenANNComponent::doForward(const MatrixPtr &in, bool) {
  input = in;
  output = in->rewrap(dims2(in->rows(), in->cols()));
  return output;
}
MatrixPtr FlattenANNComponent::doBackprop(const MatrixPtr &err) {
  error_input = err;
  error_output = err->rewrap(input->dims);
  return error_output;
}

// ------------------------------------------------------------------ convolution
ConvolutionANNComponent::ConvolutionANNComponent(const std::string &name, const std::string &wname,
                                                 const std::vector<int> &kernel, const std::vector<int> &step, int n)
    : ANNComponent(name, wname.empty() ? name : wname, 0, 0), kernel(kernel), step(step), n(n) {
  if (kernel.size() != 3) throw Error(B200_ERR_UNSUPPORTED, "convolution: only {planes,kh,kw} kernels are supported");
  if (this->step.empty()) this->step = std::vector<int>{1, 1, 1};  // bind_ann_base.lua.cc:1371-1373
  if (this->step[0] != 1) throw Error(B200_ERR_BAD_ARG, "convolution: step over planes must be 1");
}
void ConvolutionANNComponent::build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) {
  ANNComponent::build(0, 0, weights, components);
  const int ks = kernel[0] * kernel[1] * kernel[2];
  auto it = weights.find(weights_name);
  if (it != weights.end()) {
    weights_matrix = it->second;
    if (weights_matrix->dims != dims2(n, ks))
      throw Error(B200_ERR_BAD_ARG, "The weights matrix input/output sizes are not correct [" + name + "]");
  } else {
    if (!weights_matrix) weights_matrix = Matrix::create(ctx, dims2(n, ks));
    weights[weights_name] = weights_matrix;
  }
}
MatrixPtr ConvolutionANNComponent::doForward(const MatrixPtr &in, bool) {
  if (!weights_matrix) throw Error(B200_ERR_NOT_BUILT, "Not built component " + name);
  if (in->dims.size() != 4) throw Error(129, "Incorrect input matrix numDims [" + name + "]");
  if (in->dim(1) != kernel[0])
    throw Error(128, "Input matrix dim 1 must be equals to kernel dim 1 [" + name + "]");
  input = in;
  const int B = in->dim(0), C = in->dim(1), H = in->dim(2), W = in->dim(3);
  const int oH = (H - kernel[1]) / step[1] + 1, oW = (W - kernel[2]) / step[2] + 1;
  number_input_windows = oH * oW;
  output = Matrix::create(ctx, std::vector<int>{B, n, oH, oW});
  check(b200_conv2d_fwd(ctx, B, C, H, W, n, kernel[1], kernel[2], step[1], step[2], in->data,
                        weights_matrix->data, nullptr, B200_ACT_NONE, output->data));
  return output;
}
MatrixPtr ConvolutionANNComponent::doBackprop(const MatrixPtr &err) {
  error_input = err;
  const int B = input->dim(0), C = input->dim(1), H = input->dim(2), W = input->dim(3);
  error_output = Matrix::create(ctx, input->dims);
  check(b200_conv2d_bwd_data(ctx, B, C, H, W, n, kernel[1], kernel[2], step[1], step[2], err->data,
                             weights_matrix->data, error_output->data));
  return error_output;
}
void ConvolutionANNComponent::computeAllGradients(MatrixDict &grads) {
  if (!input || !error_input) throw Error(B200_ERR_BAD_ARG, "computeGradients before forward/backprop [" + name + "]");
  weights_matrix->shared_count += number_input_windows;
  float beta;
  MatrixPtr g = gradFor(grads, weights_name, weights_matrix, &beta);
  const int B = input->dim(0), C = input->dim(1), H = input->dim(2), W = input->dim(3);
  check(b200_conv2d_bwd_weight(ctx, B, C, H, W, n, kernel[1], kernel[2], step[1], step[2], error_input->data,
                               input->data, grad_scale, beta, g->data, nullptr));
}
void ConvolutionANNComponent::reset(unsigned it) {
  ANNComponent::reset(it);
  if (weights_matrix) weights_matrix->shared_count = 0;
}

ConvolutionBiasANNComponent::ConvolutionBiasANNComponent(const std::string &name, const std::string &wname, int n)
    : ANNComponent(name, wname.empty() ? name : wname, 0, 0), n(n) {}
void ConvolutionBiasANNComponent::build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &components) {
  ANNComponent::build(0, 0, weights, components);
  auto it = weights.find(weights_name);
  if (it != weights.end()) {
    bias_vector = it->second;
    if (bias_vector->dims != dims2(n, 1))
      throw Error(B200_ERR_BAD_ARG, "The weights matrix input/output sizes are not correct [" + name + "]");
  } else {
    if (!bias_vector) bias_vector = Matrix::create(ctx, dims2(n, 1));
    weights[weights_name] = bias_vector;
  }
}
MatrixPtr ConvolutionBiasANNComponent::doForward(const MatrixPtr &in, bool) {
  if (!bias_vector) throw Error(B200_ERR_NOT_BUILT, "Not built component " + name);
  if (in->dims.size() != 4 || in->dim(1) != n) throw Error(129, "Incorrect input dim[1] size [" + name + "]");
  input = in;
  number_input_windows = in->dim(2) * in->dim(3);
  output = Matrix::create(ctx, in->dims);
  check(b200_conv_bias_fwd(ctx, in->dim(0), n, in->dim(2) * in->dim(3), in->data, bias_vector->data, output->data));
  return output;
}
MatrixPtr ConvolutionBiasANNComponent::doBackprop(const MatrixPtr &err) {
  error_input = err;
  error_output = err;
  return err;
}
void ConvolutionBiasANNComponent::computeAllGradients(MatrixDict &grads) {
  if (!error_input) throw Error(B200_ERR_BAD_ARG, "computeGradients before backprop [" + name + "]");
  bias_vector->shared_count += number_input_windows;
  float beta;
  MatrixPtr g = gradFor(grads, weights_name, bias_vector, &beta);
  const MatrixPtr &e = error_input;
  check(b200_conv_bias_grad(ctx, e->dim(0), n, e->dim(2) * e->dim(3), e->data, grad_scale, beta, g->data));
}
void ConvolutionBiasANNComponent::reset(unsigned it) {
  ANNComponent::reset(it);
  if (bias_vector) bias_vector->shared_count = 0;
}

// ------------------------------------------------------------------ max pooling
MaxPoolingANNComponent::MaxPoolingANNComponent(const std::string &name, const std::vector<int> &kernel,
                                               const std::vector<int> &step)
    : ANNComponent(name, "", 0, 0), kernel(kernel), step(step) {
  if (kernel.size() != 3 || kernel[0] != 1)
    throw Error(B200_ERR_UNSUPPORTED, "max_pooling: only {1,kh,kw} kernels are supported");
  if (this->step.empty()) this->step = kernel;  // bind_ann_base.lua.cc:1485-1487
}
MatrixPtr MaxPoolingANNComponent::doForward(const MatrixPtr &in, bool) {
  if (in->dims.size() != 4) throw Error(129, "Incorrect input matrix numDims [" + name + "]");
  input = in;
  const int B = in->dim(0), C = in->dim(1), H = in->dim(2), W = in->dim(3);
  const int oH = (H - kernel[1]) / step[1] + 1, oW = (W - kernel[2]) / step[2] + 1;
  output = Matrix::create(ctx, std::vector<int>{B, C, oH, oW});
  argmax = Matrix::create(ctx, output->dims);   // 4-byte elements: holds the int32 positions
  check(b200_maxpool_fwd(ctx, B, C, H, W, kernel[1], kernel[2], step[1], step[2], in->data, output->data,
                         reinterpret_cast<int32_t *>(argmax->data)));
  return output;
}
MatrixPtr MaxPoolingANNComponent::doBackprop(const MatrixPtr &err) {
  error_input = err;
  const int B = input->dim(0), C = input->dim(1), H = input->dim(2), W = input->dim(3);
  error_output = Matrix::create(ctx, input->dims);
  if (!argmax) throw Error(B200_ERR_BAD_ARG, "backprop before forward [" + name + "]");
  check(b200_maxpool_bwd(ctx, B, C, H, W, kernel[1], kernel[2], step[1], step[2], err->data,
                         reinterpret_cast<const int32_t *>(argmax->data), error_output->data));
  return error_output;
}
void MaxPoolingANNComponent::reset(unsigned it) { ANNComponent::reset(it); }

// ------------------------------------------------------------------ stack
void StackANNComponent::setContext(b200_ctx *c) {
  ctx = c;
  for (auto &k : components) k->setContext(c);
}
void StackANNComponent::flatten(std::vector<ANNComponent *> &out) {
  for (auto &c : components) {
    if (auto *s = dynamic_cast<StackANNComponent *>(c.get())) s->flatten(out);
    else out.push_back(c.get());
  }
}
ANNComponent *StackANNComponent::lastComponent() { return flat.empty() ? nullptr : flat.back(); }

void StackANNComponent::build(unsigned in, unsigned out, MatrixDict &weights, ComponentDict &comps) {
  if (components.empty()) throw Error(B200_ERR_BAD_ARG, "stack without components [" + name + "]");
  if (!name.empty()) comps[name] = this;
  unsigned cur = in ? in : input_size;
  for (size_t i = 0; i < components.size(); ++i) {
    ANNComponent *c = components[i].get();
    c->setContext(ctx);
    const bool last = (i + 1 == components.size());
    c->build(cur, last ? out : 0, weights, comps);
    if (i == 0 && c->getInputSize()) input_size = c->getInputSize();
    cur = c->getOutputSize();  // 0 when it depends on the input shape (convolutions)
  }
  output_size = cur ? cur : out;
  flat.clear();
  flatten(flat);
}

void StackANNComponent::reset(unsigned it) {
  ANNComponent::reset(it);
  for (auto *c : flat) c->reset(it);
}

MatrixPtr StackANNComponent::doForward(const MatrixPtr &in, bool during_training) {
  if (flat.empty()) throw Error(B200_ERR_NOT_BUILT, "Not built component " + name);
  input = in;
  MatrixPtr cur = in;
  size_t n = flat.size();
  deferred_dot = nullptr;
  deferred_bias = nullptr;
  if (defer_last_actf) {
    auto *la = dynamic_cast<ActivationFunctionANNComponent *>(flat.back());
    if (la && la->rowwise()) {
      --n;  // the trainer runs it fused with the loss
      if (defer_output_layer && fuse && n >= 1) {
        size_t j = n;
        auto *b = dynamic_cast<BiasANNComponent *>(flat[j - 1]);
        if (b && j >= 2) --j;
        auto *d = dynamic_cast<DotProductANNComponent *>(flat[j - 1]);
        if (d && d->weights_matrix && d->getOutputSize() >= 3 && d->getOutputSize() <= 16 &&
            d->getInputSize() % 4 == 0 && (!b || b->bias_vector)) {
          deferred_dot = d;
          deferred_bias = (j < n) ? b : nullptr;
          n = j - 1;
        }
      }
    }
  }
  size_t i = 0;
  while (i < n) {
    ANNComponent *c = flat[i];
    if (fuse) {
      if (auto *dot = dynamic_cast<DotProductANNComponent *>(c)) {
        BiasANNComponent *b = (i + 1 < n) ? dynamic_cast<BiasANNComponent *>(flat[i + 1]) : nullptr;
        size_t ai = i + (b ? 2 : 1);
        ActivationFunctionANNComponent *a =
            (ai < n) ? dynamic_cast<ActivationFunctionANNComponent *>(flat[ai]) : nullptr;
        if (a && !a->elementwise()) a = nullptr;
        if (cur->dims.size() < 2 || (unsigned)cur->cols() != dot->getInputSize())
          throw Error(B200_ERR_BAD_ARG, "Incorrect input size [" + dot->getName() + "]");
        const int bunch = cur->rows(), K = (int)dot->getInputSize(), N = (int)dot->getOutputSize();
        MatrixPtr y = Matrix::create(ctx, dims2(bunch, N));
        check(b200_linear_fwd(ctx, bunch, N, K, cur->data, K, dot->weights_matrix->data, K,
                              b ? b->bias_vector->data : nullptr, a ? a->act : B200_ACT_NONE, y->data, N));
        dot->input = cur;
        if (b) { b->input.reset(); b->output.reset(); }
        if (a) { a->input.reset(); a->output = y; }
        (a ? (ANNComponent *)a : (b ? (ANNComponent *)b : (ANNComponent *)dot))->output = y;
        // the buffer this layer's data gradient will be accumulated into: allocated and zeroed now, on the
        // weight-gradient branch, far ahead of the backward pass (tensor-core mode, big layers, and only when the
        // derivative of the layer below is a ReLU gate or absent: those epilogues distribute over partial sums)
        dot->dx_zeroed.reset();
        if (zero_accumulate && during_training && use_branches && N > 16 && !(i == 0 && skip_input_gradient)) {
          int mode = B200_MATH_FP32;
          check(b200_get_math_mode(ctx, &mode));
          ActivationFunctionANNComponent *pa = (i >= 1) ? dynamic_cast<ActivationFunctionANNComponent *>(flat[i - 1]) : nullptr;
          const bool gate_ok = !pa || !pa->elementwise() || pa->act == B200_ACT_RELU || pa->act == B200_ACT_LINEAR;
          if (mode == B200_MATH_TF32 && gate_ok && (long)bunch * K >= 128 * 256) {
            dot->dx_zeroed = Matrix::create(ctx, dims2(bunch, K));
            check(b200_branch_begin(ctx, 2));
            dot->dx_zeroed->zeros();
            check(b200_fence_record(ctx, (int)(i & 7)));   // (ids may alias in deep nets: a later mark on the same branch implies the earlier one)
            check(b200_branch_end(ctx));
          }
        }
        cur = y;
        i = ai + (a ? 1 : 0);
        continue;
      }
      if (auto *conv = dynamic_cast<ConvolutionANNComponent *>(c)) {
        ConvolutionBiasANNComponent *b =
            (i + 1 < n) ? dynamic_cast<ConvolutionBiasANNComponent *>(flat[i + 1]) : nullptr;
        size_t ai = i + (b ? 2 : 1);
        ActivationFunctionANNComponent *a =
            (ai < n) ? dynamic_cast<ActivationFunctionANNComponent *>(flat[ai]) : nullptr;
        if (a && !a->elementwise()) a = nullptr;
        if (cur->dims.size() != 4 || cur->dim(1) != conv->kernel[0])
          throw Error(128, "Input matrix dim 1 must be equals to kernel dim 1 [" + conv->getName() + "]");
        const int B = cur->dim(0), C = cur->dim(1), H = cur->dim(2), W = cur->dim(3);
        const int kh = conv->kernel[1], kw = conv->kernel[2], sh = conv->step[1], sw = conv->step[2];
        const int oH = (H - kh) / sh + 1, oW = (W - kw) / sw + 1;
        MatrixPtr y = Matrix::create(ctx, std::vector<int>{B, conv->n, oH, oW});
        check(b200_conv2d_fwd(ctx, B, C, H, W, conv->n, kh, kw, sh, sw, cur->data, conv->weights_matrix->data,
                              b ? b->bias_vector->data : nullptr, a ? a->act : B200_ACT_NONE, y->data));
        conv->input = cur;
        conv->number_input_windows = oH * oW;
        if (b) { b->input.reset(); b->output.reset(); b->number_input_windows = oH * oW; }
        if (a) { a->input.reset(); a->output = y; }
        (a ? (ANNComponent *)a : (b ? (ANNComponent *)b : (ANNComponent *)conv))->output = y;
        cur = y;
        i = ai + (a ? 1 : 0);
        continue;
      }
    }
    cur = c->doForward(cur, during_training);
    ++i;
  }
  output = cur;
  return cur;
}

MatrixPtr StackANNComponent::doBackprop(const MatrixPtr &err) {
  error_input = err;
  MatrixPtr cur = err;
  int i = (int)flat.size() - 1;
  const int n = (int)flat.size();
  while (i >= 0) {
    ANNComponent *c = flat[i];
    int branch = -1;
    bool had_grads = false, deferred = false;
    if (interleave_grads && c->hasWeightsName() && cur) {
      had_grads = true;
      // the error input of this component is final: its weight gradients can be computed now
      c->error_input = cur;
      auto *d = dynamic_cast<DotProductANNComponent *>(c);
      const bool heavy = (d && d->getOutputSize() > 16) || dynamic_cast<ConvolutionANNComponent *>(c);
      const bool has_dgrad = d && heavy && !(i == 0 && skip_input_gradient);
      if (use_branches) branch = heavy ? 2 : 1;
      if (use_branches && has_dgrad && contraction_mode == 2) {
        // full-width contractions one after the other: the weight gradient is issued (on its branch) right
        // after the data gradient of this iteration, see Issued below
        deferred = true;
      } else {
        if (use_branches) {
          check(b200_branch_begin(ctx, branch));
          // a layer's weight gradient (branch) and data gradient (main stream) are independent: plan each
          // persistent contraction for half of the SMs so that they really run side by side
          // (also when no data gradient follows: the update of the previous layer's tensor runs beside
          // this contraction, on the other half of the SMs)
          // -- only while a contraction cannot fill the device by itself (C2: 64-128 tiles).  Contractions
          // with more 128x256 tiles than SMs (4096-wide layers, the 10 000-class layer) run at full width one
          // after the other: side by side at half width they measured 930 us against 2 x 435 us.
          if (d && heavy && contraction_mode == 1) {
            int sms = 0;
            check(b200_sm_count(ctx, &sms));
            const long tiles = (long)((d->getOutputSize() + 127) / 128) * (long)((d->getInputSize() + 255) / 256);
            if (tiles <= sms) check(b200_set_sm_budget(ctx, sms / 2));
          }
        }
        c->computeAllGradients(*interleave_grads);
        if (on_gradients_ready) on_gradients_ready(c);
        if (use_branches) check(b200_branch_end(ctx));
      }
    }
    struct Issued {   // fires on every way out of this iteration
      StackANNComponent *s; ANNComponent *c; int branch; bool had_grads, deferred;
      ~Issued() noexcept(false) {
        if (s->use_branches) b200_set_sm_budget(s->ctx, 0);
        if (std::uncaught_exceptions()) return;
        if (deferred) {
          check(b200_branch_begin(s->ctx, branch));
          c->computeAllGradients(*s->interleave_grads);
          if (s->on_gradients_ready) s->on_gradients_ready(c);
          check(b200_branch_end(s->ctx));
        }
        if (had_grads && s->on_backprop_issued) s->on_backprop_issued(c, branch);
      }
    } issued{this, c, branch, had_grads, deferred};
    if (fuse) {
      auto *la = dynamic_cast<ActivationFunctionANNComponent *>(c);
      if (la && i == n - 1 && last_actf_backprop_is_identity) {
        la->error_input = cur;
        la->error_output = cur;
        --i;
        continue;
      }
      if (auto *dot = dynamic_cast<DotProductANNComponent *>(c)) {
        dot->error_input = cur;
        ActivationFunctionANNComponent *pa =
            (i - 1 >= 0) ? dynamic_cast<ActivationFunctionANNComponent *>(flat[i - 1]) : nullptr;
        if (pa && (!pa->elementwise() || !pa->output)) pa = nullptr;
        if (i == 0 && skip_input_gradient) {
          dot->error_output.reset();
          cur.reset();
          --i;
          continue;
        }
        const int bunch = cur->rows(), K = (int)dot->getInputSize(), N = (int)dot->getOutputSize();
        MatrixPtr dx;
        if (dot == precomputed_for && precomputed_dx) {
          dx = precomputed_dx;   // produced by the fused output-layer launch (same pa decision)
        } else if (dot->dx_zeroed && dot->dx_zeroed->rows() == bunch) {
          dx = dot->dx_zeroed;   // zeroed during the forward pass on branch 2
          dot->dx_zeroed.reset();
          check(b200_fence_wait(ctx, i & 7));
          check(b200_linear_bwd_data_acc(ctx, bunch, N, K, cur->data, N, dot->weights_matrix->data, K,
                                         pa ? pa->act : B200_ACT_NONE, pa ? pa->output->data : nullptr, K, dx->data, K));
        } else {
          dx = Matrix::create(ctx, dims2(bunch, K));
          check(b200_linear_bwd_data(ctx, bunch, N, K, cur->data, N, dot->weights_matrix->data, K,
                                     pa ? pa->act : B200_ACT_NONE, pa ? pa->output->data : nullptr, K, dx->data, K));
        }
        dot->error_output = pa ? MatrixPtr() : dx;
        if (pa) {
          // dx already holds d(loss)/d(pre-activation) of the previous layer
          pa->error_input.reset();
          pa->error_output = dx;
          i -= 2;
        } else {
          i -= 1;
        }
        cur = dx;
        continue;
      }
    }
    if (!cur) throw Error(B200_ERR_BAD_ARG, "backprop reached a component without an error token [" + c->getName() + "]");
    bool only_reshapes_before = skip_input_gradient;
    if (only_reshapes_before)
      for (int j = 0; j < i; ++j)
        if (!dynamic_cast<RewrapANNComponent *>(flat[j])) { only_reshapes_before = false; break; }
    if (only_reshapes_before && dynamic_cast<ConvolutionANNComponent *>(c)) {
      c->error_input = cur;  // weight gradients still need it; the input gradient is unused
      cur.reset();
      break;
    }
    cur = c->doBackprop(cur);
    --i;
  }
  error_output = cur;
  return cur;
}

void StackANNComponent::prepareGradScales() {
  // total uses of every weights matrix in this step -> 1/sqrt(shared_count * bunch)
  std::map<std::string, int> uses;
  for (auto *c : flat)
    if (c->hasWeightsName()) uses[c->getWeightsName()] += c->sharedCountContribution();
  for (auto *c : flat) {
    if (!c->hasWeightsName()) continue;
    int nuse = uses[c->getWeightsName()];
    if (nuse <= 0) nuse = 1;
    c->grad_scale = (grad_bunch > 0.0f) ? (float)(1.0 / sqrt((double)nuse * (double)grad_bunch)) : 1.0f;
  }
}
void StackANNComponent::computeAllGradients(MatrixDict &grads) {
  prepareGradScales();
  for (auto *c : flat)
    if (c->hasWeightsName()) c->computeAllGradients(grads);
}

ComponentPtr makeHyperplane(const std::string &name, unsigned in, unsigned out, const std::string &dot_name,
                            const std::string &bias_name, const std::string &dot_weights,
                            const std::string &bias_weights) {
  auto s = std::make_shared<StackANNComponent>(name);
  s->input_size = in;
  s->output_size = out;
  s->pushComponent(std::make_shared<DotProductANNComponent>(dot_name, dot_weights, in, out));
  s->pushComponent(std::make_shared<BiasANNComponent>(bias_name, bias_weights, out));
  return s;
}

std::shared_ptr<StackANNComponent> mlpAllAllGenerate(const std::string &topology) {
  std::istringstream is(topology);
  std::vector<std::string> tok;
  std::string t;
  while (is >> t) tok.push_back(t);
  if (tok.size() < 2 || tok[1] != "inputs") throw Error(B200_ERR_BAD_ARG, "'inputs' string is required");
  auto net = std::make_shared<StackANNComponent>("stack");
  unsigned prev = (unsigned)std::stoul(tok[0]);
  net->input_size = prev;
  int count = 1;
  for (size_t i = 2; i + 1 < tok.size(); i += 2) {
    unsigned size = (unsigned)std::stoul(tok[i]);
    const std::string &kind = tok[i + 1];
    std::string c = std::to_string(count);
    net->pushComponent(makeHyperplane("layer" + c, prev, size, "w" + c, "b" + c, "w" + c, "b" + c));
    {
      // parameter defaults of the bindings: leaky_relu leak 0.01, hardtanh [-1, 1]
      const int act = actfFromName(kind);
      const float p0 = act == B200_ACT_LEAKY_RELU ? 0.01f : (act == B200_ACT_HARDTANH ? -1.0f : 0.0f);
      const float p1 = act == B200_ACT_HARDTANH ? 1.0f : 0.0f;
      net->pushComponent(std::make_shared<ActivationFunctionANNComponent>("actf" + c, act, p0, p1));
    }
    prev = size;
    ++count;
  }
  if ((tok.size() % 2) != 0) throw Error(B200_ERR_BAD_ARG, "Incorrect topology string");
  return net;
}

}  // namespace b200
