// TEST INFRASTRUCTURE (oracle/): a flat C face over the reference's own C++
// components and loss functions, compiled in place from /root/reference, so
// that tests/ can run the reference itself beside the numpy restatement
// (oracle/april.py).  Nothing in the product links or loads this.
//
// What is the reference's and what is not:
//   - every forward / backprop / gradient / loss number returned here is
//     computed by the reference's classes (packages/ann/ann/c_src/*.cc,
//     packages/ann/loss/c_src/*.cc, activation_function_kernels.cu,
//     loss_kernels.cu) on its own Matrix / GPUMirroredMemoryBlock types;
//   - BLAS is oracle/ref_build/cblas_min.cc (the reference vendors none);
//   - the trainer and the optimizers are Lua in the reference
//     (trainable/lua_src/supervised.lua, ann/optimizer/lua_src/*.lua) and are
//     NOT reachable from here: those stay pinned by the golden curves.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
#include "lauxlib.h"
#include "lua.h"
#include "lualib.h"
}

#include "MersenneTwister.h"
#include "activation_function_component.h"
#include "ann_component.h"
#include "base.h"
#include "bias_component.h"
#include "convolution_bias_component.h"
#include "convolution_component.h"
#include "cross_entropy_loss_function.h"
#include "dot_product_component.h"
#include "dropout_component.h"
#include "flatten_component.h"
#include "hardtanh_actf_component.h"
#include "hyperplane_component.h"
#include "leaky_relu_actf_component.h"
#include "linear_actf_component.h"
#include "log_logistic_actf_component.h"
#include "log_softmax_actf_component.h"
#include "logistic_actf_component.h"
#include "loss_function.h"
#include "lua_table.h"
#include "matrixFloat.h"
#include "maxpooling_component.h"
#include "mse_loss_function.h"
#include "multiclass_cross_entropy_loss_function.h"
#include "prelu_actf_component.h"
#include "relu_actf_component.h"
#include "rewrap_component.h"
#include "softmax_actf_component.h"
#include "softplus_actf_component.h"
#include "softsign_actf_component.h"
#include "stack_component.h"
#include "tanh_actf_component.h"
#include "token_matrix.h"
#include "zero_one_loss_function.h"

using ANN::ANNComponent;
using AprilUtils::LuaTable;
using Basics::MatrixFloat;
using Basics::Token;
using Basics::TokenMatrixFloat;

namespace {

lua_State* g_L = 0;

// ERROR_EXIT in the reference throws a heap-allocated char* message
// (packages/basics/util/c_src/error_print.cc:132-141); it is caught at this
// boundary and handed to the caller through ref_last_error().
std::string g_error;
const int REF_FAILED = -2147483647 - 1;

#define REF_TRY try {
#define REF_CATCH(failure_value)              \
  }                                           \
  catch (char* msg) {                         \
    g_error = msg ? msg : "unknown error";    \
    delete[] msg;                             \
    return failure_value;                     \
  }                                           \
  catch (const char* msg) {                   \
    g_error = msg ? msg : "unknown error";    \
    return failure_value;                     \
  }

struct Net {
  ANNComponent* root;
  LuaTable* weights;
  LuaTable* components;
  LuaTable* grads;
};

const char* opt_name(const char* s) { return (s && *s) ? s : 0; }

MatrixFloat* matrix_from(const float* data, int ndims, const int* dims) {
  MatrixFloat* m = new MatrixFloat(ndims, dims);
  if (m->getIsContiguous()) {   // a fresh matrix is row-major and dense: one block copy
    std::memcpy(m->getRawDataAccess()->getPPALForWrite() + m->getOffset(), data, sizeof(float) * (size_t)m->size());
    return m;
  }
  const float* p = data;
  for (MatrixFloat::iterator it(m->begin()); it != m->end(); ++it) *it = *p++;
  return m;
}

// Copies m out in logical (last index fastest) order; returns the element count
// and, if dims_out is given, the shape.
int matrix_to(const MatrixFloat* m, float* out, int cap, int* ndims_out, int* dims_out) {
  if (ndims_out) *ndims_out = m->getNumDim();
  if (dims_out)
    for (int i = 0; i < m->getNumDim(); ++i) dims_out[i] = m->getDimSize(i);
  if (out) {
    if (m->size() > cap) return -m->size();
    if (m->getIsContiguous()) {
      std::memcpy(out, m->getRawDataAccess()->getPPALForRead() + m->getOffset(), sizeof(float) * (size_t)m->size());
    } else {
      float* p = out;
      for (MatrixFloat::const_iterator it(m->begin()); it != m->end(); ++it) *p++ = *it;
    }
  }
  return m->size();
}

int token_to(Token* tok, float* out, int cap, int* ndims_out, int* dims_out) {
  if (tok == 0) return 0;
  if (tok->getTokenCode() != Basics::table_of_token_codes::token_matrix) return 0;
  return matrix_to(tok->convertTo<TokenMatrixFloat*>()->getMatrix(), out, cap, ndims_out, dims_out);
}

}  // namespace

extern "C" {

// Once per process: the reference keeps one global lua_State for its tables
// (packages/basics/base/c_src/base.cc:28-45).
int ref_init(void) {
  if (g_L) return 0;
  g_L = luaL_newstate();
  if (!g_L) return 1;
  luaL_openlibs(g_L);
  Base::registerGlobalLuaState(g_L);
  return 0;
}

// 1 when this library is the reference's USE_CUDA build (integration/Makefile).
int ref_built_with_cuda(void) {
#ifdef USE_CUDA
  return 1;
#else
  return 0;
#endif
}
// mathcore.set_use_cuda_default (mathcore/binding/bind_mathcore.lua.cc:157-163):
// matrices and components created from now on compute on the device.
void ref_set_use_cuda_default(int flag) {
  AprilMath::GPUMirroredMemoryBlockBase::USE_CUDA_DEFAULT = (flag != 0);
}
// component:set_use_cuda(flag) (ann_component.h:378-388); recursive for a stack.
void ref_component_set_use_cuda(void* comp, int flag) {
  static_cast<ANNComponent*>(comp)->setUseCuda(flag != 0);
}

// ---- components --------------------------------------------------------
// Each constructor returns a new ANNComponent* with one reference held by the
// caller; ref_stack_push hands a second one to the stack.

void* ref_stack_new(void) {
  ANNComponent* c = new ANN::StackANNComponent();
  IncRef(c);
  return c;
}
void ref_stack_push(void* stack, void* comp) {
  static_cast<ANN::StackANNComponent*>(stack)->pushComponent(static_cast<ANNComponent*>(comp));
}
void* ref_hyperplane_new(int in, int out, const char* wname, const char* bname, int transpose) {
  ANNComponent* c = new ANN::HyperplaneANNComponent(0, 0, 0, wname, bname, in, out, transpose != 0);
  IncRef(c);
  return c;
}
void* ref_dot_product_new(int in, int out, const char* wname, int transpose) {
  ANNComponent* c = new ANN::DotProductANNComponent(0, wname, in, out, transpose != 0);
  IncRef(c);
  return c;
}
void* ref_bias_new(int size, const char* wname) {
  ANNComponent* c = new ANN::BiasANNComponent(size, 0, wname);
  IncRef(c);
  return c;
}
void* ref_actf_new(const char* kind, float p0, float p1) {
  std::string k(kind);
  ANNComponent* c = 0;
  if (k == "logistic") c = new ANN::LogisticActfANNComponent(0);
  else if (k == "tanh") c = new ANN::TanhActfANNComponent(0);
  else if (k == "relu") c = new ANN::ReLUActfANNComponent(0);
  else if (k == "softmax") c = new ANN::SoftmaxActfANNComponent(0);
  else if (k == "log_softmax") c = new ANN::LogSoftmaxActfANNComponent(0);
  else if (k == "linear") c = new ANN::LinearActfANNComponent(0);
  else if (k == "log_logistic") c = new ANN::LogLogisticActfANNComponent(0);
  else if (k == "softplus") c = new ANN::SoftplusActfANNComponent(0);
  else if (k == "softsign") c = new ANN::SoftsignActfANNComponent(0);
  else if (k == "leaky_relu") c = new ANN::LeakyReLUActfANNComponent(p0, 0);
  else if (k == "hardtanh") c = new ANN::HardtanhActfANNComponent(0, p0, p1);
  if (c) IncRef(c);
  return c;
}
void* ref_prelu_new(int scalar, int size, const char* wname) {
  ANNComponent* c = new ANN::PReLUActfANNComponent(scalar != 0, size, 0, wname);
  IncRef(c);
  return c;
}
void* ref_convolution_new(int ndims, const int* kernel, const int* step, int n, const char* wname) {
  ANNComponent* c = new ANN::ConvolutionANNComponent(ndims, kernel, step, n, 0, wname);
  IncRef(c);
  return c;
}
void* ref_convolution_bias_new(int ndims, int n, const char* wname) {
  ANNComponent* c = new ANN::ConvolutionBiasANNComponent(ndims, n, 0, wname);
  IncRef(c);
  return c;
}
void* ref_max_pooling_new(int ndims, const int* kernel, const int* step) {
  ANNComponent* c = new ANN::MaxPoolingANNComponent(ndims, kernel, step, 0);
  IncRef(c);
  return c;
}
void* ref_flatten_new(void) {
  ANNComponent* c = new ANN::FlattenANNComponent(0);
  IncRef(c);
  return c;
}
void* ref_rewrap_new(const int* dims, int n) {
  ANNComponent* c = new ANN::RewrapANNComponent(dims, n, 0);
  IncRef(c);
  return c;
}
void* ref_dropout_new(unsigned int seed, float value, float prob, int normalize) {
  Basics::MTRand* rng = new Basics::MTRand(seed);
  ANNComponent* c = new ANN::DropoutANNComponent(rng, value, prob, normalize != 0, 0, 0);
  IncRef(c);
  return c;
}
void ref_component_free(void* comp) { DecRef(static_cast<ANNComponent*>(comp)); }

// ---- a built network ---------------------------------------------------

const char* ref_last_error(void) { return g_error.c_str(); }

void* ref_net_build(void* root, int input_size, int output_size) {
  REF_TRY
  Net* n = new Net;
  n->root = static_cast<ANNComponent*>(root);
  IncRef(n->root);
  n->weights = new LuaTable(g_L);
  n->components = new LuaTable(g_L);
  n->grads = new LuaTable(g_L);
  n->root->build(input_size, output_size, *n->weights, *n->components);
  return n;
  REF_CATCH((void*)0)
}
void ref_net_free(void* net) {
  Net* n = static_cast<Net*>(net);
  delete n->grads;
  delete n->components;
  delete n->weights;
  DecRef(n->root);
  delete n;
}

// which: 0 = weights, 1 = gradients.  Returns the element count (0 if absent,
// negative required size if cap is too small).
int ref_net_tensor_get(void* net, int which, const char* name, float* out, int cap,
                       int* ndims_out, int* dims_out) {
  Net* n = static_cast<Net*>(net);
  LuaTable& t = which ? *n->grads : *n->weights;
  MatrixFloat* m = t.opt<MatrixFloat*>(name, 0);
  if (!m) return 0;
  return matrix_to(m, out, cap, ndims_out, dims_out);
}
int ref_net_weight_set(void* net, const char* name, const float* data, int count) {
  Net* n = static_cast<Net*>(net);
  MatrixFloat* m = n->weights->opt<MatrixFloat*>(name, 0);
  if (!m || m->size() != count) return 1;
  const float* p = data;
  for (MatrixFloat::iterator it(m->begin()); it != m->end(); ++it) *it = *p++;
  return 0;
}
// Number of components that share the tensor, as the trainer reads it for the
// gradient smoothing (trainable/lua_src/supervised.lua:797-803).
int ref_net_shared_count(void* net, const char* name) {
  Net* n = static_cast<Net*>(net);
  MatrixFloat* m = n->weights->opt<MatrixFloat*>(name, 0);
  return m ? (int)m->getSharedCount() : -1;
}

int ref_net_forward(void* net, const float* x, int ndims, const int* dims, int during_training,
                    float* out, int cap, int* ndims_out, int* dims_out) {
  REF_TRY
  Net* n = static_cast<Net*>(net);
  MatrixFloat* m = matrix_from(x, ndims, dims);
  Token* in = new TokenMatrixFloat(m);
  IncRef(in);
  Token* y = n->root->doForward(in, during_training != 0);
  int r = token_to(y, out, cap, ndims_out, dims_out);
  DecRef(in);
  return r;
  REF_CATCH(REF_FAILED)
}
int ref_net_backprop(void* net, const float* e, int ndims, const int* dims, float* out, int cap,
                     int* ndims_out, int* dims_out) {
  REF_TRY
  Net* n = static_cast<Net*>(net);
  MatrixFloat* m = matrix_from(e, ndims, dims);
  Token* in = new TokenMatrixFloat(m);
  IncRef(in);
  Token* d = n->root->doBackprop(in);
  int r = token_to(d, out, cap, ndims_out, dims_out);
  DecRef(in);
  return r;
  REF_CATCH(REF_FAILED)
}
// Gradients are accumulated by the components into zeroed matrices on the
// first call after a reset, like the trainer's loop does.
int ref_net_compute_gradients(void* net) {
  REF_TRY
  Net* n = static_cast<Net*>(net);
  n->root->computeAllGradients(*n->grads);
  return 0;
  REF_CATCH(REF_FAILED)
}
void ref_net_reset(void* net, unsigned int it) { static_cast<Net*>(net)->root->reset(it); }

// ---- loss functions ----------------------------------------------------

void* ref_loss_new(const char* kind, int size, float param) {
  std::string k(kind);
  ANN::LossFunction* l = 0;
  if (k == "mse") l = new ANN::MSELossFunction(size);
  else if (k == "multi_class_cross_entropy") l = new ANN::MultiClassCrossEntropyLossFunction(size);
  else if (k == "cross_entropy") l = new ANN::CrossEntropyLossFunction(size);
  else if (k == "zero_one") l = new ANN::ZeroOneLossFunction(size, param);
  if (l) IncRef(l);
  return l;
}
void ref_loss_free(void* loss) { DecRef(static_cast<ANN::LossFunction*>(loss)); }

// Per-row losses into loss_vec[rows]; returns 0 on success.
int ref_loss_compute(void* loss, const float* out, const float* tgt, int rows, int cols,
                     int tgt_cols, float* loss_vec) {
  REF_TRY
  ANN::LossFunction* l = static_cast<ANN::LossFunction*>(loss);
  int od[2] = {rows, cols}, td[2] = {rows, tgt_cols};
  Token* o = new TokenMatrixFloat(matrix_from(out, 2, od));
  Token* t = new TokenMatrixFloat(matrix_from(tgt, 2, td));
  IncRef(o);
  IncRef(t);
  MatrixFloat* v = l->computeLoss(o, t);
  int rc = 1;
  if (v) {
    IncRef(v);
    rc = matrix_to(v, loss_vec, rows, 0, 0) == rows ? 0 : 2;
    DecRef(v);
  }
  DecRef(o);
  DecRef(t);
  return rc;
  REF_CATCH(REF_FAILED)
}
int ref_loss_gradient(void* loss, const float* out, const float* tgt, int rows, int cols,
                      float* grad) {
  REF_TRY
  ANN::LossFunction* l = static_cast<ANN::LossFunction*>(loss);
  int od[2] = {rows, cols};
  Token* o = new TokenMatrixFloat(matrix_from(out, 2, od));
  Token* t = new TokenMatrixFloat(matrix_from(tgt, 2, od));
  IncRef(o);
  IncRef(t);
  Token* g = l->computeGradient(o, t);
  int rc = token_to(g, grad, rows * cols, 0, 0) == rows * cols ? 0 : 1;
  DecRef(o);
  DecRef(t);
  l->reset();
  return rc;
  REF_CATCH(REF_FAILED)
}

// ---- the reference's MT19937 stream (dropout masks, shuffles) -----------

void* ref_random_new(unsigned int seed) {
  Basics::MTRand* r = new Basics::MTRand(seed);
  IncRef(r);
  return r;
}
void ref_random_free(void* r) { DecRef(static_cast<Basics::MTRand*>(r)); }
double ref_random_rand(void* r) { return static_cast<Basics::MTRand*>(r)->rand(); }
unsigned int ref_random_randint(void* r) { return static_cast<Basics::MTRand*>(r)->randInt(); }
double ref_random_rand_n(void* r, double n) { return static_cast<Basics::MTRand*>(r)->rand(n); }
unsigned int ref_random_randint_n(void* r, unsigned int n) { return static_cast<Basics::MTRand*>(r)->randInt(n); }
// the permutation trainable.supervised_trainer draws for an epoch (0-based here; the Lua binding adds 1)
void ref_random_shuffle(void* r, int size, int* out) { static_cast<Basics::MTRand*>(r)->shuffle(size, out); }

// ---- raw BLAS-level seam (AprilMath::doGemm through the Matrix API) ------

}  // extern "C"

#include "matrix_ext.h"

extern "C" {

// C[m,n] = alpha * op(A) * op(B) + beta * C through MatrixExt::BLAS::matGemm
// (packages/basics/matrix/c_src/matrix_ext_blas.cu:150-245), row-major.
int ref_gemm(int ta, int tb, int m, int n, int k, float alpha, const float* a, const float* b,
             float beta, float* c) {
  REF_TRY
  int ad[2] = {ta ? k : m, ta ? m : k}, bd[2] = {tb ? n : k, tb ? k : n}, cd[2] = {m, n};
  MatrixFloat* A = matrix_from(a, 2, ad);
  MatrixFloat* B = matrix_from(b, 2, bd);
  MatrixFloat* C = matrix_from(c, 2, cd);
  IncRef(A);
  IncRef(B);
  IncRef(C);
  AprilMath::MatrixExt::BLAS::matGemm(C, ta ? CblasTrans : CblasNoTrans,
                                      tb ? CblasTrans : CblasNoTrans, alpha, A, B, beta);
  matrix_to(C, c, m * n, 0, 0);
  DecRef(A);
  DecRef(B);
  DecRef(C);
  return 0;
  REF_CATCH(REF_FAILED)
}

}  // extern "C"
