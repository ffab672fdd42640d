// See b200_dot_product_component.h.
#include "b200_dot_product_component.h"

#include "b200_bridge.h"
#include "matrixFloat.h"

using Basics::MatrixFloat;

namespace ANN {

  B200DotProductANNComponent::B200DotProductANNComponent(const char *name, const char *weights_name,
                                                         unsigned int input_size,
                                                         unsigned int output_size,
                                                         bool transpose_weights, MatrixFloat *matrix) :
    DotProductANNComponent(name, weights_name, input_size, output_size, transpose_weights, matrix),
    b200_weights(0) {
  }

  B200DotProductANNComponent::~B200DotProductANNComponent() {
    if (b200_weights) DecRef(b200_weights);
  }

  ANNComponent *B200DotProductANNComponent::clone(AprilUtils::LuaTable &copies) {
    UNUSED_VARIABLE(copies);
    return new B200DotProductANNComponent(getName().c_str(), getWeightsName().c_str(),
                                          getInputSize(), getOutputSize(), transposed());
  }

  void B200DotProductANNComponent::build(unsigned int input_size, unsigned int output_size,
                                         AprilUtils::LuaTable &weights_dict,
                                         AprilUtils::LuaTable &components_dict) {
    DotProductANNComponent::build(input_size, output_size, weights_dict, components_dict);
    MatrixFloat *w = weights_dict.get<MatrixFloat*>(getWeightsName());
    AssignRef(b200_weights, w);
  }

#ifdef USE_B200

  namespace {
    inline bool plain2d(const MatrixFloat *m) {
      return m->getNumDim() == 2 && m->getStrideSize(1) == 1 && m->getStrideSize(0) >= m->getDimSize(1);
    }
  }

  bool B200DotProductANNComponent::onDevice(const MatrixFloat *a, const MatrixFloat *b) {
    return getUseCuda() && !transposed() && b200_weights != 0 && plain2d(b200_weights) &&
      plain2d(a) && (b == 0 || plain2d(b)) && a->getDimSize(0) > 1;
  }

  // Y[bunch, out] = X[bunch, in] . W[out, in]^T
  MatrixFloat *B200DotProductANNComponent::privateDoDenseForward(MatrixFloat *input_mat,
                                                                 bool during_training) {
    if (!onDevice(input_mat, 0))
      return DotProductANNComponent::privateDoDenseForward(input_mat, during_training);
    const int bunch = input_mat->getDimSize(0);
    int dims[2] = { bunch, static_cast<int>(getOutputSize()) };
    MatrixFloat *output_mat = new MatrixFloat(2, dims);
    output_mat->setUseCuda(true);
    const float *x = input_mat->getRawDataAccess()->getGPUForRead() + input_mat->getOffset();
    const float *w = b200_weights->getRawDataAccess()->getGPUForRead() + b200_weights->getOffset();
    float *y = output_mat->getRawDataAccess()->getGPUForWrite() + output_mat->getOffset();
    AprilMath::B200::StreamOrder order_guard;
    AprilMath::B200::check(b200_linear_fwd(AprilMath::B200::context(), bunch,
                                           static_cast<int>(getOutputSize()),
                                           static_cast<int>(getInputSize()),
                                           x, input_mat->getStrideSize(0),
                                           w, b200_weights->getStrideSize(0),
                                           /*bias*/ 0, B200_ACT_NONE,
                                           y, output_mat->getStrideSize(0)));
    return output_mat;
  }

  // dX[bunch, in] = dY[bunch, out] . W[out, in]
  MatrixFloat *B200DotProductANNComponent::privateDoDenseBackprop(MatrixFloat *error_input_mat) {
    if (!onDevice(error_input_mat, 0))
      return DotProductANNComponent::privateDoDenseBackprop(error_input_mat);
    const int bunch = error_input_mat->getDimSize(0);
    int dims[2] = { bunch, static_cast<int>(getInputSize()) };
    MatrixFloat *error_output_mat = new MatrixFloat(2, dims);
    error_output_mat->setUseCuda(true);
    const float *dy = error_input_mat->getRawDataAccess()->getGPUForRead() + error_input_mat->getOffset();
    const float *w = b200_weights->getRawDataAccess()->getGPUForRead() + b200_weights->getOffset();
    float *dx = error_output_mat->getRawDataAccess()->getGPUForWrite() + error_output_mat->getOffset();
    AprilMath::B200::StreamOrder order_guard;
    AprilMath::B200::check(b200_linear_bwd_data(AprilMath::B200::context(), bunch,
                                                static_cast<int>(getOutputSize()),
                                                static_cast<int>(getInputSize()),
                                                dy, error_input_mat->getStrideSize(0),
                                                w, b200_weights->getStrideSize(0),
                                                B200_ACT_NONE, 0, 0,
                                                dx, error_output_mat->getStrideSize(0)));
    return error_output_mat;
  }

  // dW[out, in] += dY[bunch, out]^T . X[bunch, in]   (the trainer applies its 1/sqrt(N bunch) afterwards)
  void B200DotProductANNComponent::privateDenseComputeGradients(const char *name,
                                                                AprilUtils::LuaTable &grads_mat_dict) {
    MatrixFloat *error_input_mat = getErrorInputMatrix();
    MatrixFloat *input_mat = getInputMatrix();
    if (!onDevice(input_mat, error_input_mat)) {
      DotProductANNComponent::privateDenseComputeGradients(name, grads_mat_dict);
      return;
    }
    // shared count, allocation and zeroing of a new gradient matrix: the reference's helper (:170-190)
    MatrixFloat *grads_mat = initializeComputeGradients(name, grads_mat_dict);
    if (!plain2d(grads_mat)) ERROR_EXIT(128, "b200: gradient matrix is not a plain row-major matrix\n");
    const int bunch = error_input_mat->getDimSize(0);
    const float *dy = error_input_mat->getRawDataAccess()->getGPUForRead() + error_input_mat->getOffset();
    const float *x = input_mat->getRawDataAccess()->getGPUForRead() + input_mat->getOffset();
    float *dw = grads_mat->getRawDataAccess()->getGPUForReadAndWrite() + grads_mat->getOffset();
    AprilMath::B200::StreamOrder order_guard;
    AprilMath::B200::check(b200_linear_bwd_weight(AprilMath::B200::context(), bunch,
                                                  static_cast<int>(getOutputSize()),
                                                  static_cast<int>(getInputSize()),
                                                  dy, error_input_mat->getStrideSize(0),
                                                  x, input_mat->getStrideSize(0),
                                                  /*scale*/ 1.0f, /*beta*/ 1.0f,
                                                  dw, grads_mat->getStrideSize(0), /*db*/ 0));
  }

#else // !USE_B200: the class is the reference's component under another name

  bool B200DotProductANNComponent::onDevice(const MatrixFloat *, const MatrixFloat *) {
    return false;
  }
  MatrixFloat *B200DotProductANNComponent::privateDoDenseForward(MatrixFloat *input_mat,
                                                                 bool during_training) {
    return DotProductANNComponent::privateDoDenseForward(input_mat, during_training);
  }
  MatrixFloat *B200DotProductANNComponent::privateDoDenseBackprop(MatrixFloat *error_input_mat) {
    return DotProductANNComponent::privateDoDenseBackprop(error_input_mat);
  }
  void B200DotProductANNComponent::privateDenseComputeGradients(const char *name,
                                                                AprilUtils::LuaTable &grads_mat_dict) {
    DotProductANNComponent::privateDenseComputeGradients(name, grads_mat_dict);
  }

#endif

} // namespace ANN
