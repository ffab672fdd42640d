// Component seam of the integration: the reference's
// dot-product component with its three dense methods routed to the fused
// dense-layer entry points of libb200ann.so.
//
// A maintainer adds this file pair to packages/ann/ann/c_src/ and registers the
// class next to ann.components.dot_product
// (packages/ann/ann/binding/bind_ann_base.lua.cc); everything the class does not
// override -- build, weight registration, sparse inputs, clone, serialisation --
// is the reference's DotProductANNComponent
// (packages/ann/ann/c_src/dot_product_component.{h,cc}).
#ifndef B200_DOT_PRODUCT_COMPONENT_H
#define B200_DOT_PRODUCT_COMPONENT_H

#include "dot_product_component.h"

namespace ANN {

  class B200DotProductANNComponent : public DotProductANNComponent {
    APRIL_DISALLOW_COPY_AND_ASSIGN(B200DotProductANNComponent);

    /// The weights as registered in the dictionary at build(): the parent keeps
    /// its own pointer private.
    Basics::MatrixFloat *b200_weights;

    /// True when this call can go to the library: use_cuda is on, the weights
    /// are [out, in] (not stored transposed) and every operand is a plain
    /// row-major 2-D matrix (unit stride along rows).
    bool onDevice(const Basics::MatrixFloat *a, const Basics::MatrixFloat *b);

  protected:
    // dot_product_component.cc:63-98
    virtual Basics::MatrixFloat *privateDoDenseForward(Basics::MatrixFloat *input,
                                                       bool during_training);
    // dot_product_component.cc:123-152
    virtual Basics::MatrixFloat *privateDoDenseBackprop(Basics::MatrixFloat *error_input);
    // dot_product_component.cc:194-216
    virtual void privateDenseComputeGradients(const char *name,
                                              AprilUtils::LuaTable &grads_mat_dict);

  public:
    B200DotProductANNComponent(const char *name = 0, const char *weights_name = 0,
                               unsigned int input_size = 0, unsigned int output_size = 0,
                               bool transpose_weights = false,
                               Basics::MatrixFloat *matrix = 0);
    virtual ~B200DotProductANNComponent();
    virtual ANNComponent *clone(AprilUtils::LuaTable &copies);
    virtual void build(unsigned int input_size, unsigned int output_size,
                       AprilUtils::LuaTable &weights_dict,
                       AprilUtils::LuaTable &components_dict);
  };

} // namespace ANN

#endif // B200_DOT_PRODUCT_COMPONENT_H
