// C face of the integration-only classes, next to oracle/ref_build/ref_api.cc
// (same conventions: the caller holds one reference on what a constructor returns).
#include "b200_dot_product_component.h"

extern "C" {

// ann.components.dot_product with its dense methods on libb200ann.so
// (integration/ann/b200_dot_product_component.h).
void* ref_b200_dot_product_new(int in, int out, const char* wname, int transpose) {
  ANN::ANNComponent* c = new ANN::B200DotProductANNComponent(0, wname, in, out, transpose != 0);
  IncRef(c);
  return c;
}

}  // extern "C"
