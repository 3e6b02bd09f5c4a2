// DoubleIntModel<1..3> instantiations of the model-templated kernels (isls_kernels.cuh)
#include "isls_kernels.cuh"
const isls_model_ops *isls_ops_double_integrator(int m) {
  if (m == 1) return ModelImpl<DoubleIntModel<1>>::ops();
  if (m == 2) return ModelImpl<DoubleIntModel<2>>::ops();
  if (m == 3) return ModelImpl<DoubleIntModel<3>>::ops();
  return nullptr;
}
