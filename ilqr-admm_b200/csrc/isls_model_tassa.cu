// TassaCarModel instantiation of the model-templated kernels (isls_kernels.cuh)
#include "isls_kernels.cuh"
const isls_model_ops *isls_ops_tassa_car() { return ModelImpl<TassaCarModel>::ops(); }
