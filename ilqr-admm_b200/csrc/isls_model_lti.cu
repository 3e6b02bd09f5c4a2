// LtiModel<n, m> instantiations (generic constant A, B) of the model-templated kernels (isls_kernels.cuh).  The matrices
// are uploaded into this translation unit's constant-memory block c_lti before every launch sequence.
#include "isls_kernels.cuh"

static int lti_prepare(const isls_plan *plan, cudaStream_t s) {
  LtiConst h;
  memcpy(h.A, plan->lti, sizeof(h.A));
  memcpy(h.B, plan->lti + 36, sizeof(h.B));
  CK(cudaMemcpyToSymbolAsync(c_lti, &h, sizeof(h), 0, cudaMemcpyHostToDevice, s));
  return 0;
}

template <class M>
static const isls_model_ops *with_prepare() {
  static isls_model_ops o = *ModelImpl<M>::ops();
  o.prepare = &lti_prepare;
  return &o;
}

const isls_model_ops *isls_ops_lti(int n, int m) {
  if (n == 2 && m == 1) return with_prepare<LtiModel<2, 1>>();
  if (n == 4 && m == 2) return with_prepare<LtiModel<4, 2>>();
  if (n == 6 && m == 3) return with_prepare<LtiModel<6, 3>>();
  return nullptr;
}
