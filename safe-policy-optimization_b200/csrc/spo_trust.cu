// Trust-region pieces (CPO / TRPO-Lag): surrogate gradient, Fisher-vector product,
// line-search evaluation, conjugate gradient.  (implemented in the next milestone)
#include "spo_common.cuh"

extern "C" {

int spo_surrogate_grad(const spo_dims*, const float*, const float*, const float*, const float*, const float*, int64_t,
                       float*, float*, void*) {
  spo_set_error("spo_surrogate_grad: not built yet");
  return SPO_ERR_UNSUPPORTED;
}
int spo_fvp(const spo_dims*, const float*, const float*, int64_t, const float*, float, float*, void*) {
  spo_set_error("spo_fvp: not built yet");
  return SPO_ERR_UNSUPPORTED;
}
int spo_linesearch_eval(const spo_dims*, const float*, const float*, const float*, const float*, const float*, const float*,
                        const float*, const float*, int64_t, float*, void*) {
  spo_set_error("spo_linesearch_eval: not built yet");
  return SPO_ERR_UNSUPPORTED;
}
int spo_conjugate_gradient(const spo_dims*, const float*, const float*, int64_t, const float*, int, float, float, float,
                           float*, float*, void*) {
  spo_set_error("spo_conjugate_gradient: not built yet");
  return SPO_ERR_UNSUPPORTED;
}
}
