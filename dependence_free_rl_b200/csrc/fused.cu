// fused.cu -- fused small-MLP rollout / learn kernels (placeholder: not attached yet; every
// trainer currently runs on the layered kernels of layers.cu).
#include "trainer.h"

int dfrl_fused_try_attach(dfrl_trainer *t) {
  t->fused_impl = nullptr;
  return DFRL_ERR_UNSUPPORTED;
}
void dfrl_fused_detach(dfrl_trainer *t) { t->fused_impl = nullptr; }
int dfrl_fused_rollout(dfrl_trainer *, const uint8_t *, const uint8_t *, const double *) {
  return DFRL_ERR_UNSUPPORTED;
}
int dfrl_fused_learn(dfrl_trainer *) { return DFRL_ERR_UNSUPPORTED; }
int dfrl_fused_eval_argmax(dfrl_ctx *, dfrl_env *, dfrl_mlp *, int, double *, long long *) {
  return DFRL_ERR_UNSUPPORTED;
}
