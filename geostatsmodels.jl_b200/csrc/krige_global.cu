// Global Kriging (fit once on all samples, predict every target): placeholder until the tiled
// DMMA contraction lands; the symbols exist so the ABI is complete and callers get a clean code.
#include "gsm_internal.cuh"

extern "C" {
int gsm_global_krige_fit(gsm_ctx* ctx, const gsm_variogram*, const gsm_model*, gsm_fit** out) {
  if (out) *out = nullptr;
  gsm_set_error(ctx, "global Kriging is not implemented in this build");
  return GSM_ERR_UNSUPPORTED;
}
int gsm_global_krige_status(const gsm_fit* fit) { return fit ? fit->status : 0; }
int gsm_global_krige_predict(gsm_fit*, const gsm_targets*, double*, double*) { return GSM_ERR_UNSUPPORTED; }
int gsm_global_krige_free(gsm_fit* fit) {
  delete fit;
  return GSM_OK;
}
}
