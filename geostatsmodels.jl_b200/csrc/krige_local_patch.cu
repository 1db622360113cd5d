// Patch-level local Kriging (krige_local_patch.cuh): dispatcher over the instantiated constraint counts.
#include "krige_local_patch.cuh"

int gsm_build_jobs(gsm_ctx* ctx, const TargetsDev& tg, int k, int nbc, int smax, int minneighbors, const int* d_pos,
                   const int* d_cnt, const JobIdx** out_jobs, long long* out_ngroups);

int gsm_launch_local_krige_patch(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered,
                                 const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                                 double* d_mean, double* d_var, unsigned char* d_status) {
  if (k > 32 || k < 1) return GSM_ERR_UNSUPPORTED;
  if (tg.count <= 0) return GSM_OK;
  if (tg.count >= (1LL << 31)) return GSM_ERR_UNSUPPORTED;   // descriptors hold 32-bit relative indices
  switch (ncon) {
    case 0: case 1: case 2: case 3: case 4: case 6: case 10: break;
    default: return GSM_ERR_UNSUPPORTED;
  }
  const int nbc = (k + 7) / 8;
  const bool d3 = ctx->dim >= 3;
  const JobIdx* jobs = nullptr;
  long long ngroups = 0;
  GSM_CHECK(gsm_build_jobs(ctx, tg, k, nbc, std::min(nbc - 1, 3), minneighbors, d_pos, d_cnt, &jobs, &ngroups));
#define GSM_PATCH_ARGS ctx, v, m, tg, jobs, ngroups, k, nbc, centered, d_pos, d_mean, d_var, d_status
#define GSM_PATCH_CASE(N)                                                                                  \
  case N:                                                                                                  \
    if (two) return d3 ? gsm_patch::launch_patch<N, true, 2>(GSM_PATCH_ARGS) : gsm_patch::launch_patch<N, false, 2>(GSM_PATCH_ARGS); \
    return d3 ? gsm_patch::launch_patch<N, true, 1>(GSM_PATCH_ARGS) : gsm_patch::launch_patch<N, false, 1>(GSM_PATCH_ARGS);
  const bool two = ctx->opt_local_impl == GSM_IMPL_PATCH2;   // developer option: two warps per patch
  switch (ncon) {
    GSM_PATCH_CASE(0)
    GSM_PATCH_CASE(1)
    GSM_PATCH_CASE(2)
    GSM_PATCH_CASE(3)
    GSM_PATCH_CASE(4)
    GSM_PATCH_CASE(6)
    GSM_PATCH_CASE(10)
  }
#undef GSM_PATCH_CASE
  return GSM_ERR_UNSUPPORTED;
}
