// Dispatcher of the pipelined local-Kriging kernel (krige_local_pipe.cuh) over its instantiation units.
#include "gsm_internal.cuh"

int gsm_pipe_entry_d0_b1(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, int, int, const int*, const int*, double*, double*, unsigned char*);
int gsm_pipe_entry_d0_b2(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, int, int, const int*, const int*, double*, double*, unsigned char*);
int gsm_pipe_entry_d0_b4(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, int, int, const int*, const int*, double*, double*, unsigned char*);
int gsm_pipe_entry_d1_b1(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, int, int, const int*, const int*, double*, double*, unsigned char*);
int gsm_pipe_entry_d1_b2(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, int, int, const int*, const int*, double*, double*, unsigned char*);
int gsm_pipe_entry_d1_b4(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, int, int, const int*, const int*, double*, double*, unsigned char*);

int gsm_launch_local_krige_pipe(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered,
                                const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                                double* d_mean, double* d_var, unsigned char* d_status) {
  if (k > 32 || k < 1) return GSM_ERR_UNSUPPORTED;
  const int nbc = k <= 8 ? 1 : (k <= 16 ? 2 : 4);
  const bool d3 = ctx->dim >= 3;
#define GSM_PIPE_CALL(D, B) gsm_pipe_entry_d##D##_b##B(ctx, v, m, ncon, centered, tg, k, minneighbors, d_pos, d_cnt, d_mean, d_var, d_status)
  if (!d3) return nbc == 1 ? GSM_PIPE_CALL(0, 1) : (nbc == 2 ? GSM_PIPE_CALL(0, 2) : GSM_PIPE_CALL(0, 4));
  return nbc == 1 ? GSM_PIPE_CALL(1, 1) : (nbc == 2 ? GSM_PIPE_CALL(1, 2) : GSM_PIPE_CALL(1, 4));
#undef GSM_PIPE_CALL
}
