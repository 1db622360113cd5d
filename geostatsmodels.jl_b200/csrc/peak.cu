// FP64 roofline denominators measured on the device the context owns: a register-resident DFMA chain
// (vector FP64 pipe) and a DMMA m8n8k4 chain (FP64 tensor path).  MEASURED_PEAKS.json records only HBM
// and bf16, so bench.py calls this in the same run that produces the Kriging numbers.
#include "gsm_internal.cuh"

namespace {

constexpr int PK_ITERS = 4096;
constexpr int PK_ILP = 8;

__global__ void __launch_bounds__(256) dfma_kernel(double* out, double seed) {
  double a[PK_ILP];
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) a[i] = seed + i + threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < PK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) s += a[i];
  if (s == 123.456) out[0] = s;  // keep the chain alive
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, double seed) {
  double acc[PK_ILP][2];
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) acc[i][0] = acc[i][1] = 0.0;
  const double a = seed * 1e-3 + threadIdx.x * 1e-6, b = 1.0 + seed * 1e-6;
  for (int it = 0; it < PK_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < PK_ILP; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < PK_ILP; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}

}  // namespace

int gsm_fp64_peak(gsm_ctx* ctx, double* dfma, double* dmma) {
  GSM_CHECK(gsm_reserve(ctx, ctx->red, 64));
  cudaStream_t st = ctx->stream;
  cudaEvent_t e0, e1;
  GSM_CUDA(ctx, cudaEventCreate(&e0));
  GSM_CUDA(ctx, cudaEventCreate(&e1));
  const int blocks = 148 * 8, threads = 256;
  double best[2] = {0, 0};
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 6; ++rep) {
      GSM_CUDA(ctx, cudaEventRecord(e0, st));
      if (which == 0) dfma_kernel<<<blocks, threads, 0, st>>>((double*)ctx->red.p, 1.0);
      else dmma_kernel<<<blocks, threads, 0, st>>>((double*)ctx->red.p, 1.0);
      GSM_CUDA(ctx, cudaEventRecord(e1, st));
      GSM_CUDA(ctx, cudaStreamSynchronize(st));
      GSM_CUDA(ctx, cudaGetLastError());
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      double flops = which == 0 ? 2.0 * blocks * threads * (double)PK_ITERS * PK_ILP
                                : 2.0 * 256.0 * blocks * (threads / 32) * (double)PK_ITERS * PK_ILP;
      if (rep > 0) best[which] = std::max(best[which], flops / (ms * 1e-3) / 1e12);
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (dfma) *dfma = best[0];
  if (dmma) *dmma = best[1];
  return GSM_OK;
}
