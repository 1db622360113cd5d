// Do DMMA and DFMA share one pipe on B200?  Time DFMA-only, DMMA-only and an interleaved mix.
// Also print how the warps of 288-thread CTAs are placed on hardware warp slots (%warpid) per SM.
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 4096
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NF, int NM>
__global__ void mix(double* out, double s) {
  double f[8], c[8][2];
  for (int i = 0; i < 8; ++i) { f[i] = s + i + threadIdx.x; c[i][0] = c[i][1] = 0; }
  double a = s * 1e-3 + threadIdx.x * 1e-6;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NM) dmma884(c[i][0], c[i][1], a, 1.0);
#pragma unroll
      for (int j = 0; j < NF; ++j) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[(i + j) & 7]) : "d"(1.0000001), "d"(1e-9));
    }
  }
  double r = 0; for (int i = 0; i < 8; ++i) r += f[i] + c[i][0] + c[i][1];
  if (r == 1.2345) out[0] = r;
}
__global__ void place(int* out) {
  if ((threadIdx.x & 31) == 0) {
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    int w = threadIdx.x >> 5;
    out[(blockIdx.x * 16 + w) * 2] = smid;
    out[(blockIdx.x * 16 + w) * 2 + 1] = warpid;
  }
  // keep CTAs resident long enough that two share an SM
  long long t0 = clock64(); while (clock64() - t0 < 200000) {}
}
template <int NF, int NM> float run(double* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); mix<NF, NM><<<148 * 8, 256>>>(d, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  return best;
}
int main() {
  double* d; cudaMalloc(&d, 64);
  float t_f = run<8, 0>(d);   // 64 DFMA per iter
  float t_m = run<0, 8>(d);   // 8 DMMA per iter
  float t_x = run<8, 8>(d);   // both
  printf("DFMA only (64/iter): %.3f ms\nDMMA only (8/iter): %.3f ms\nboth: %.3f ms  (sum %.3f, max %.3f)\n", t_f, t_m, t_x, t_f + t_m, t_f > t_m ? t_f : t_m);
  int* p; cudaMalloc(&p, 296 * 16 * 2 * 4); cudaMemset(p, 0xff, 296 * 16 * 2 * 4);
  place<<<296, 288>>>(p); cudaDeviceSynchronize();
  static int h[296 * 16 * 2]; cudaMemcpy(h, p, sizeof(h), cudaMemcpyDeviceToHost);
  for (int b = 0; b < 296; ++b) if (h[b * 32] == 0 || h[b*32] == 5) { printf("cta %d sm %d warpids:", b, h[b * 32]); for (int w = 0; w < 9; ++w) printf(" %d", h[(b * 16 + w) * 2 + 1]); printf("\n"); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
