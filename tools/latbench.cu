// Single-warp dependent-chain latencies on B200 (cycles per op), clock64-timed.
#include <cuda_runtime.h>
#include <cstdio>
#define N 2048
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-(d * y), y, 1.0);
  const double p = e * fma(e, 0.375, 0.5);
  return fma(y, p, y);
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* cyc, double s) {
  __shared__ double sm[64];
  sm[threadIdx.x] = threadIdx.x * 1e-9;
  sm[threadIdx.x + 32] = 0;
  __syncwarp();
  long long t0, t1; double a = s + threadIdx.x;
  // 0: DFMA chain
  t0 = clock64(); for (int i = 0; i < N; ++i) a = fma(a, 1.0000001, 1e-9); t1 = clock64(); cyc[0] = t1 - t0;
  // 1: DMUL chain
  t0 = clock64(); for (int i = 0; i < N; ++i) a = a * 1.0000001; t1 = clock64(); cyc[1] = t1 - t0;
  // 2: rsqrt() chain
  t0 = clock64(); for (int i = 0; i < N; ++i) a = rsqrt(a) + 1.5; t1 = clock64(); cyc[2] = t1 - t0;
  // 3: fast_rsqrt chain
  t0 = clock64(); for (int i = 0; i < N; ++i) a = fast_rsqrt(a) + 1.5; t1 = clock64(); cyc[3] = t1 - t0;
  // 4: mufu.rsq64h alone chain
  t0 = clock64(); for (int i = 0; i < N; ++i) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = y; } t1 = clock64(); cyc[4] = t1 - t0;
  // 5: LDS chain (dependent address)
  int idx = threadIdx.x & 31; double acc = 0;
  t0 = clock64(); for (int i = 0; i < N; ++i) { double v = sm[idx]; acc += v; idx = (idx + (int)(v * 1e-30)) & 31; } t1 = clock64(); cyc[5] = t1 - t0; a += acc;
  // 6: DMMA chain
  double d0 = 0, d1 = 0;
  t0 = clock64(); for (int i = 0; i < N; ++i) dmma884(d0, d1, a * 1e-30, 1.0); t1 = clock64(); cyc[6] = t1 - t0; a += d0 + d1;
  // 7: SHFL chain (64-bit)
  t0 = clock64(); for (int i = 0; i < N; ++i) a = __shfl_sync(0xffffffffu, a, (i + 1) & 31); t1 = clock64(); cyc[7] = t1 - t0;
  // 8: DFMA throughput single warp, 8 independent
  double b[8]; for (int j = 0; j < 8; ++j) b[j] = a + j;
  t0 = clock64(); for (int i = 0; i < N; ++i) { for (int j = 0; j < 8; ++j) b[j] = fma(b[j], 1.0000001, 1e-9); } t1 = clock64(); cyc[8] = t1 - t0;
  for (int j = 0; j < 8; ++j) a += b[j];
  // 9: DMMA throughput single warp, 8 independent
  double c[8][2]; for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0;
  t0 = clock64(); for (int i = 0; i < N; ++i) { for (int j = 0; j < 8; ++j) dmma884(c[j][0], c[j][1], a * 1e-30, 1.0); } t1 = clock64(); cyc[9] = t1 - t0;
  for (int j = 0; j < 8; ++j) a += c[j][0] + c[j][1];
  // 10: 1/x chain
  t0 = clock64(); for (int i = 0; i < N; ++i) a = 1.0 / a + 1.5; t1 = clock64(); cyc[10] = t1 - t0;
  if (a == 1.2345) out[0] = a;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 64); cudaMalloc(&c, 16 * 8);
  k<<<1, 32>>>(d, c, 1.0); k<<<1, 32>>>(d, c, 1.0); cudaDeviceSynchronize();
  long long h[16]; cudaMemcpy(h, c, 16 * 8, cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA dep", "DMUL dep", "rsqrt()+add dep", "fast_rsqrt+add dep", "MUFU.RSQ64H dep", "LDS dep (+DADD+cvt)", "DMMA m8n8k4 dep", "SHFL64 dep", "DFMA x8 indep (per 8)", "DMMA x8 indep (per 8)", "1/x + add dep"};
  for (int i = 0; i < 11; ++i) printf("%-28s %8.2f cycles/iter\n", names[i], (double)h[i] / N);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
