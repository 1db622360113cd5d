// B200 micro-benchmarks that size the local-Kriging kernel design: throughput of the FP64 vector pipe,
// FP64 DMMA, warp shuffles, shared-memory broadcast loads and FP64 special functions, all in
// warp-instructions per ns over the whole chip (and relative to DFMA).  Standalone: nvcc -> binary.
#include <cuda_runtime.h>
#include <cstdio>
#include <functional>

#define ITERS 2048
#define ILP 8

__global__ void k_dfma(double* out, double s) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = s + i + threadIdx.x;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], 1.0000001, 1e-9);
  double r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_dfma_lat(double* out, double s) {
  double a = s + threadIdx.x;
  for (int it = 0; it < ITERS * ILP; ++it) a = fma(a, 1.0000001, 1e-9);
  if (a == 1.2345) out[0] = a;
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__global__ void k_dmma884(double* out, double s) {
  double acc[ILP][2];
  for (int i = 0; i < ILP; ++i) acc[i][0] = acc[i][1] = 0;
  double a = s * 1e-3 + threadIdx.x * 1e-6, b = 1.0 + s * 1e-6;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  double r = 0; for (int i = 0; i < ILP; ++i) r += acc[i][0] + acc[i][1];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_dmma884_lat(double* out, double s) {
  double d0 = 0, d1 = 0;
  double a = s * 1e-3 + threadIdx.x * 1e-6, b = 1.0 + s * 1e-6;
  for (int it = 0; it < ITERS * ILP; ++it) dmma884(d0, d1, a, b);
  if (d0 + d1 == 1.2345) out[0] = d0;
}
__global__ void k_dmma1688(double* out, double s) {
  double acc[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0;
  double a[4] = {s * 1e-3, s * 2e-3, threadIdx.x * 1e-6, 1e-4}, b[2] = {1.0 + s * 1e-6, 0.5};
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma1688(acc[i], a, b);
  double r = 0; for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) r += acc[i][j];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_dmma16816(double* out, double s) {
  double acc[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0;
  double a[8], b[4];
  for (int j = 0; j < 8; ++j) a[j] = s * 1e-3 * (j + 1) + threadIdx.x * 1e-6;
  for (int j = 0; j < 4; ++j) b[j] = 1.0 + s * 1e-6 * j;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma16816(acc[i], a, b);
  double r = 0; for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) r += acc[i][j];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_shfl(double* out, double s) {
  int a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = __shfl_sync(0xffffffffu, a[i], (it + i) & 31);
  int r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 12345678) out[0] = r;
}
__global__ void k_shfl_lat(double* out, double s) {
  int a = threadIdx.x;
  for (int it = 0; it < ITERS * ILP; ++it) a = __shfl_sync(0xffffffffu, a, (a + 1) & 31);
  if (a == 12345678) out[0] = a;
}
__global__ void k_lds64_bcast(double* out, double s) {
  __shared__ double sm[256];
  sm[threadIdx.x] = s + threadIdx.x;
  __syncthreads();
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = 0;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] += *(volatile double*)&sm[(it + i * 7) & 255];
  double r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_lds128_bcast(double* out, double s) {
  __shared__ __align__(16) double sm[256];
  sm[threadIdx.x] = s + threadIdx.x;
  __syncthreads();
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = 0;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      double2 v;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(&sm[((it + i * 7) & 127) * 2])));
      a[i] += v.x + v.y;
    }
  double r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_lds64_rows(double* out, double s) {  // each lane its own address, conflict-free
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = s + i;
  __syncthreads();
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = 0;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] += *(volatile double*)&sm[(((it + i) & 31) * 32 + lane) & 1023];
  double r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_rsqrt(double* out, double s) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = s + i + threadIdx.x + 1.0;
  for (int it = 0; it < ITERS / 8; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = rsqrt(a[i]) + 1.5;
  double r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_sqrt(double* out, double s) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = s + i + threadIdx.x + 1.0;
  for (int it = 0; it < ITERS / 8; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = sqrt(a[i]) + 1.5;
  double r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_div(double* out, double s) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = s + i + threadIdx.x + 1.0;
  for (int it = 0; it < ITERS / 8; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 / a[i] + 1.5;
  double r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_exp(double* out, double s) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = -(s + i + threadIdx.x * 1e-3);
  for (int it = 0; it < ITERS / 8; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = exp(a[i]) - 1.5;
  double r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 1.2345) out[0] = r;
}
__global__ void k_mufu_rsq64h(double* out, double s) {
  int a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = 0x3ff00000 + threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      double x = __hiloint2double(a[i], 0);
      double y;
      asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
      a[i] = __double2hiint(y) + 0x00100000;
    }
  int r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 12345678) out[0] = r;
}
__global__ void k_dsetp_sel(double* out, double s) {
  double a[ILP]; double b = s;
  for (int i = 0; i < ILP; ++i) a[i] = s + i + threadIdx.x;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = a[i] < b + it ? a[i] + 1.0 : a[i];
  double r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 1.2345) out[0] = r;
}

struct Bench { const char* name; void (*fn)(double*, double); double ops_per_thread_iter; const char* unit; };

int main() {
  double* d; cudaMalloc(&d, 64);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device: %s, SMs=%d, clock(max)=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  const int blocks = p.multiProcessorCount * 8, threads = 256;
  struct Row { const char* name; void (*fn)(double*, double); double warp_instr; double flops_per_warp_instr; };
  Row rows[] = {
    {"dfma (ILP8)", k_dfma, (double)ITERS * ILP, 64},
    {"dfma latency chain", k_dfma_lat, (double)ITERS * ILP, 64},
    {"dmma m8n8k4 (ILP8)", k_dmma884, (double)ITERS * ILP, 512},
    {"dmma m8n8k4 latency chain", k_dmma884_lat, (double)ITERS * ILP, 512},
    {"dmma m16n8k8 (ILP8)", k_dmma1688, (double)ITERS * ILP, 2048},
    {"dmma m16n8k16 (ILP8)", k_dmma16816, (double)ITERS * ILP, 4096},
    {"shfl.b32 (ILP8)", k_shfl, (double)ITERS * ILP, 0},
    {"shfl.b32 latency chain", k_shfl_lat, (double)ITERS * ILP, 0},
    {"lds.64 broadcast", k_lds64_bcast, (double)ITERS * ILP, 0},
    {"lds.128 broadcast", k_lds128_bcast, (double)ITERS * ILP, 0},
    {"lds.64 lane-private (256B/warp)", k_lds64_rows, (double)ITERS * ILP, 0},
    {"rsqrt(double)", k_rsqrt, (double)(ITERS / 8) * ILP, 0},
    {"sqrt(double)", k_sqrt, (double)(ITERS / 8) * ILP, 0},
    {"1/x (double)", k_div, (double)(ITERS / 8) * ILP, 0},
    {"exp(double)", k_exp, (double)(ITERS / 8) * ILP, 0},
    {"mufu.rsq64h", k_mufu_rsq64h, (double)ITERS * ILP, 0},
    {"dsetp+sel (double compare-select)", k_dsetp_sel, (double)ITERS * ILP, 0},
  };
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double dfma_rate = 0;
  for (auto& r : rows) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      r.fn<<<blocks, threads>>>(d, 1.0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    double warps = (double)blocks * threads / 32;
    double winstr_per_ns = warps * r.warp_instr / (best * 1e6);
    double per_sm_clk = winstr_per_ns / p.multiProcessorCount / (p.clockRate * 1e-6);
    if (dfma_rate == 0) dfma_rate = winstr_per_ns;
    printf("%-36s %8.3f ms  %9.2f warp-instr/ns  %6.3f /SM/clk(max)  x%.2f of DFMA issue", r.name, best, winstr_per_ns, per_sm_clk, winstr_per_ns / dfma_rate);
    if (r.flops_per_warp_instr > 0) printf("  %7.2f TFLOP/s", winstr_per_ns * r.flops_per_warp_instr * 1e-3);
    if (err != cudaSuccess) printf("  ERR %s", cudaGetErrorString(err));
    printf("\n");
  }
  return 0;
}
