// Global Kriging: fit once on all N samples, then predict every target (reference fitpredictfull,
// src/models.jl:200-234: fit krig.jl:58-75 once; per target weights krig.jl:316-339, krigmean
// krig.jl:245-258, predictvar krig.jl:264-293).
//
// fit (one-off, O(N^3)):   C = covariance matrix (column-major, padded to a multiple of 128 with an
//   identity block), blocked right-looking Cholesky C = L L' (64-wide panels: in-CTA diagonal factor
//   + inverse, then DMMA GEMMs), explicit Linv = L^-1 by blocked forward substitution,
//   D = C^-1 [z | F] = Linv' (Linv [z | F]), and the small Schur pieces G = F'C^-1F, hw = (C^-1F)'z.
// predict (hot, O(M N^2/2)): for a tile of 64 targets the kernel builds the covariance vectors
//   r (N x 64) block by block in shared memory and contracts them against the rows of Linv on the FP64
//   DMMA path (mma.sync m8n8k4), accumulating |Linv r|^2 = r'C^-1 r; the O(N) dot products alpha'r and
//   (C^-1F)'r ride along on the same r tiles.  Then with g = (C^-1F)'r - f0 and nu = G^-1 g:
//       mean = alpha'r - hw'nu          variance = C0 - r'C^-1r + g'nu + 1e-10
//   which is block elimination of the reference's saddle-point system [C F; F' 0][lambda; nu] = [r; f0].
#include <algorithm>
#include <cmath>
#include <cstring>

#include <cstdlib>
#include <string>

#include "gsm_internal.cuh"

namespace {

constexpr int NB = 64;    // Cholesky panel width
constexpr int PADN = 128; // N is padded to a multiple of this
constexpr size_t POTF2_SMEM = 2 * NB * (NB + 1) * sizeof(double);

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ---- covariance matrix ----------------------------------------------------------------------
__global__ void cov_matrix_kernel(const double* __restrict__ px, const double* __restrict__ py,
                                  const double* __restrict__ pz, int n, int np, VarioDev v, double* __restrict__ C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= np) return;
  double val;
  if (i >= n || j >= n) {
    val = (i == j) ? 1.0 : 0.0;
  } else if (i == j) {
    val = v.sill;
  } else {
    const double dx = px[i] - px[j], dy = py[i] - py[j], dz = pz[i] - pz[j];
    val = gsm_cov_vec(v, dx, dy, dz);
  }
  C[(size_t)j * np + i] = val;
}

// ---- diagonal block: Cholesky + inverse of a 64x64 block in one CTA --------------------------
// A (ld) holds the block in place (lower part becomes L); Linv receives inv(L) (64x64, ld 64).
__global__ void __launch_bounds__(256) potf2_inv_kernel(double* __restrict__ A, int ld, double* __restrict__ Linv,
                                                        int* __restrict__ info) {
  extern __shared__ double potf2_smem[];
  double (*S)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potf2_smem);
  double (*X)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potf2_smem + NB * (NB + 1));
  __shared__ double piv;
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += 256) {
    int r = e % NB, c = e / NB;
    S[r][c] = A[(size_t)c * ld + r];
  }
  __syncthreads();
  for (int j = 0; j < NB; ++j) {
    if (tid == 0) {
      double d = S[j][j];
      if (!(d > 0.0)) {
        atomicExch(info, 1);
        d = 1.0;
      }
      piv = sqrt(d);
      S[j][j] = piv;
    }
    __syncthreads();
    if (tid > j && tid < NB) S[tid][j] /= piv;
    __syncthreads();
    // trailing update of the lower triangle
    for (int e = tid; e < NB * NB; e += 256) {
      int r = e % NB, c = e / NB;
      if (c > j && r >= c) S[r][c] -= S[r][j] * S[c][j];
    }
    __syncthreads();
  }
  // inverse of L: thread c solves L x = e_c
  if (tid < NB) {
    const int c = tid;
    for (int i = 0; i < NB; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= S[i][k] * X[k][c];
      X[i][c] = (i < c) ? 0.0 : s / S[i][i];
    }
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) {
    int r = e % NB, c = e / NB;
    A[(size_t)c * ld + r] = (r >= c) ? S[r][c] : 0.0;
    Linv[(size_t)c * NB + r] = X[r][c];
  }
}

// ---- generic column-major DGEMM on DMMA:  C = beta*C + alpha * A * op(B) ----------------------
// CTA tile 64x64, 4 warps of 32x32, K step 16.  M, N multiples of 64, K multiple of 16.
// TRANSB: op(B) = B' (B is N x K);  else B is K x N.  lower_only: skip tiles strictly above the diagonal.
constexpr int GT = 64, GK = 16, GLDA = GT + 8, GLDB = GK + 4;

template <bool TRANSB>
__global__ void __launch_bounds__(128) dgemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A,
                                                    int lda, const double* __restrict__ B, int ldb, double beta,
                                                    double* __restrict__ C, int ldc, int lower_only) {
  const int bm = blockIdx.x * GT, bn = blockIdx.y * GT;
  if (lower_only && bn > bm) return;
  __shared__ __align__(16) double As[GK * GLDA];   // As[k][m]
  __shared__ __align__(16) double Bs[GT * GLDB];   // Bs[n][k]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int k0 = 0; k0 < K; k0 += GK) {
    for (int e = tid; e < GK * GT; e += 128) {   // A tile: 64 rows x 16 k, coalesced along rows
      int m = e % GT, k = e / GT;
      As[k * GLDA + m] = A[(size_t)(k0 + k) * lda + bm + m];
    }
    if (TRANSB) {
      for (int e = tid; e < GK * GT; e += 128) { // B is N x K: element (n, k) at B[k*ldb + n]
        int n = e % GT, k = e / GT;
        Bs[n * GLDB + k] = B[(size_t)(k0 + k) * ldb + bn + n];
      }
    } else {
      for (int e = tid; e < GK * GT; e += 128) { // B is K x N: element (k, n) at B[n*ldb + k]
        int k = e % GK, n = e / GK;
        Bs[n * GLDB + k] = B[(size_t)(bn + n) * ldb + k0 + k];
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[(kk + t) * GLDA + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[(wn + j * 8 + g) * GLDB + kk + t];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int r = bm + wm + i * 8 + g, c = bn + wn + j * 8 + 2 * t + s;
        double* p = &C[(size_t)c * ldc + r];
        *p = (beta == 0.0 ? 0.0 : beta * *p) + alpha * acc[i][j][s];
      }
}

// identity minus T in place:  T <- I - T  restricted to a 64-row block starting at global row r0
__global__ void eye_minus_kernel(double* __restrict__ T, int ld, int r0, int ncols) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (c >= ncols) return;
  double* p = &T[(size_t)c * ld + r];
  *p = ((r0 + r == c) ? 1.0 : 0.0) - *p;
}

// RHS block [z - mean | F] (np x 64, zero padded)
__global__ void rhs_kernel(const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pz,
                           const double* __restrict__ z, int n, int np, ModelDev m, int ncon, double* __restrict__ R) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  for (int c = 0; c < NB; ++c) {
    double v = 0.0;
    if (i < n) {
      if (c == 0) v = (m.kind == GSM_KRIG_SIMPLE) ? z[i] - m.mean : z[i];
      else if (c <= ncon) v = (m.kind == GSM_KRIG_ORDINARY) ? 1.0 : gsm_drift(m, c - 1, px[i], py[i], pz[i]);
    }
    R[(size_t)c * np + i] = v;
  }
}

// ---- prediction kernel ------------------------------------------------------------------------
constexpr int PT = 64;          // targets per CTA
constexpr int PM = 128;         // Linv rows per pass
constexpr int PK = 16;          // k step
constexpr int PLDA = PM + 8;    // As[k][m]
constexpr int PLDB = PK + 4;    // Bs[n][k]
constexpr int PTHREADS = 256;

struct GlobalDev {
  const double* Linv;   // np x np column-major, lower
  const double* D;      // np x 64 column-major: col 0 = alpha, cols 1..ncon = C^-1 F
  const double* px;
  const double* py;
  const double* pz;
  int n, np, ncon;
  double Ginv[GSM_MAX_DRIFT * GSM_MAX_DRIFT];
  double hw[GSM_MAX_DRIFT];
};

// ---- batched prediction (default): covariance tile once, then a triangular DMMA GEMM with a norm epilogue ----
// For a batch of Mb targets:
//   1. global_r_kernel     R[k][n] = cov(sample k, target n), k-major (np x Mb), written once; and the inner
//                          products dots[c][n] = sum_k D[c][k] R[k][n] (mean and drift terms).
//   2. global_gemm_norm    Y = Linv R is never stored: each CTA forms a 128 x 64 block of it on the DMMA path
//                          (3-stage cp.async pipeline, Linv lower triangular so K stops at the block's last row)
//                          and adds the column sums of squares to uu[n] = |Linv c0(n)|^2 = c0' C^-1 c0.
//   3. global_finish       mean, variance and the drift correction per target.
// (The first version rebuilt the covariance tile for every row block, 16x redundant at N = 4000, with no load /
// compute overlap: 969 ms against 582 ms at C2; removed in round 2.)
constexpr int QK = 16;            // k step
constexpr int QLDA = PM + 4;      // As[k][m]: stride = 4 mod 16 doubles -> the (t, g) fragment loads are conflict-free
constexpr int QLDB = PT + 4;      // Bs[k][n]
constexpr int QSTAGES = 3;
constexpr int QTHREADS = 256;

__global__ void __launch_bounds__(256)
global_r_kernel(GlobalDev gd, VarioDev v, TargetsDev tg, BlockDev bk, long long t0, int mb, double* __restrict__ R,
                double* __restrict__ dots) {
  // block: 64 targets x 128 samples; thread: target tid % 64, 32 samples of group tid / 64
  __shared__ double part[4][1 + GSM_MAX_DRIFT][64];
  const int tid = threadIdx.x, nl = tid & 63, grp = tid >> 6;
  const int n = blockIdx.x * 64 + nl;
  const int k0 = blockIdx.y * 128 + grp * 32;
  double tx = 0, ty = 0, tz = 0;
  const bool tlive = t0 + n < tg.count;
  if (tlive) gsm_target_coords(tg, tg.first + t0 + n, tx, ty, tz);
  double acc[1 + GSM_MAX_DRIFT];
  for (int c = 0; c <= gd.ncon; ++c) acc[c] = 0.0;
  for (int kk = 0; kk < 32; ++kk) {
    const int s = k0 + kk;
    double r = 0.0;
    if (s < gd.n && tlive) {
      const double dx = gd.px[s] - tx, dy = gd.py[s] - ty, dz = gd.pz[s] - tz;
      r = bk.nq > 0 ? gsm_cov_block(v, bk, dx, dy, dz) : gsm_cov_vec(v, dx, dy, dz);
      for (int c = 0; c <= gd.ncon; ++c) acc[c] = fma(gd.D[(size_t)c * gd.np + s], r, acc[c]);
    }
    R[(size_t)s * mb + n] = r;
  }
  for (int c = 0; c <= gd.ncon; ++c) part[grp][c][nl] = acc[c];
  __syncthreads();
  // partial sums of this sample block; global_finish_kernel adds the blocks in a fixed order (reproducible)
  if (grp == 0)
    for (int c = 0; c <= gd.ncon; ++c)
      dots[((size_t)blockIdx.y * (1 + GSM_MAX_DRIFT) + c) * mb + n] =
          (part[0][c][nl] + part[1][c][nl]) + (part[2][c][nl] + part[3][c][nl]);
}

__global__ void __launch_bounds__(QTHREADS, 2)
global_gemm_norm_kernel(const double* __restrict__ Linv, int np, const double* __restrict__ R, int mb, int ncb,
                        double* __restrict__ uu) {
  extern __shared__ __align__(16) double gsm_smem[];
  double* As = gsm_smem;                              // [QSTAGES][QK][QLDA]
  double* Bs = gsm_smem + QSTAGES * QK * QLDA;        // [QSTAGES][QK][QLDB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  const int nib = np / PM;
  // heavy row blocks (long K) first
  const int ib = nib - 1 - (int)(blockIdx.x / ncb), cb = (int)(blockIdx.x % ncb);
  const int ksteps = (ib + 1) * PM / QK;
  const double* Ag = Linv + (size_t)ib * PM;          // + k * np + m
  const double* Bg = R + (size_t)cb * PT;             // + k * mb + n

  auto load_stage = [&](int st, int ks) {
    const int k0 = ks * QK;
    // A: QK x 128 doubles = 1024 16-byte chunks, 4 per thread; B: QK x 64 doubles = 512 chunks, 2 per thread
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ch = tid + q * QTHREADS, k = ch >> 6, m2 = (ch & 63) * 2;
      const unsigned d = (unsigned)__cvta_generic_to_shared(&As[(st * QK + k) * QLDA + m2]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(Ag + (size_t)(k0 + k) * np + m2) : "memory");
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int ch = tid + q * QTHREADS, k = ch >> 5, n2 = (ch & 31) * 2;
      const unsigned d = (unsigned)__cvta_generic_to_shared(&Bs[(st * QK + k) * QLDB + n2]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(Bg + (size_t)(k0 + k) * mb + n2) : "memory");
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < QSTAGES - 1; ++s) {
    if (s < ksteps) load_stage(s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int ks = 0; ks < ksteps; ++ks) {
    asm volatile("cp.async.wait_group %0;" ::"n"(QSTAGES - 2) : "memory");
    __syncthreads();   // stage ks landed for everyone; stage ks-1 (about to be refilled) is no longer read
    if (ks + QSTAGES - 1 < ksteps) load_stage((ks + QSTAGES - 1) % QSTAGES, ks + QSTAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double* a = &As[(ks % QSTAGES) * QK * QLDA];
    const double* b = &Bs[(ks % QSTAGES) * QK * QLDB];
#pragma unroll
    for (int kk = 0; kk < QK; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = a[(kk + t) * QLDA + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = b[(kk + t) * QLDB + wn + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  // column sums of squares of this 128 x 64 block of Y: one partial per (row block, column), fixed order
  __syncthreads();                       // the tile buffers are free: reuse them for the 4 x 64 warp-row partials
  double* red = gsm_smem;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      double p = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) p = fma(acc[i][j][s], acc[i][j][s], p);
      p += __shfl_xor_sync(0xffffffffu, p, 4);
      p += __shfl_xor_sync(0xffffffffu, p, 8);
      p += __shfl_xor_sync(0xffffffffu, p, 16);
      if (g == 0) red[(warp & 3) * PT + wn + j * 8 + 2 * t + s] = p;
    }
  __syncthreads();
  if (tid < PT) uu[(size_t)ib * mb + cb * PT + tid] = (red[tid] + red[PT + tid]) + (red[2 * PT + tid] + red[3 * PT + tid]);
}

__global__ void global_finish_kernel(GlobalDev gd, double c0, ModelDev m, TargetsDev tg, long long t0, int mb,
                                     const double* __restrict__ dots_part, const double* __restrict__ uu_part,
                                     double* __restrict__ out_mean, double* __restrict__ out_var) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= mb || t0 + n >= tg.count) return;
  const int ncon = gd.ncon, nib = gd.np / PM;
  double dots[1 + GSM_MAX_DRIFT];
  double usum = 0.0;
  for (int c = 0; c <= ncon; ++c) dots[c] = 0.0;
  for (int ib = 0; ib < nib; ++ib) {
    for (int c = 0; c <= ncon; ++c) dots[c] += dots_part[((size_t)ib * (1 + GSM_MAX_DRIFT) + c) * mb + n];
    if (out_var) usum += uu_part[(size_t)ib * mb + n];
  }
  double mean = dots[0], var = c0 - usum;   // c0: covzero (krig.jl:295-301), the sill for a point target
  if (m.kind == GSM_KRIG_SIMPLE) {
    mean += m.mean;
  } else {
    double tx, ty, tz;
    gsm_target_coords(tg, tg.first + t0 + n, tx, ty, tz);
    double gq[GSM_MAX_DRIFT], nu[GSM_MAX_DRIFT];
    for (int c = 0; c < ncon; ++c) {
      const double f0 = (m.kind == GSM_KRIG_ORDINARY) ? 1.0 : gsm_drift(m, c, tx, ty, tz);
      gq[c] = dots[1 + c] - f0;
    }
    for (int c = 0; c < ncon; ++c) {
      double s = 0.0;
      for (int d = 0; d < ncon; ++d) s = fma(gd.Ginv[c * ncon + d], gq[d], s);
      nu[c] = s;
    }
    for (int c = 0; c < ncon; ++c) {
      mean = fma(-gd.hw[c], nu[c], mean);
      var = fma(gq[c], nu[c], var);
    }
  }
  out_mean[t0 + n] = mean;
  if (out_var) out_var[t0 + n] = var + 1e-10;
}


__global__ void split_coords_kernel(const double* __restrict__ coords, int dim, int n, int np, double* __restrict__ px,
                                    double* __restrict__ py, double* __restrict__ pz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  px[i] = i < n ? coords[(size_t)i * dim] : 0.0;
  py[i] = (i < n && dim > 1) ? coords[(size_t)i * dim + 1] : 0.0;
  pz[i] = (i < n && dim > 2) ? coords[(size_t)i * dim + 2] : 0.0;
}

// ---- transpose helper (np multiple of 32) -----------------------------------------------------
__global__ void transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int n) {
  __shared__ double tile[32][33];
  int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
  for (int j = 0; j < 32; j += 8) tile[threadIdx.y + j][threadIdx.x] = in[(size_t)(y + j) * n + x];
  __syncthreads();
  x = blockIdx.y * 32 + threadIdx.x;
  y = blockIdx.x * 32 + threadIdx.y;
  for (int j = 0; j < 32; j += 8) out[(size_t)(y + j) * n + x] = tile[threadIdx.x][threadIdx.y + j];
}
void transpose_launch(const double* in, double* out, int n, cudaStream_t st) {
  transpose_kernel<<<dim3(n / 32, n / 32), dim3(32, 8), 0, st>>>(in, out, n);
}

int gemm(gsm_ctx* ctx, bool transb, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
         int ldb, double beta, double* C, int ldc, int lower_only) {
  if (M <= 0 || N <= 0) return GSM_OK;
  dim3 grid(M / GT, N / GT);
  if (K <= 0) K = 0;
  if (transb) dgemm_kernel<true><<<grid, 128, 0, ctx->stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
  else dgemm_kernel<false><<<grid, 128, 0, ctx->stream>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

}  // namespace

struct FitBufs {
  DevBuf C, Linv, D, T, diag, px, py, pz, info;
};

// gsm_fit carries its own device buffers (declared in gsm_internal.cuh as DevBuf members Linv/W/schur);
// the extra scratch lives here for the duration of the fit.
extern "C" {

int gsm_global_krige_fit(gsm_ctx* ctx, const gsm_variogram* vario, const gsm_model* model, gsm_fit** out) {
  if (!ctx || !out) return GSM_ERR_INVALID;
  *out = nullptr;
  if (ctx->n <= 0) {
    gsm_set_error(ctx, "no samples: call gsm_set_samples first");
    return GSM_ERR_NO_SAMPLES;
  }
  GSM_CHECK(gsm_check_vario(ctx, vario));
  GsmRange nv("gsm_global_krige_fit");
  if (!model || model->kind < 0 || model->kind > 2 ||
      (model->kind == GSM_KRIG_UNIVERSAL && (model->ndrift < 1 || model->ndrift > GSM_MAX_DRIFT))) {
    gsm_set_error(ctx, "invalid Kriging model descriptor");
    return GSM_ERR_INVALID;
  }
  if (model->kind == GSM_KRIG_UNIVERSAL)   // same validation as the neighbourhood entry points (universal.jl:56-57)
    for (int a = 0; a < model->ndrift; ++a)
      for (int c = 0; c < 3; ++c)
        if (model->exps[a][c] < 0 || (c >= ctx->dim && model->exps[a][c] != 0)) {
          gsm_set_error(ctx, "UniversalKriging: invalid monomial exponent");
          return GSM_ERR_INVALID;
        }
  if (ctx->n_missing > 0) {
    gsm_set_error(ctx, "global Kriging with missing values: drop the missing samples on the host (the reference's zeroed "
                       "rows / columns are the reduced system; see the host mirror's fitpredict)");
    return GSM_ERR_UNSUPPORTED;
  }
  const long long n = ctx->n;
  if (n > 60000) {
    gsm_set_error(ctx, "global Kriging with more than 60000 samples does not fit the dense factor in HBM");
    return GSM_ERR_UNSUPPORTED;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->device);
  ctx->tim = gsm_timings{};
  cudaStream_t st = ctx->stream;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, st);

  const int np = (int)((n + PADN - 1) / PADN * PADN);
  const int ncon = model->kind == GSM_KRIG_SIMPLE ? 0 : (model->kind == GSM_KRIG_ORDINARY ? 1 : model->ndrift);
  gsm_fit* f = new gsm_fit();
  f->ctx = ctx;
  f->vario = *vario;
  f->model = *model;
  f->n = n;
  f->ncon = ncon;
  FitBufs b;
  int rc = GSM_OK;
  auto fail = [&](int code) {
    DevBuf* all[] = {&b.C, &b.T, &b.diag, &b.info, &f->Linv, &f->W, &f->schur};
    for (DevBuf* d : all)
      if (d->p) cudaFree(d->p);
    delete f;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaSetDevice(prev);
    return code;
  };
#define FIT_TRY(expr)                  \
  do {                                 \
    rc = (expr);                       \
    if (rc != GSM_OK) return fail(rc); \
  } while (0)
// a CUDA call inside the fit: on failure record the message and leave through fail() (frees the fit's buffers and
// events, restores the device) instead of GSM_CUDA's bare return
#define FIT_CUDA(call)                                                                                  \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      gsm_set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(e_) + " (global Kriging fit)"); \
      return fail(GSM_ERR_CUDA);                                                                        \
    }                                                                                                   \
  } while (0)
  FIT_TRY(gsm_reserve(ctx, b.C, (size_t)np * np * sizeof(double)));
  FIT_TRY(gsm_reserve(ctx, f->Linv, (size_t)np * np * sizeof(double)));
  FIT_TRY(gsm_reserve(ctx, f->W, (size_t)np * NB * sizeof(double)));        // D = C^-1 [z | F]
  FIT_TRY(gsm_reserve(ctx, b.T, (size_t)np * NB * sizeof(double)));
  FIT_TRY(gsm_reserve(ctx, b.diag, (size_t)(np / NB) * NB * NB * sizeof(double)));
  FIT_TRY(gsm_reserve(ctx, b.info, sizeof(int)));
  FIT_TRY(gsm_reserve(ctx, f->schur, (size_t)3 * np * sizeof(double)));     // px, py, pz (SoA, padded)
  double* C = (double*)b.C.p;
  double* Linv = (double*)f->Linv.p;
  double* D = (double*)f->W.p;
  double* T = (double*)b.T.p;
  double* diag = (double*)b.diag.p;
  double* px = (double*)f->schur.p;
  double* py = px + np;
  double* pz = py + np;
  cudaMemsetAsync(b.info.p, 0, sizeof(int), st);

  const VarioDev vd = gsm_make_vario(*vario);
  ModelDev md{};
  md.kind = model->kind;
  md.mean = model->mean;
  md.ndrift = model->kind == GSM_KRIG_UNIVERSAL ? model->ndrift : 0;
  for (int a = 0; a < md.ndrift; ++a)
    for (int d = 0; d < 3; ++d) md.exps[a][d] = model->exps[a][d];

  split_coords_kernel<<<(np + 255) / 256, 256, 0, st>>>((const double*)ctx->coords.p, ctx->dim, (int)n, np, px, py, pz);
  cov_matrix_kernel<<<dim3((np + 255) / 256, np), 256, 0, st>>>(px, py, pz, (int)n, np, vd, C);
  ctx->tim.launches += 2;

  // ---- blocked Cholesky, lower, in place in C ----
  cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTF2_SMEM);
  const int nblk = np / NB;
  for (int kb = 0; kb < nblk; ++kb) {
    double* Akk = C + (size_t)kb * NB * np + kb * NB;
    double* Dk = diag + (size_t)kb * NB * NB;
    potf2_inv_kernel<<<1, 256, POTF2_SMEM, st>>>(Akk, np, Dk, (int*)b.info.p);
    ctx->tim.launches++;
    const int rows = np - (kb + 1) * NB;
    if (rows > 0) {
      // panel: L_ik = A_ik * inv(L_kk)'   (T is scratch because the GEMM is not in-place safe)
      double* Aik = C + (size_t)kb * NB * np + (kb + 1) * NB;
      FIT_TRY(gemm(ctx, true, rows, NB, NB, 1.0, Aik, np, Dk, NB, 0.0, T, np, 0));
      FIT_CUDA(cudaMemcpy2DAsync(Aik, (size_t)np * sizeof(double), T, (size_t)np * sizeof(double),
                                      (size_t)rows * sizeof(double), NB, cudaMemcpyDeviceToDevice, st));
      // trailing update: A_ij -= L_ik L_jk'   (lower tiles only)
      double* Atr = C + (size_t)(kb + 1) * NB * np + (kb + 1) * NB;
      FIT_TRY(gemm(ctx, true, rows, rows, NB, -1.0, Aik, np, Aik, np, 1.0, Atr, np, 1));
    }
  }
  // ---- Linv = L^-1 : blocked forward substitution on the identity ----
  FIT_CUDA(cudaMemsetAsync(Linv, 0, (size_t)np * np * sizeof(double), st));
  for (int ib = 0; ib < nblk; ++ib) {
    const int ncols = (ib + 1) * NB;   // inv(L) is lower triangular: block row ib has ib+1 nonzero block columns
    double* Ti = T;                    // 64 x ncols scratch (ld = NB) -- reuse T as 64-row matrix
    // T = L[ib, 0:ib] * Linv[0:ib, 0:ncols]
    if (ib > 0) {
      FIT_TRY(gemm(ctx, false, NB, ncols, ib * NB, 1.0, C + (size_t)ib * NB, np, Linv, np, 0.0, Ti, NB, 0));
    } else {
      FIT_CUDA(cudaMemsetAsync(Ti, 0, (size_t)NB * ncols * sizeof(double), st));
    }
    eye_minus_kernel<<<dim3((ncols + 127) / 128, NB), 128, 0, st>>>(Ti, NB, ib * NB, ncols);
    ctx->tim.launches++;
    // Linv[ib, 0:ncols] = inv(L_ii) * (I - T)
    FIT_TRY(gemm(ctx, false, NB, ncols, NB, 1.0, diag + (size_t)ib * NB * NB, NB, Ti, NB, 0.0, Linv + (size_t)ib * NB, np, 0));
  }
  // ---- D = C^-1 [z | F] = Linv' (Linv R) ----
  rhs_kernel<<<(np + 255) / 256, 256, 0, st>>>(px, py, pz, (const double*)ctx->values.p, (int)n, np, md, ncon, D);
  ctx->tim.launches++;
  FIT_TRY(gemm(ctx, false, np, NB, np, 1.0, Linv, np, D, np, 0.0, T, np, 0));     // T = Linv R
  // D = Linv' T: transpose Linv into C (free after the factorisation) and reuse the NN GEMM
#undef FIT_TRY
#undef FIT_CUDA
  transpose_launch(Linv, C, np, st);
  ctx->tim.launches++;
  rc = gemm(ctx, false, np, NB, np, 1.0, C, np, T, np, 0.0, D, np, 0);
  if (rc != GSM_OK) return fail(rc);

  // ---- Schur pieces on the host: G = F' (C^-1 F), hw = (C^-1 F)' z ----
  std::vector<double> hD((size_t)np * (1 + ncon)), hR((size_t)np * (1 + ncon));
  rhs_kernel<<<(np + 255) / 256, 256, 0, st>>>(px, py, pz, (const double*)ctx->values.p, (int)n, np, md, ncon, T);
  ctx->tim.launches++;
  cudaMemcpyAsync(hD.data(), D, hD.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(hR.data(), T, hR.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
  int info = 0;
  cudaMemcpyAsync(&info, b.info.p, sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaEventRecord(e1, st);
  cudaError_t ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) {
    gsm_set_error(ctx, std::string("global fit failed: ") + cudaGetErrorString(ce));
    return fail(GSM_ERR_CUDA);
  }
  f->status = info == 0 ? 1 : 0;
  f->S.assign((size_t)GSM_MAX_DRIFT * GSM_MAX_DRIFT, 0.0);
  f->h.assign(GSM_MAX_DRIFT, 0.0);
  if (ncon > 0) {
    // G = F' C^-1 F (symmetric); invert by Gauss-Jordan with partial pivoting in long double
    std::vector<long double> G((size_t)ncon * ncon), Gi((size_t)ncon * ncon, 0.0L);
    for (int a = 0; a < ncon; ++a)
      for (int c = 0; c < ncon; ++c) {
        long double s = 0;
        for (long long i = 0; i < n; ++i) s += (long double)hR[(size_t)(1 + a) * np + i] * hD[(size_t)(1 + c) * np + i];
        G[a * ncon + c] = s;
      }
    for (int a = 0; a < ncon; ++a) Gi[a * ncon + a] = 1.0L;
    for (int col = 0; col < ncon; ++col) {
      int piv = col;
      for (int r = col + 1; r < ncon; ++r)
        if (fabsl(G[r * ncon + col]) > fabsl(G[piv * ncon + col])) piv = r;
      if (G[piv * ncon + col] == 0.0L) {
        f->status = 0;
        break;
      }
      for (int c = 0; c < ncon; ++c) {
        std::swap(G[col * ncon + c], G[piv * ncon + c]);
        std::swap(Gi[col * ncon + c], Gi[piv * ncon + c]);
      }
      long double d = 1.0L / G[col * ncon + col];
      for (int c = 0; c < ncon; ++c) {
        G[col * ncon + c] *= d;
        Gi[col * ncon + c] *= d;
      }
      for (int r = 0; r < ncon; ++r) {
        if (r == col) continue;
        long double fct = G[r * ncon + col];
        for (int c = 0; c < ncon; ++c) {
          G[r * ncon + c] -= fct * G[col * ncon + c];
          Gi[r * ncon + c] -= fct * Gi[col * ncon + c];
        }
      }
    }
    for (int a = 0; a < ncon; ++a) {
      for (int c = 0; c < ncon; ++c) f->S[a * ncon + c] = (double)Gi[a * ncon + c];
      long double s = 0;
      const double* zc = &hR[0];
      for (long long i = 0; i < n; ++i) s += (long double)hD[(size_t)(1 + a) * np + i] * zc[i];
      f->h[a] = (double)s;
    }
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ctx->tim.solve_ms = ms;
  ctx->tim.total_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(b.C.p);
  cudaFree(b.T.p);
  cudaFree(b.diag.p);
  cudaFree(b.info.p);
  cudaSetDevice(prev);
  *out = f;
  if (!f->status) {
    gsm_set_error(ctx, "global Kriging factorisation failed (non-positive pivot)");
  }
  return GSM_OK;
}

int gsm_global_krige_status(const gsm_fit* fit) { return fit ? fit->status : 0; }

int gsm_global_krige_predict(gsm_fit* f, const gsm_targets* t, double* out_mean, double* out_var) {
  if (!f || !t || !out_mean) return GSM_ERR_INVALID;
  gsm_ctx* ctx = f->ctx;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->device);
  struct Restore {
    int d;
    ~Restore() { cudaSetDevice(d); }
  } restore{prev};
  ctx->tim = gsm_timings{};
  cudaStream_t st = ctx->stream;
  if (t->kind != GSM_TARGETS_GRID && t->kind != GSM_TARGETS_POINTS) {
    gsm_set_error(ctx, "invalid target domain descriptor");
    return GSM_ERR_INVALID;
  }
  if (t->dim != ctx->dim) {
    gsm_set_error(ctx, "target domain and samples have different dimensions");
    return GSM_ERR_INVALID;
  }
  long long total = 1;
  if (t->kind == GSM_TARGETS_GRID) for (int d = 0; d < t->dim; ++d) total *= t->dims[d];
  else total = t->npoints;
  const long long first = t->first, count = t->count < 0 ? total - t->first : t->count;
  if (first < 0 || count < 0 || first + count > total) {
    gsm_set_error(ctx, "target shard [first, first+count) is outside the domain");
    return GSM_ERR_INVALID;
  }
  cudaEvent_t e0, e1, e2, e3;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventCreate(&e2);
  cudaEventCreate(&e3);
  cudaEventRecord(e0, st);
  TargetsDev td{};
  td.kind = t->kind;
  td.dim = t->dim;
  for (int c = 0; c < 3; ++c) {
    td.origin[c] = c < t->dim ? t->origin[c] : 0.0;
    td.spacing[c] = c < t->dim ? t->spacing[c] : 0.0;
    td.dims[c] = c < t->dim ? t->dims[c] : 1;
  }
  td.first = first;
  td.count = count;
  if (t->kind == GSM_TARGETS_POINTS) {
    const double* src = t->points + first * t->dim;
    if (gsm_is_device_ptr(t->points)) {
      td.points = src;
    } else {
      GSM_CHECK(gsm_reserve(ctx, ctx->tgt_points, (size_t)std::max<long long>(count, 1) * t->dim * sizeof(double)));
      GSM_CUDA(ctx, cudaMemcpyAsync(ctx->tgt_points.p, src, (size_t)count * t->dim * sizeof(double),
                                    cudaMemcpyHostToDevice, st));
      td.points = (const double*)ctx->tgt_points.p;
    }
    td.first = 0;
  }
  cudaEventRecord(e1, st);
  const bool mean_dev = gsm_is_device_ptr(out_mean), var_dev = out_var && gsm_is_device_ptr(out_var);
  double* dmean = out_mean;
  double* dvar = out_var;
  if (!mean_dev) {
    GSM_CHECK(gsm_reserve(ctx, ctx->out_mean, (size_t)std::max<long long>(count, 1) * sizeof(double)));
    dmean = (double*)ctx->out_mean.p;
  }
  if (out_var && !var_dev) {
    GSM_CHECK(gsm_reserve(ctx, ctx->out_var, (size_t)std::max<long long>(count, 1) * sizeof(double)));
    dvar = (double*)ctx->out_var.p;
  }
  if (count > 0) {
    GlobalDev gd{};
    const int np = (int)((f->n + PADN - 1) / PADN * PADN);
    gd.Linv = (const double*)f->Linv.p;
    gd.D = (const double*)f->W.p;
    gd.px = (const double*)f->schur.p;
    gd.py = gd.px + np;
    gd.pz = gd.py + np;
    gd.n = (int)f->n;
    gd.np = np;
    gd.ncon = f->ncon;
    for (int a = 0; a < f->ncon * f->ncon; ++a) gd.Ginv[a] = f->S[a];
    for (int a = 0; a < f->ncon; ++a) gd.hw[a] = f->h[a];
    const VarioDev vd = gsm_make_vario(f->vario);
    GSM_CHECK(gsm_block_prepare(ctx, vd));
    ModelDev md{};
    md.kind = f->model.kind;
    md.mean = f->model.mean;
    md.ndrift = f->model.kind == GSM_KRIG_UNIVERSAL ? f->model.ndrift : 0;
    for (int a = 0; a < md.ndrift; ++a)
      for (int d = 0; d < 3; ++d) md.exps[a][d] = f->model.exps[a][d];
    {
      // batches of mb targets: the covariance tile R (np x mb, k-major) stays around 256 MB
      long long mb = (256LL << 20) / ((long long)np * 8) / PT * PT;
      mb = std::max<long long>(PT, std::min<long long>(mb, 8192));
      mb = std::min<long long>(mb, (count + PT - 1) / PT * PT);
      const int ncb = (int)(mb / PT), nib = np / PM;
      GSM_CHECK(gsm_reserve(ctx, ctx->glob_R, (size_t)np * mb * sizeof(double)));
      GSM_CHECK(gsm_reserve(ctx, ctx->glob_acc, (size_t)nib * (2 + GSM_MAX_DRIFT) * mb * sizeof(double)));
      double* Rb = (double*)ctx->glob_R.p;
      double* dots = (double*)ctx->glob_acc.p;           // per sample block: (1 + GSM_MAX_DRIFT) x mb partial sums
      double* uu = dots + (size_t)nib * (1 + GSM_MAX_DRIFT) * mb;   // per row block: mb partial sums of squares
      const size_t smem = (size_t)QSTAGES * QK * (QLDA + QLDB) * sizeof(double);
      GSM_CUDA(ctx, cudaFuncSetAttribute(global_gemm_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (long long t0 = 0; t0 < count; t0 += mb) {
        global_r_kernel<<<dim3((unsigned)ncb, (unsigned)nib), 256, 0, st>>>(gd, vd, td, ctx->block_cur, t0, (int)mb, Rb, dots);
        if (dvar)
          global_gemm_norm_kernel<<<(unsigned)(ncb * nib), QTHREADS, smem, st>>>(gd.Linv, np, Rb, (int)mb, ncb, uu);
        global_finish_kernel<<<(unsigned)((mb + 127) / 128), 128, 0, st>>>(gd, ctx->block_cur.c0, md, td, t0, (int)mb, dots, uu, dmean, dvar);
        ctx->tim.launches += dvar ? 3 : 2;
      }
      GSM_CUDA(ctx, cudaGetLastError());
    }
  }
  cudaEventRecord(e2, st);
  if (!mean_dev && count > 0)
    GSM_CUDA(ctx, cudaMemcpyAsync(out_mean, dmean, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (out_var && !var_dev && count > 0)
    GSM_CUDA(ctx, cudaMemcpyAsync(out_var, dvar, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, st));
  cudaEventRecord(e3, st);
  GSM_CUDA(ctx, cudaStreamSynchronize(st));
  float a = 0, b2 = 0, c = 0;
  cudaEventElapsedTime(&a, e0, e1);
  cudaEventElapsedTime(&b2, e1, e2);
  cudaEventElapsedTime(&c, e2, e3);
  ctx->tim.h2d_ms = a;
  ctx->tim.solve_ms = b2;
  ctx->tim.d2h_ms = c;
  ctx->tim.total_ms = a + b2 + c;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaEventDestroy(e2);
  cudaEventDestroy(e3);
  return GSM_OK;
}

int gsm_global_krige_free(gsm_fit* fit) {
  if (!fit) return GSM_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(fit->ctx->device);
  cudaStreamSynchronize(fit->ctx->stream);
  if (fit->Linv.p) cudaFree(fit->Linv.p);
  if (fit->W.p) cudaFree(fit->W.p);
  if (fit->schur.p) cudaFree(fit->schur.p);
  cudaSetDevice(prev);
  delete fit;
  return GSM_OK;
}

}  // extern "C"

