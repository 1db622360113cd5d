// Local Kriging for 32 < k <= 64 neighbours (config C5 sweeps k up to 64): one CTA per system, the
// augmented matrix [C ; B'] in shared memory.  Same mathematics as the warp-level kernels (block
// elimination of the reference's saddle system, krig.jl:58-75 / 316-339 / 245-293): right-looking
// Cholesky of the covariance block with the right-hand-side rows (c0, z, drift monomials) carried as
// extra rows, so the trailing NR x NR block ends up as -B'C^-1B, from which mean and variance follow.
// This is the capacity path, not the tuned one: k <= 32 runs on krige_local_pipe.cuh.
#include <algorithm>

#include "gsm_internal.cuh"

namespace {

constexpr int BK_THREADS = 128;
constexpr int BK_KMAX = GSM_MAX_NEIGHBORS;        // 64
constexpr int BK_NRMAX = 2 + GSM_MAX_DRIFT;       // 12
constexpr int BK_ROWS = BK_KMAX + BK_NRMAX;       // 76
constexpr int BK_LD = BK_ROWS + 1;                // odd leading dimension: column walks spread over banks

__global__ void __launch_bounds__(BK_THREADS)
local_krige_big_kernel(BinsDev b, VarioDev v, ModelDev m, int ncon, int centered, TargetsDev tg, int k, int minn,
                       const int* __restrict__ pos, const int* __restrict__ cnt, double* __restrict__ out_mean,
                       double* __restrict__ out_var, unsigned char* __restrict__ out_status) {
  extern __shared__ __align__(16) double sm[];
  double* A = sm;                          // A[c * BK_LD + r], lower part used, r in [0, k + nr)
  double* px = A + BK_ROWS * BK_LD;        // neighbour coordinates / values
  double* py = px + BK_KMAX;
  double* pz = py + BK_KMAX;
  double* pv = pz + BK_KMAX;
  __shared__ int s_pos[BK_KMAX];
  __shared__ int s_bad;
  __shared__ double s_scale[4];

  const long long t = blockIdx.x;
  const int tid = threadIdx.x;
  const int nr = 2 + ncon;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const int n = cnt[t];
  if (n < minn || n <= 0) {
    if (tid == 0) {
      out_mean[t] = qnan;
      if (out_var) out_var[t] = qnan;
      if (out_status) out_status[t] = GSM_ST_MISSING;
    }
    return;
  }
  double tx, ty, tz;
  gsm_target_coords(tg, tg.first + t, tx, ty, tz);

  // ---- canonical neighbour order: ascending record position (rank by counting, n <= 64) ----------
  if (tid < k) s_pos[tid] = tid < n ? pos[t * (long long)k + tid] : 0x7fffffff;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  if (tid < n) {
    const int mine = s_pos[tid];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (s_pos[j] < mine) ? 1 : 0;   // positions are distinct
    const double4 r = b.rec[mine];
    px[rank] = r.x;
    py[rank] = r.y;
    pz[rank] = r.z;
    pv[rank] = r.w;
  }
  __syncthreads();
  // drift scaling: centre at the target, scale by the neighbourhood radius (lower-set exponent tables)
  if (tid == 0) {
    double dmax = 0.0;
    for (int i = 0; i < n; ++i) {
      const double dx = px[i] - tx, dy = py[i] - ty, dz = pz[i] - tz;
      dmax = fmax(dmax, dx * dx + dy * dy + dz * dz);
    }
    s_scale[0] = centered ? tx : 0.0;
    s_scale[1] = centered ? ty : 0.0;
    s_scale[2] = centered ? tz : 0.0;
    s_scale[3] = (centered && dmax > 0.0) ? rsqrt(dmax) : 1.0;
  }
  __syncthreads();
  const double ox = s_scale[0], oy = s_scale[1], oz = s_scale[2], isc = s_scale[3];

  // ---- assemble: covariance block (lower), then the right-hand-side rows -------------------------
  // thread layout for every 2-D sweep below: 16 row lanes x 8 column lanes (no integer divisions in the loops)
  const int rx = tid & 15, cy = tid >> 4;
  for (int c = cy; c < n; c += 8)
    for (int r = c + rx; r < n; r += 16) {
      double val = v.sill;
      if (r != c) {
        const double dx = px[r] - px[c], dy = py[r] - py[c], dz = pz[r] - pz[c];
        val = gsm_cov_d2(v, dx * dx + dy * dy + dz * dz);
      }
      A[c * BK_LD + r] = val;
    }
  for (int q = cy; q < nr; q += 8)
    for (int c = rx; c < n; c += 16) {
      double val;
      if (q == 0) {
        const double dx = px[c] - tx, dy = py[c] - ty, dz = pz[c] - tz;
        val = gsm_cov_d2(v, dx * dx + dy * dy + dz * dz);
      } else if (q == 1) {
        val = (m.kind == GSM_KRIG_SIMPLE) ? pv[c] - m.mean : pv[c];
      } else {
        val = gsm_drift(m, q - 2, (px[c] - ox) * isc, (py[c] - oy) * isc, (pz[c] - oz) * isc);
      }
      A[c * BK_LD + n + q] = val;
    }
  // trailing block starts at zero
  for (int q = cy; q < nr; q += 8)
    for (int c = rx; c < nr; c += 16) A[(n + q) * BK_LD + n + c] = 0.0;
  __syncthreads();

  // ---- right-looking Cholesky over the n covariance columns, RHS rows included ------------------
  const int rows = n + nr;
  for (int j = 0; j < n; ++j) {
    const double d = A[j * BK_LD + j];
    const bool okp = d > GSM_PIVOT_RTOL * v.sill;
    if (!okp) {
      if (tid == 0) s_bad = 1;
    }
    const double rs = rsqrt(okp ? d : 1.0);
    // A[r][c] -= l_r l_c for j < c <= r with l = A[.][j] * rs.  Column j is never read again (only the trailing
    // block is needed), so it is not scaled in place: one barrier per column.
    for (int c = j + 1 + cy; c < rows; c += 8) {
      const double lc = A[j * BK_LD + c] * rs;
      for (int r = c + rx; r < rows; r += 16) A[c * BK_LD + r] = fma(-(A[j * BK_LD + r] * rs), lc, A[c * BK_LD + r]);
    }
    __syncthreads();
  }

  // ---- epilogue (one thread): G = B'C^-1B is minus the trailing block ---------------------------
  if (tid == 0) {
    bool singular = s_bad != 0;
    auto G = [&](int a, int bb) { return -A[(n + bb) * BK_LD + n + a]; };   // a >= bb
    const double uu = G(0, 0), wu = G(1, 0);
    const double c0 = v.sill;
    double mean, var;
    if (ncon == 0) {
      mean = m.mean + wu;
      var = c0 - uu;
    } else {
      double Gs[GSM_MAX_DRIFT * GSM_MAX_DRIFT], gq[GSM_MAX_DRIFT], hw[GSM_MAX_DRIFT], y[GSM_MAX_DRIFT], hy[GSM_MAX_DRIFT];
      for (int p = 0; p < ncon; ++p) {
        const double f0 = (m.kind == GSM_KRIG_ORDINARY) ? 1.0
                                                         : gsm_drift(m, p, (tx - ox) * isc, (ty - oy) * isc, (tz - oz) * isc);
        gq[p] = G(2 + p, 0) - f0;
        hw[p] = G(2 + p, 1);
        for (int q = 0; q <= p; ++q) Gs[p * ncon + q] = G(2 + p, 2 + q);
      }
      double gd0[GSM_MAX_DRIFT];   // original diagonal: the singularity test is relative to it
      for (int p0_ = 0; p0_ < ncon; ++p0_) gd0[p0_] = Gs[p0_ * ncon + p0_];
      for (int j = 0; j < ncon; ++j) {
        double d = Gs[j * ncon + j];
        if (!(d > GSM_SCHUR_RTOL * gd0[j])) {
          singular = true;
          d = 1.0;
        }
        const double rs = rsqrt(d);
        Gs[j * ncon + j] = rs;
        for (int i = j + 1; i < ncon; ++i) Gs[i * ncon + j] *= rs;
        for (int i = j + 1; i < ncon; ++i)
          for (int q = j + 1; q <= i; ++q) Gs[i * ncon + q] = fma(-Gs[i * ncon + j], Gs[q * ncon + j], Gs[i * ncon + q]);
      }
      double gnu = 0.0, hnu = 0.0;
      for (int i = 0; i < ncon; ++i) {
        double s1 = gq[i], s2 = hw[i];
        for (int q = 0; q < i; ++q) {
          s1 = fma(-Gs[i * ncon + q], y[q], s1);
          s2 = fma(-Gs[i * ncon + q], hy[q], s2);
        }
        y[i] = s1 * Gs[i * ncon + i];
        hy[i] = s2 * Gs[i * ncon + i];
        gnu = fma(y[i], y[i], gnu);
        hnu = fma(hy[i], y[i], hnu);
      }
      mean = wu - hnu;
      var = c0 - uu + gnu;
    }
    var += 1e-10;  // krig.jl:278-279
    out_mean[t] = singular ? qnan : mean;
    if (out_var) out_var[t] = singular ? qnan : var;
    if (out_status) out_status[t] = singular ? GSM_ST_SINGULAR : GSM_ST_OK;
  }
}

}  // namespace

int gsm_launch_local_krige_big(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered,
                               const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                               double* d_mean, double* d_var, unsigned char* d_status) {
  if (k > BK_KMAX || ncon > GSM_MAX_DRIFT) return GSM_ERR_UNSUPPORTED;
  if (tg.count <= 0) return GSM_OK;
  const size_t smem = ((size_t)BK_ROWS * BK_LD + 4 * BK_KMAX) * sizeof(double);
  GSM_CUDA(ctx, cudaFuncSetAttribute(local_krige_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // one CTA per system; split very large launches so the grid dimension stays in range
  const long long maxgrid = 1LL << 30;
  for (long long off = 0; off < tg.count; off += maxgrid) {
    TargetsDev part = tg;
    part.count = std::min(maxgrid, tg.count - off);
    if (tg.kind == GSM_TARGETS_POINTS) part.points = tg.points + off * tg.dim;
    else part.first = tg.first + off;
    local_krige_big_kernel<<<(unsigned)part.count, BK_THREADS, smem, ctx->stream>>>(
        ctx->bins, v, m, ncon, centered, part, k, minneighbors, d_pos + off * k, d_cnt + off, d_mean + off,
        d_var ? d_var + off : nullptr, d_status ? d_status + off : nullptr);
    ctx->tim.launches++;
  }
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}
