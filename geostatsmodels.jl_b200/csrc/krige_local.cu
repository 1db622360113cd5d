// Batched local Kriging: for every target, gather its k neighbours, assemble the covariance system,
// factor, solve, and emit mean and variance.  Replaces the per-target body of fitpredictneigh
// (reference src/models.jl:162-184): fit (krig.jl:58-75: setlhs!/lhsfactorize!), weights
// (krig.jl:316-339), krigmean (krig.jl:245-258, simple.jl:55-69) and predictvar (krig.jl:264-293).
//
// Formulation (one warp per system, k <= 32, lane r owns row r of the covariance block C):
//   The reference factors the saddle-point matrix [C F; F' 0] with Bunch-Kaufman.  Here C (SPD) is
//   Cholesky-factored in registers, and every right-hand side is forward-substituted in the same
//   sweep:  u = L^-1 c0,  w = L^-1 z,  V = L^-1 F.  With  G = V'V,  g = V'u - f0,  G nu = g:
//       mean = w'u - (V'w)' nu            variance = C0 - u'u + g' nu + 1e-10
//   which equals lambda'z and C0 - c0'lambda - f0'nu of the reference (block elimination of the same
//   system), so no back-substitution and no explicit weights are needed.
//   Drift monomials are evaluated on coordinates centred at the target and scaled by the neighbourhood
//   radius when the exponent set is a lower set (then the spanned polynomial space, hence the
//   prediction, is unchanged while G stays well conditioned); otherwise on the raw coordinates.
#include <cstdlib>

#include <cmath>

#include "gsm_internal.cuh"

namespace {

constexpr int LK_WARPS = 4;
constexpr int LK_THREADS = LK_WARPS * 32;
constexpr int LK_LD = 34;  // padded leading dimension of the per-warp covariance tile (doubles)

struct WarpSmem {
  double C[32 * LK_LD];
  double l[32];
  double px[32], py[32], pz[32];
};

// MISS: samples whose value is NaN are `missing` (reference: missingindices / lhsmissings! / rhsmissings!,
// krig.jl:182-209, 381-388, and the SVD minimum-norm solve of krig.jl:157-161).  The reference zeroes the covariance
// rows / columns and right-hand-side entries of those samples but keeps their constraint entries.  Block by block
// that system says (n = samples with a value, m = missing ones, F = drift matrix, nu = Lagrange multipliers):
//     C_nn lambda_n + F_n nu = c_n ,    F_m nu = 0 ,    F_n' lambda_n + F_m' lambda_m = f0 .
// lambda_m multiplies missing values and zeroed right-hand-side entries, so neither mean nor variance sees it; the
// middle equation confines nu to null(F_m).  With nu = N y (N a basis of that null space) the prediction is the
// usual block elimination on the samples that have a value with G, g, h replaced by N'GN, N'g, N'h -- for
// OrdinaryKriging with any missing neighbour that is no constraint at all (SimpleKriging with mean 0), which is what
// the reference's tests pin (test/krig.jl:284-312).  Whenever (lambda_n, nu) is unique this equals the reference's
// minimum-norm solution; where it is not, the system is flagged singular.
template <int NCON, bool MISS>
__global__ void __launch_bounds__(LK_THREADS, 4)
local_krige_kernel(BinsDev b, VarioDev v, ModelDev m, TargetsDev tg, BlockDev bk, int k, int minn, int centered,
                   const int* __restrict__ pos, const int* __restrict__ cnt, double* __restrict__ out_mean,
                   double* __restrict__ out_var, unsigned char* __restrict__ out_status) {
  __shared__ __align__(16) WarpSmem smem[LK_WARPS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long t = blockIdx.x * (long long)LK_WARPS + warp;
  if (t >= tg.count) return;
  WarpSmem& sm = smem[warp];
  constexpr int NR = 2 + NCON;  // right-hand sides: c0, z, drift columns
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  const int n = cnt[t];
  if (n < minn || n <= 0) {
    if (lane == 0) {
      out_mean[t] = qnan;
      if (out_var) out_var[t] = qnan;
      if (out_status) out_status[t] = GSM_ST_MISSING;
    }
    return;
  }
  double tx, ty, tz;
  gsm_target_coords(tg, tg.first + t, tx, ty, tz);

  // ---- gather ------------------------------------------------------------------------------
  bool live = lane < n;
  double4 r = make_double4(0.0, 0.0, 0.0, 0.0);
  {
    int mypos = lane < k ? pos[t * (long long)k + lane] : -1;
    const int key = gsm_warp_sort_int(mypos < 0 ? 0x7fffffff : mypos, lane);   // canonical order
    if (live) r = b.rec[key];
  }
  // a missing neighbour keeps its slot (padding row: sill on the diagonal, zeros elsewhere) and only contributes
  // its drift row to F_m
  const bool miss = MISS && live && (r.w != r.w);
  const unsigned missmask = MISS ? __ballot_sync(0xffffffffu, miss) : 0u;
  const bool nearby = live;          // neighbours that define the drift frame (missing ones included)
  if (MISS && miss) live = false;
  sm.px[lane] = r.x;
  sm.py[lane] = r.y;
  sm.pz[lane] = r.z;
  __syncwarp();

  // ---- assemble C (covariance form), half the pairs per lane, mirrored through shared memory --
  sm.C[lane * LK_LD + lane] = v.sill;   // (padding rows too: the pivot test is relative to the sill)
#pragma unroll 4
  for (int s = 1; s <= 16; ++s) {
    const int c = (lane + s) & 31;
    double val = 0.0;
    if (live && c < n && !((missmask >> c) & 1u)) {
      const double dx = r.x - sm.px[c], dy = r.y - sm.py[c], dz = r.z - sm.pz[c];
      val = gsm_cov_vec(v, dx, dy, dz);
    }
    sm.C[lane * LK_LD + c] = val;
    sm.C[c * LK_LD + lane] = val;
  }
  __syncwarp();
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    const double2 p2 = *reinterpret_cast<const double2*>(&sm.C[lane * LK_LD + c]);
    a[c] = p2.x;
    a[c + 1] = p2.y;
  }

  // ---- right-hand sides ----------------------------------------------------------------------
  double rhs[NR];
  double d2t = 0.0;
  {
    const double dx = r.x - tx, dy = r.y - ty, dz = r.z - tz;
    d2t = dx * dx + dy * dy + dz * dz;
    rhs[0] = live ? (bk.nq > 0 ? gsm_cov_block(v, bk, dx, dy, dz) : gsm_cov_vec(v, dx, dy, dz)) : 0.0;
    rhs[1] = live ? (NCON == 0 ? r.w - m.mean : r.w) : 0.0;
  }
  double f0[NCON > 0 ? NCON : 1];
  double fm[(MISS && NCON > 0) ? NCON : 1];    // drift row of a missing neighbour
  if (NCON > 0) {
    double ox = 0.0, oy = 0.0, oz = 0.0, isc = 1.0;
    if (centered) {
      double dmax = nearby ? d2t : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, gsm_shfl_xor(dmax, o));
      isc = dmax > 0.0 ? rsqrt(dmax) : 1.0;
      ox = tx;
      oy = ty;
      oz = tz;
    }
    const double sx = (r.x - ox) * isc, sy = (r.y - oy) * isc, sz = (r.z - oz) * isc;
    const double gx = (tx - ox) * isc, gy = (ty - oy) * isc, gz = (tz - oz) * isc;
#pragma unroll
    for (int q = 0; q < NCON; ++q) {
      const double fq = gsm_drift(m, q, sx, sy, sz);
      rhs[2 + q] = live ? fq : 0.0;
      if (MISS) fm[q] = miss ? fq : 0.0;
      f0[q] = gsm_drift(m, q, gx, gy, gz);
    }
  }

  // ---- Cholesky of C with fused forward substitution ----------------------------------------
  bool singular = false;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    double d = gsm_shfl(a[j], j);
    if (!(d > GSM_PIVOT_RTOL * v.sill)) {
      singular = true;
      d = 1.0;
    }
    const double rs = rsqrt(d);
    const double l = (lane > j) ? a[j] * rs : 0.0;  // L[r][j], r > j
#pragma unroll
    for (int q = 0; q < NR; ++q) {
      const double yj = gsm_shfl(rhs[q], j) * rs;  // y_j = b_j / L_jj
      rhs[q] = (lane == j) ? yj : fma(-l, yj, rhs[q]);
    }
    if (j < 31) {
      sm.l[lane] = l;
      __syncwarp();
#pragma unroll
      for (int cp = ((j + 1) >> 1) << 1; cp < 32; cp += 2) {
        const double2 l2 = *reinterpret_cast<const double2*>(&sm.l[cp]);
        if (cp > j) a[cp] = fma(-l, l2.x, a[cp]);
        a[cp + 1] = fma(-l, l2.y, a[cp + 1]);
      }
      __syncwarp();
    }
  }

  // ---- inner products, Schur complement, outputs ---------------------------------------------
  const double uu = gsm_warp_sum(rhs[0] * rhs[0]);
  const double wu = gsm_warp_sum(rhs[1] * rhs[0]);
  double mean, var;
  const double c0 = bk.c0;  // covzero (krig.jl:295-301): sill - gamma(0) = sill for a point target
  if (NCON == 0) {
    mean = m.mean + wu;
    var = c0 - uu;
  } else {
    double G[NCON > 0 ? NCON * NCON : 1];
    double g[NCON > 0 ? NCON : 1], hw[NCON > 0 ? NCON : 1];
#pragma unroll
    for (int p = 0; p < NCON; ++p) {
      g[p] = gsm_warp_sum(rhs[2 + p] * rhs[0]) - f0[p];
      hw[p] = gsm_warp_sum(rhs[2 + p] * rhs[1]);
#pragma unroll
      for (int q = 0; q <= p; ++q) G[p * NCON + q] = gsm_warp_sum(rhs[2 + p] * rhs[2 + q]);
    }
    int nc = NCON;   // constraints left after the projection onto null(F_m)
    if (MISS && missmask != 0u) {
      // M = F_m' F_m; symmetric full-pivot Cholesky M[pi, pi] = [L1; L2][L1' L2'] finds its rank r and the null-space
      // basis N[pi] = [-L1^-T L2'; I]; then G, g, h <- N'GN, N'g, N'h (every lane redundantly; local memory)
      constexpr int NC1 = NCON > 0 ? NCON : 1;
      double Mm[NC1 * NC1], Nb[NC1 * NC1], T1[NC1 * NC1];
      int perm[NC1];
      for (int a = 0; a < NCON; ++a) {
        perm[a] = a;
        for (int c = 0; c <= a; ++c) {
          const double sacc = gsm_warp_sum(fm[a] * fm[c]);
          Mm[a * NCON + c] = sacc;
          Mm[c * NCON + a] = sacc;
        }
      }
      double dmax0 = 0.0;
      for (int a = 0; a < NCON; ++a) dmax0 = fmax(dmax0, Mm[a * NCON + a]);
      int rk = 0;
      for (; rk < NCON; ++rk) {
        int piv = rk;
        for (int a = rk + 1; a < NCON; ++a)
          if (Mm[a * NCON + a] > Mm[piv * NCON + piv]) piv = a;
        if (!(Mm[piv * NCON + piv] > 1e-12 * dmax0)) break;
        if (piv != rk) {   // symmetric row / column swap
          for (int a = 0; a < NCON; ++a) {
            const double tmp = Mm[rk * NCON + a];
            Mm[rk * NCON + a] = Mm[piv * NCON + a];
            Mm[piv * NCON + a] = tmp;
          }
          for (int a = 0; a < NCON; ++a) {
            const double tmp = Mm[a * NCON + rk];
            Mm[a * NCON + rk] = Mm[a * NCON + piv];
            Mm[a * NCON + piv] = tmp;
          }
          const int tp = perm[rk];
          perm[rk] = perm[piv];
          perm[piv] = tp;
        }
        const double dd = sqrt(Mm[rk * NCON + rk]);
        Mm[rk * NCON + rk] = dd;
        for (int a = rk + 1; a < NCON; ++a) Mm[a * NCON + rk] /= dd;
        for (int a = rk + 1; a < NCON; ++a)
          for (int c = rk + 1; c <= a; ++c) {
            Mm[a * NCON + c] -= Mm[a * NCON + rk] * Mm[c * NCON + rk];
            Mm[c * NCON + a] = Mm[a * NCON + c];
          }
      }
      nc = NCON - rk;
      // null-space basis, column q (q < nc): N[perm[rk + q]] = 1, N[perm[0..rk)] = -L1^-T (L2 row q)'
      for (int q = 0; q < nc; ++q) {
        for (int a = 0; a < NCON; ++a) Nb[a * NCON + q] = 0.0;
        Nb[perm[rk + q] * NCON + q] = 1.0;
        for (int a = rk - 1; a >= 0; --a) {      // back substitution with L1'
          double sacc = -Mm[(rk + q) * NCON + a];
          for (int c = a + 1; c < rk; ++c) sacc -= Mm[c * NCON + a] * Nb[perm[c] * NCON + q];
          Nb[perm[a] * NCON + q] = sacc / Mm[a * NCON + a];
        }
      }
      // T1 = G N (G symmetric, lower stored), then G <- N' T1, g <- N' g, hw <- N' hw
      for (int a = 0; a < NCON; ++a)
        for (int q = 0; q < nc; ++q) {
          double sacc = 0.0;
          for (int c = 0; c < NCON; ++c) sacc += (c <= a ? G[a * NCON + c] : G[c * NCON + a]) * Nb[c * NCON + q];
          T1[a * NCON + q] = sacc;
        }
      double g2[NC1], h2[NC1];
      for (int q = 0; q < nc; ++q) {
        double sg = 0.0, sh = 0.0;
        for (int a = 0; a < NCON; ++a) {
          sg += Nb[a * NCON + q] * g[a];
          sh += Nb[a * NCON + q] * hw[a];
        }
        g2[q] = sg;
        h2[q] = sh;
      }
      for (int q = 0; q < nc; ++q)
        for (int q2 = 0; q2 <= q; ++q2) {
          double sacc = 0.0;
          for (int a = 0; a < NCON; ++a) sacc += Nb[a * NCON + q] * T1[a * NCON + q2];
          G[q * NCON + q2] = sacc;
        }
      for (int q = 0; q < NCON; ++q) {
        g[q] = q < nc ? g2[q] : 0.0;
        hw[q] = q < nc ? h2[q] : 0.0;
      }
      for (int q = nc; q < NCON; ++q) {   // inert rows: identity
        for (int q2 = 0; q2 < q; ++q2) G[q * NCON + q2] = 0.0;
        G[q * NCON + q] = 1.0;
      }
    }
    // Cholesky of G (every lane redundantly, registers), solve G nu = g
    double gd0[NCON > 0 ? NCON : 1];   // original diagonal: the singularity test is relative to it
#pragma unroll
    for (int p0_ = 0; p0_ < NCON; ++p0_) gd0[p0_] = G[p0_ * NCON + p0_];
#pragma unroll
    for (int j = 0; j < NCON; ++j) {
      double d = G[j * NCON + j];
      if (!(d > GSM_SCHUR_RTOL * gd0[j])) {
        singular = true;
        d = 1.0;
      }
      const double rs = rsqrt(d);
      G[j * NCON + j] = rs;  // store 1/L_jj
#pragma unroll
      for (int i = j + 1; i < NCON; ++i) G[i * NCON + j] *= rs;
#pragma unroll
      for (int i = j + 1; i < NCON; ++i)
#pragma unroll
        for (int q = j + 1; q <= i; ++q) G[i * NCON + q] = fma(-G[i * NCON + j], G[q * NCON + j], G[i * NCON + q]);
    }
    double y[NCON > 0 ? NCON : 1], hy[NCON > 0 ? NCON : 1];
    // forward-substitute both g and hw:  nu'g = |L^-1 g|^2 ,  hw'nu = (L^-1 hw)'(L^-1 g)
#pragma unroll
    for (int i = 0; i < NCON; ++i) {
      double s1 = g[i], s2 = hw[i];
#pragma unroll
      for (int q = 0; q < i; ++q) {
        s1 = fma(-G[i * NCON + q], y[q], s1);
        s2 = fma(-G[i * NCON + q], hy[q], s2);
      }
      y[i] = s1 * G[i * NCON + i];
      hy[i] = s2 * G[i * NCON + i];
    }
    double gnu = 0.0, hnu = 0.0;
#pragma unroll
    for (int i = 0; i < NCON; ++i) {
      gnu = fma(y[i], y[i], gnu);
      hnu = fma(hy[i], y[i], hnu);
    }
    mean = wu - hnu;
    var = c0 - uu + gnu;
  }
  var += 1e-10;  // krig.jl:278-279
  if (lane == 0) {
    out_mean[t] = singular ? qnan : mean;
    if (out_var) out_var[t] = singular ? qnan : var;
    if (out_status) out_status[t] = singular ? GSM_ST_SINGULAR : GSM_ST_OK;
  }
}

template <int NCON, bool MISS = false>
int launch(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, const TargetsDev& tg, int k, int minn, int centered,
           const int* pos, const int* cnt, double* mean, double* var, unsigned char* status) {
  long long blocks = (tg.count + LK_WARPS - 1) / LK_WARPS;
  local_krige_kernel<NCON, MISS><<<(unsigned)blocks, LK_THREADS, 0, ctx->stream>>>(ctx->bins, v, m, tg, ctx->block_cur, k, minn, centered,
                                                                            pos, cnt, mean, var, status);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

// exponent set closed under decrement => monomial span is translation invariant
bool is_lower_set(const ModelDev& m) {
  for (int a = 0; a < m.ndrift; ++a)
    for (int d = 0; d < 3; ++d) {
      if (m.exps[a][d] == 0) continue;
      int e[3] = {m.exps[a][0], m.exps[a][1], m.exps[a][2]};
      e[d]--;
      bool found = false;
      for (int c = 0; c < m.ndrift; ++c)
        if (m.exps[c][0] == e[0] && m.exps[c][1] == e[1] && m.exps[c][2] == e[2]) found = true;
      if (!found) return false;
    }
  return true;
}

}  // namespace

int gsm_launch_local_krige(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m_in, const TargetsDev& tg, int k,
                           int minneighbors, const int* d_pos, const int* d_cnt, double* d_mean, double* d_var,
                           unsigned char* d_status) {
  if (tg.count <= 0) return GSM_OK;
  ModelDev m = m_in;
  int ncon = 0;
  if (m.kind == GSM_KRIG_ORDINARY) {  // OK == UK with the single constant drift (ordinary.jl:33-77)
    ncon = 1;
    m.ndrift = 1;
    m.exps[0][0] = m.exps[0][1] = m.exps[0][2] = 0;
  } else if (m.kind == GSM_KRIG_UNIVERSAL) {
    ncon = m.ndrift;
  }
  bool nonconst = false;
  for (int a = 0; a < m.ndrift; ++a) nonconst |= (m.exps[a][0] | m.exps[a][1] | m.exps[a][2]) != 0;
  const int centered = (ncon > 0 && nonconst && is_lower_set(m)) ? 1 : 0;
  if (ctx->n_missing > 0) {
    // samples with a missing value (NaN): the kernel that restates lhsmissings! / rhsmissings! / the SVD solve
    if (k > 32) {
      gsm_set_error(ctx, "missing values are claimed for maxneighbors <= 32 only");
      return GSM_ERR_UNSUPPORTED;
    }
#define GSM_LKM_CASE(N) \
  case N:               \
    return launch<N, true>(ctx, v, m, tg, k, minneighbors, centered, d_pos, d_cnt, d_mean, d_var, d_status);
    switch (ncon) {
      GSM_LKM_CASE(0)
      GSM_LKM_CASE(1)
      GSM_LKM_CASE(2)
      GSM_LKM_CASE(3)
      GSM_LKM_CASE(4)
      GSM_LKM_CASE(5)
      GSM_LKM_CASE(6)
      GSM_LKM_CASE(7)
      GSM_LKM_CASE(8)
      GSM_LKM_CASE(9)
      GSM_LKM_CASE(10)
    }
#undef GSM_LKM_CASE
    return GSM_ERR_UNSUPPORTED;
  }
  if (k > 32)   // 32 < k <= 64: the matrix lives in shared memory (krige_local_mid.cu)
    return gsm_launch_local_krige_mid(ctx, v, m, ncon, centered, tg, k, minneighbors, d_pos, d_cnt, d_mean, d_var,
                                      d_status);
  // Default: the shared-factor kernel for grid targets with more than one block column (k > 8: neighbouring
  // targets share most neighbours, krige_local_shr.cuh), the pipelined kernel otherwise (k <= 8, or explicit
  // points whose order says nothing about their proximity).  gsm_set_option(ctx, "local_impl", ...) overrides
  // (A/B profiling and the per-kernel parity tests); no environment variable is read on a product call.
  bool patches = tg.kind == GSM_TARGETS_GRID && tg.dim >= 2 && tg.dims[1] > 1;
  if (patches) {
    // Sharing pays when the neighbourhood is wide compared with a patch: radius of the k nearest samples in units
    // of the target spacing, from the mean sample density of the bins.  Measured crossover (tools/ab_density_sweep.py,
    // tools/run_configs.py): ~5 cells in 2-D (4 x 2 patches), ~3 in 3-D (2 x 2 x 2); below it the pools overflow
    // or share little and the per-system pipelined kernel is faster.
    const BinsDev& bn = ctx->bins;
    const int d = tg.dim >= 3 ? 3 : 2;
    double cells = (double)bn.nc[0] * bn.nc[1] * (d == 3 ? bn.nc[2] : 1);
    double vol = cells * std::pow(bn.h, d);
    double cellvol = std::fabs(tg.spacing[0] * tg.spacing[1] * (d == 3 ? tg.spacing[2] : 1.0));
    if (vol > 0.0 && cellvol > 0.0 && ctx->n > 0) {
      const double s = (double)ctx->n / vol * cellvol;   // samples per target cell
      const double rk = d == 3 ? std::cbrt(3.0 * k / (4.0 * M_PI * s)) : std::sqrt(k / (M_PI * s));
      if (rk < (d == 3 ? 3.0 : 5.0)) patches = false;
    }
  }
  int impl = ctx->opt_local_impl;
  if (v.general) impl = GSM_IMPL_SIMT;   // composite / anisotropic / newer families: the general covariance function
  if (ctx->block_cur.nq > 0) impl = GSM_IMPL_SIMT;   // block targets: right-hand sides averaged over the block's points
  if (impl == GSM_IMPL_AUTO) impl = (k > 8 && patches) ? GSM_IMPL_SHR : GSM_IMPL_PIPE;
  if (impl == GSM_IMPL_PATCH || impl == GSM_IMPL_PATCH2) {
    int rc = gsm_launch_local_krige_patch(ctx, v, m, ncon, centered, tg, k, minneighbors, d_pos, d_cnt, d_mean, d_var,
                                          d_status);
    if (rc != GSM_ERR_UNSUPPORTED) return rc;
    impl = GSM_IMPL_PIPE;
  }
  if (impl == GSM_IMPL_SHR) {
    int rc = gsm_launch_local_krige_shr(ctx, v, m, ncon, centered, tg, k, minneighbors, d_pos, d_cnt, d_mean, d_var,
                                        d_status);
    if (rc != GSM_ERR_UNSUPPORTED) return rc;
  }
  if (impl == GSM_IMPL_PIPE || impl == GSM_IMPL_SHR) {
    int rc = gsm_launch_local_krige_pipe(ctx, v, m, ncon, centered, tg, k, minneighbors, d_pos, d_cnt, d_mean, d_var,
                                         d_status);
    if (rc != GSM_ERR_UNSUPPORTED) return rc;
  }
#define GSM_LK_CASE(N) \
  case N:              \
    return launch<N>(ctx, v, m, tg, k, minneighbors, centered, d_pos, d_cnt, d_mean, d_var, d_status);
  switch (ncon) {
    GSM_LK_CASE(0)
    GSM_LK_CASE(1)
    GSM_LK_CASE(2)
    GSM_LK_CASE(3)
    GSM_LK_CASE(4)
    GSM_LK_CASE(5)
    GSM_LK_CASE(6)
    GSM_LK_CASE(7)
    GSM_LK_CASE(8)
    GSM_LK_CASE(9)
    GSM_LK_CASE(10)
  }
#undef GSM_LK_CASE
  gsm_set_error(ctx, "unsupported number of drift functions");
  return GSM_ERR_UNSUPPORTED;
}
