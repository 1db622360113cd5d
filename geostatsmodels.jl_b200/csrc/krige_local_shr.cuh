// Batched local Kriging on the FP64 DMMA path with a SHARED LEADING FACTOR (k <= 32).
// Mathematics as in krige_local_pipe.cuh (reference: fit krig.jl:58-75, weights krig.jl:316-339, krigmean
// krig.jl:245-258, predictvar krig.jl:264-293 per target of models.jl:162-184).  What is new is the grouping:
//
//   * A job is a compact PATCH of 8 grid targets (4 x 2 in 2-D, 2 x 2 x 2 in 3-D).  Neighbouring targets share
//     most of their k nearest samples, and the Cholesky factor of a leading principal block only depends on that
//     block.  Every system of the patch is ordered [samples common to all 8 | its private samples] (each part
//     by ascending record position), the leading 8*S x 8*S block (S <= 3 block columns) is factored ONCE per
//     patch, and each system only factors its private trailing part: 4 - S hand-offs to the diagonal-block
//     warp instead of 4, and 23 instead of 30 DMMA block products when S = 3.
//   * The index work (hash de-duplication of the 8 lists, common samples, ranks) is a separate, fully parallel
//     kernel (krige_local_shr.cu: one warp per job, thousands of warps in flight); this kernel only streams the
//     384-byte job descriptors and the pooled records in, 5 and 4 jobs ahead.
//   * The SHARED FACTOR is a 3-stage pipeline: in the first hand-off round of job j, lanes 8, 9, 10 of the
//     diagonal warp factor the diagonal blocks (0,0) of job j+3, (1,1) of job j+2 and (2,2) of job j+1
//     thread-per-block next to the 8 private blocks; the few DMMA panel / update steps that connect the stages
//     are spread over six system warps at the end of every job.  Nothing of it is on the critical path of the
//     job being solved.
//   * Tried first and dropped: one front-end warp per CTA doing hash + ranks + factor
//     (4x too slow: a lone warp runs dependent instructions at ~6 cycles each on a busy SM), and the index phase
//     inside the system warps' wait slot (its barriers and atomics end up on the critical path).
#include <algorithm>
#include <type_traits>

#include "gsm_internal.cuh"
#include "krige_local_jobs.cuh"

#ifndef GSM_TRACE_DIM3
#define GSM_TRACE_DIM3 0
#endif
#if defined(GSM_TRACE) && GSM_SHR_DIM3 == GSM_TRACE_DIM3 && GSM_SHR_NBC == 4   // developer phase trace: one unit only (k <= 32; 2-D unless GSM_TRACE_DIM3=1)
__device__ long long* g_trace3 = nullptr;
extern "C" __attribute__((visibility("default"))) int gsm_debug_set_trace3(long long* p) {
  return cudaMemcpyToSymbol(g_trace3, &p, sizeof(p)) == cudaSuccess ? 0 : 1;
}
#define TR(slot)                                                                            \
  do {                                                                                      \
    if (g_trace3 && blockIdx.x == 0 && lane == 0 && ji >= 20 && ji < 24)                    \
      g_trace3[((ji - 20) * 10 + warp) * 16 + (slot)] = clock64();                          \
  } while (0)
#else
#define TR(slot) do {} while (0)
#endif

namespace {

constexpr int SW = 8;                      // system warps per CTA
constexpr int WD = SW;                     // diagonal-block warp
constexpr int NTH = (SW + 1) * 32;
constexpr int NSD = NTH;                   // participants of the system <-> diagonal barriers
constexpr int NSYS = SW * 32;
constexpr int EXS = 66;                    // padded stride (doubles) of one exchanged 8x8 block
constexpr int EXL = SW + 3;                // exchange lanes: 8 private blocks + 3 shared-factor stages
constexpr int PMAX = GSM_JOB_PMAX;         // pool capacity (distinct samples per job)
constexpr int PLD = PMAX + 1;              // leading dimension of the pooled matrix (odd: spreads banks)
constexpr int NBJ = 4;                     // pool / shared-factor buffers (pools are loaded 4 jobs ahead)
constexpr int NBH = 8;                     // job descriptors in shared memory (loaded 5 jobs ahead)
constexpr int BAR_READY = 1, BAR_DONE = 2, BAR_POOL = 3, BAR_EPI = 4, BAR_START = 5, BAR_FREE0 = 6;   // (+ BAR_FREE0 + 1)
constexpr int WE = SW + 1;                 // epilogue warp (only with many drift functions, see EPIW below)

__device__ __forceinline__ void dmma(double2& d, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d.x), "+d"(d.y)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ void bar_arrive(int id, int n) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 1/sqrt(d), d > 0: hardware seed (rel. error ~2^-22) + one third-order step -> full double precision
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-(d * y), y, 1.0);
  const double p = e * fma(e, 0.375, 0.5);
  return fma(y, p, y);
}

struct CovFast {
  int kind;
  double sill, c1, k1, k2, inv_r, inv_r2, cs, range2, cfar;
};
__host__ __device__ __forceinline__ CovFast make_cov(const VarioDev& v) {
  CovFast c;
  c.kind = v.kind;
  c.sill = v.sill;
  c.c1 = v.sill - v.nadd;
  c.cs = v.cs;
  c.k1 = -1.5 * v.cs;
  c.k2 = 0.5 * v.cs;
  c.inv_r = v.inv_range;
  c.inv_r2 = v.inv_range * v.inv_range;
  c.range2 = v.range * v.range;
  c.cfar = v.sill - (v.cs + v.nadd);
  return c;
}
// d2 >= 0 clamped to >= ~1e-290 with ONE integer max on the high word (fmax on doubles is ~8 instructions); keeps
// 1/sqrt finite for an exact hit, whose covariance the final select replaces by the sill anyway
__device__ __forceinline__ double clamp_d2(double d2) {
  return __hiloint2double(max(__double2hiint(d2), 0x03e00000), __double2loint(d2));
}
// the same with the variogram kind as a compile-time constant (no branch per evaluation)
template <int KIND>
__device__ __forceinline__ double cov_kind(const CovFast& c, double d2) {
  double v;
  if (KIND == GSM_VARIO_SPHERICAL) {
    const double t2 = d2 * c.inv_r2;
    const double y = fast_rsqrt(clamp_d2(d2));
    const double t = (d2 * c.inv_r) * y;
    v = fma(t, fma(c.k2, t2, c.k1), c.c1);
    v = (d2 < c.range2) ? v : c.cfar;
  } else if (KIND == GSM_VARIO_GAUSSIAN) {
    const double t2 = d2 * c.inv_r2;
    v = c.sill - (c.cs * (1.0 - exp(-3.0 * t2)) + (c.sill - c.c1));
  } else {
    const double y = fast_rsqrt(clamp_d2(d2));
    const double t = (d2 * c.inv_r) * y;
    v = c.sill - (c.cs * (1.0 - exp(-3.0 * t)) + (c.sill - c.c1));
  }
  return (d2 > 0.0) ? v : c.sill;
}
__device__ __forceinline__ double cov_fast(const CovFast& c, double d2) {
  if (c.kind == GSM_VARIO_SPHERICAL) {
    const double t2 = d2 * c.inv_r2;
    const double y = fast_rsqrt(clamp_d2(d2));
    const double t = (d2 * c.inv_r) * y;
    double v = fma(t, fma(c.k2, t2, c.k1), c.c1);
    v = (d2 < c.range2) ? v : c.cfar;
    return (d2 > 0.0) ? v : c.sill;
  } else if (c.kind == GSM_VARIO_GAUSSIAN) {
    const double t2 = d2 * c.inv_r2;
    const double v = c.sill - (c.cs * (1.0 - exp(-3.0 * t2)) + (c.sill - c.c1));
    return (d2 > 0.0) ? v : c.sill;
  } else {
    const double y = fast_rsqrt(clamp_d2(d2));
    const double t = (d2 * c.inv_r) * y;
    const double v = c.sill - (c.cs * (1.0 - exp(-3.0 * t)) + (c.sill - c.c1));
    return (d2 > 0.0) ? v : c.sill;
  }
}

// thread-per-block Cholesky + inverse of an 8x8 SPD block (diagonal warp)
#define LIDX(i, j) ((i) * ((i) + 1) / 2 + (j))
__device__ __forceinline__ bool diag_block(const double* __restrict__ blk, double* __restrict__ out, double thr) {
  double a[36];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) a[LIDX(i, j)] = blk[i * 8 + j];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double d = a[LIDX(j, j)];
    if (!(d > thr)) {
      bad = true;
      d = 1.0;
    }
    const double rs = fast_rsqrt(d);
    a[LIDX(j, j)] = rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[LIDX(i, j)] *= rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i)
#pragma unroll
      for (int k = j + 1; k <= i; ++k) a[LIDX(i, k)] = fma(-a[LIDX(i, j)], a[LIDX(k, j)], a[LIDX(i, k)]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double x[8];
    x[j] = a[LIDX(j, j)];
    out[j * 8 + j] = x[j];
#pragma unroll
    for (int i = j + 1; i < 8; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) s = fma(a[LIDX(i, k)], x[k], s);
      x[i] = -s * a[LIDX(i, i)];
      out[i * 8 + j] = x[i];
      out[j * 8 + i] = 0.0;   // the shared-factor slots are recycled: write the whole block
    }
  }
  return bad;
}

// Cold paths, kept out of line so the hot path stays dense in the instruction cache.
// General gather of two adjacent elements (row, col), (row, col+1) of a system's covariance block:
// pooled with empty slots (identity padding), or direct assembly from the neighbour coordinates (pool overflow).
template <bool DIM3>
__device__ __noinline__ double2 gath_cold(const double* __restrict__ PC, bool pooled, int ir, int ix, int iy, int row,
                                          int col, int n, const double* __restrict__ px,
                                          const double* __restrict__ py, const double* __restrict__ pz,
                                          const VarioDev* vp) {
  double c0 = (row == col) ? vp->sill : 0.0, c1 = (row == col + 1) ? vp->sill : 0.0;   // padding: sill * identity
  if (pooled) {
    if (ir >= 0 && ix >= 0) c0 = PC[ir * PLD + ix];
    if (ir >= 0 && iy >= 0) c1 = PC[ir * PLD + iy];
  } else {
    const CovFast cf = make_cov(*vp);
    const double xr = px[row], yr = py[row], zr = DIM3 ? pz[row] : 0.0;
    const bool rowlive = row < n;
    double dx = xr - px[col], dy = yr - py[col], dz = DIM3 ? zr - pz[col] : 0.0;
    const double e0 = cov_fast(cf, DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx));
    dx = xr - px[col + 1], dy = yr - py[col + 1], dz = DIM3 ? zr - pz[col + 1] : 0.0;
    const double e1 = cov_fast(cf, DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx));
    c0 = (row == col) ? vp->sill : ((rowlive && col < n) ? e0 : 0.0);
    c1 = (row == col + 1) ? vp->sill : ((rowlive && col + 1 < n) ? e1 : 0.0);
  }
  return make_double2(c0, c1);
}
// Pool overflow: the neighbour record of this lane's slot, slots in ascending record position.
__device__ __noinline__ double4 q1_direct(const double4* __restrict__ rec, const int* __restrict__ pos, long long tt,
                                          int k, int n, int lane) {
  int p = 0x7fffffff;
  if (lane < n) p = pos[tt * (long long)k + lane];
  p = gsm_warp_sort_int(p, lane);
  double4 r = make_double4(0.0, 0.0, 0.0, 0.0);
  if (lane < n) r = rec[p];
  return r;
}

template <int NR>
struct alignas(128) ShrSmem {
  static constexpr int NRB = NR > 8 ? 2 : 1;            // right-hand-side block rows
  static constexpr int NGB = NRB * (NRB + 1) / 2;       // Schur blocks handed to the epilogue
  JobIdx J[NBH];                                        // job descriptors
  double exin[EXL * EXS];
  double exout[SW * EXS];
  double px[2][SW][32], py[2][SW][32], pz[2][SW][32];   // neighbour coordinates (direct assembly: pool overflow)
  double rhs[2][SW][NR][32];                            // right-hand-side rows: c0, z, drifts
  double PC[PMAX * PLD];                                // pooled covariances (rows of the non-shared samples)
  double Gx[2][SW * NGB * EXS];                         // trailing Schur blocks handed to the epilogue
  double4 prec[NBJ][PMAX];                              // pooled records {x, y, z, value} by rank
  double fW[NBJ][3][64];                                // inverse diagonal factors of the shared columns
  double fL[NBJ][3][64];                                // shared blocks (1,0), (2,0), (2,1)
  int hbad[NBH][4];                                     // shared diagonal block not positive definite
  int flag[4][SW];                                      // private diagonal-block failures, by job mod 4
};

// NCON: constraints (drift monomials); NBC: 8-wide block columns of the covariance block (k <= 8 NBC)
// EPIW (more than 6 drift functions, i.e. two right-hand-side block rows): the thread-per-system epilogue -- a fully
// unrolled Cholesky of the NCON x NCON drift Schur complement, ~2000 instructions for 10 drifts -- runs in a TENTH
// warp instead of the diagonal-block warp, whose hand-off rounds are on every system warp's critical path.
template <int NCON, bool DIM3, int NBC>
__global__ void __launch_bounds__(NTH + ((2 + NCON > 8) ? 32 : 0), (2 + NCON > 8) ? 1 : 2)
local_krige_shr_kernel(const __grid_constant__ BinsDev b, const __grid_constant__ VarioDev v,
                       const __grid_constant__ ModelDev m, const __grid_constant__ TargetsDev tg,
                       const __grid_constant__ CovFast cf, const JobIdx* __restrict__ jobs, long long ngroups, int k,
                       int centered,
                       const int* __restrict__ pos, double* __restrict__ out_mean, double* __restrict__ out_var,
                       unsigned char* __restrict__ out_status) {
  constexpr int NR = 2 + NCON;
  constexpr int NRB = NR > 8 ? 2 : 1;       // right-hand-side block rows
  constexpr bool EPIW = NR > 8;             // epilogue in its own warp
  constexpr int NTHK = NTH + (EPIW ? 32 : 0);
  constexpr int NGB = NRB * (NRB + 1) / 2;
  constexpr int NROWS = NBC + NRB;          // block rows of the augmented matrix
  constexpr int KMAX = 8 * NBC;             // neighbour slots
  constexpr int SMAX = NBC - 1;             // shared block columns at most (the last column stays private)
  constexpr int S2SPEC = (DIM3 && SMAX >= 2) ? SMAX - 1 : -1;   // second specialised job shape (3-D patches share less)
  static_assert(NR <= 16 && NBC >= 1 && NBC <= 4, "unsupported shape");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ShrSmem<NR>& S = *reinterpret_cast<ShrSmem<NR>*>(smem_raw);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  // jobs of this CTA: patches blockIdx.x, blockIdx.x + gridDim.x, ...
  const int niter = (int)((ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x);

  for (int e = threadIdx.x; e < SW * EXS; e += NTHK) S.exout[e] = 0.0;
  for (int e = threadIdx.x; e < NBH * 4; e += NTHK) (&S.hbad[0][0])[e] = 0;
  __syncthreads();

  // (cf: the covariance constants, a kernel parameter so they are constant-bank operands, not 20 live registers)
  // X <- X W'  (panel solve); two independent accumulators so the two k-chunks overlap
  auto panel = [&](double2& X, const double2& W) {
    double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
    dmma(a0, X.x, W.x);
    dmma(a1, X.y, W.y);
    X = make_double2(a0.x + a1.x, a0.y + a1.y);
  };
  // A -= X Y'
  auto upd = [&](double2& A, const double2& X, const double2& Y) {
    dmma(A, -X.x, Y.x);
    dmma(A, -X.y, Y.y);
  };
  auto frag = [&](const double* src) -> double2 { return *reinterpret_cast<const double2*>(&src[g * 8 + 2 * t]); };
  auto put = [&](double* dst, const double2& val) { *reinterpret_cast<double2*>(&dst[g * 8 + 2 * t]) = val; };
  // covariance block (I,J) of the shared samples of job q (pool ranks 0 .. 8S-1), fragment layout
  auto covs = [&](int q, int I, int J) -> double2 {
    const int qb = q & (NBJ - 1);
    const int r = 8 * I + g, c0 = 8 * J + 2 * t, c1 = c0 + 1;
    const double4* pr = S.prec[qb];
    const double2 pxy = *reinterpret_cast<const double2*>(&pr[r]);
    const double2 axy = *reinterpret_cast<const double2*>(&pr[c0]), bxy = *reinterpret_cast<const double2*>(&pr[c1]);
    const double xr = pxy.x, yr = pxy.y, zr = DIM3 ? pr[r].z : 0.0;
    double dx = xr - axy.x, dy = yr - axy.y, dz = DIM3 ? zr - pr[c0].z : 0.0;
    const double e0 = cov_fast(cf, DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx));
    dx = xr - bxy.x, dy = yr - bxy.y, dz = DIM3 ? zr - pr[c1].z : 0.0;
    const double e1 = cov_fast(cf, DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx));
    return make_double2(r == c0 ? v.sill : e0, r == c1 ? v.sill : e1);
  };
  auto stage_on = [&](int q, int s) -> bool { return q >= 0 && q < niter && S.J[q & (NBH - 1)].S > s; };
  // coordinates of member w's target of a job: grid index from the descriptor (same rounding as
  // gsm_target_coords: (origin + i * spacing) + spacing / 2, every operation rounded once)
  auto target_xyz = [&](const JobIdx& Jq, int w, double& x, double& y, double& z) {
    if (tg.kind == GSM_TARGETS_GRID) {
      x = __dadd_rn(__dadd_rn(tg.origin[0], __dmul_rn((double)Jq.ci[w][0], tg.spacing[0])), __dmul_rn(tg.spacing[0], 0.5));
      y = __dadd_rn(__dadd_rn(tg.origin[1], __dmul_rn((double)Jq.ci[w][1], tg.spacing[1])), __dmul_rn(tg.spacing[1], 0.5));
      z = __dadd_rn(__dadd_rn(tg.origin[2], __dmul_rn((double)Jq.ci[w][2], tg.spacing[2])), __dmul_rn(tg.spacing[2], 0.5));
    } else {
      x = y = z = 0.0;
      if (Jq.tt[w] >= 0) gsm_target_coords(tg, tg.first + Jq.tt[w], x, y, z);
    }
  };

  if (warp == WD || (EPIW && warp == WE)) {
    // ---------------- diagonal-block warp (and, with EPIW, the epilogue warp) ----------------
    // one hand-off round: lanes 0..7 private blocks (when `priv`), lanes 8..10 the shared stages (when `shared`)
    auto round = [&](int j, bool priv, bool shared) __attribute__((always_inline)) {
      const double* in = nullptr;
      double* out = nullptr;
      int* badp = nullptr;
      if (lane < SW) {
        if (priv) {
          in = &S.exin[lane * EXS];
          out = &S.exout[lane * EXS];
          badp = &S.flag[j & 3][lane];
        }
      } else if (lane < SW + 3 && shared && SMAX >= 1) {
        const int s = lane - SW, q = j + 3 - s;
        if (s < SMAX && stage_on(q, s)) {
          in = &S.exin[lane * EXS];
          out = S.fW[q & (NBJ - 1)][s];
          badp = &S.hbad[q & (NBH - 1)][s];
        }
      }
      if (in != nullptr) {
        if (diag_block(in, out, GSM_PIVOT_RTOL * v.sill)) *badp = 1;
      }
    };
    auto epilogue = [&](int j) __attribute__((always_inline)) {
      const JobIdx& Jj = S.J[j & (NBH - 1)];
      const int jh = j & (NBH - 1);
      if (lane < SW && (Jj.fl[lane] & 1)) {
        const double* Gj = S.Gx[j & 1];
        // -G[a][b], a >= b, stored as 8x8 blocks (0,0), (1,0), (1,1)
        auto Gm = [&](int a, int bb) -> double {
          const int br = a >> 3, bc = bb >> 3;
          return Gj[(lane * NGB + br * (br + 1) / 2 + bc) * EXS + (a & 7) * 8 + (bb & 7)];
        };
        const long long tt = Jj.tt[lane];
        bool singular = S.flag[j & 3][lane] != 0 || (S.hbad[jh][0] | S.hbad[jh][1] | S.hbad[jh][2]) != 0;
        const double uu = -Gm(0, 0), wu = -Gm(1, 0);
        const double c0 = v.sill;
        double mean, var;
        if (NCON == 0) {
          mean = m.mean + wu;
          var = c0 - uu;
        } else {
          double Gs[NCON > 0 ? NCON * NCON : 1], gq[NCON > 0 ? NCON : 1], hw[NCON > 0 ? NCON : 1];
#pragma unroll
          for (int p = 0; p < NCON; ++p) {
            gq[p] = -Gm(2 + p, 0);
            hw[p] = -Gm(2 + p, 1);
#pragma unroll
            for (int q = 0; q <= p; ++q) Gs[p * NCON + q] = -Gm(2 + p, 2 + q);
          }
          if (centered) {
#pragma unroll
            for (int p = 0; p < NCON; ++p)
              gq[p] -= (m.exps[p][0] == 0 && m.exps[p][1] == 0 && m.exps[p][2] == 0) ? 1.0 : 0.0;
          } else {
            double tx, ty, tz;
            target_xyz(Jj, lane, tx, ty, tz);
#pragma unroll
            for (int p = 0; p < NCON; ++p) gq[p] -= gsm_drift(m, p, tx, ty, tz);
          }
          double gd0[NCON > 0 ? NCON : 1];   // original diagonal: the singularity test is relative to it
#pragma unroll
          for (int p0_ = 0; p0_ < NCON; ++p0_) gd0[p0_] = Gs[p0_ * NCON + p0_];
#pragma unroll
          for (int j2 = 0; j2 < NCON; ++j2) {
            double d = Gs[j2 * NCON + j2];
            if (!(d > GSM_SCHUR_RTOL * gd0[j2])) {
              singular = true;
              d = 1.0;
            }
            const double rs = rsqrt(d);
            Gs[j2 * NCON + j2] = rs;
#pragma unroll
            for (int i = j2 + 1; i < NCON; ++i) Gs[i * NCON + j2] *= rs;
#pragma unroll
            for (int i = j2 + 1; i < NCON; ++i)
#pragma unroll
              for (int q = j2 + 1; q <= i; ++q)
                Gs[i * NCON + q] = fma(-Gs[i * NCON + j2], Gs[q * NCON + j2], Gs[i * NCON + q]);
          }
          double y[NCON > 0 ? NCON : 1], hy[NCON > 0 ? NCON : 1];
#pragma unroll
          for (int i = 0; i < NCON; ++i) {
            double s1 = gq[i], s2 = hw[i];
#pragma unroll
            for (int q = 0; q < i; ++q) {
              s1 = fma(-Gs[i * NCON + q], y[q], s1);
              s2 = fma(-Gs[i * NCON + q], hy[q], s2);
            }
            y[i] = s1 * Gs[i * NCON + i];
            hy[i] = s2 * Gs[i * NCON + i];
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
        if (Jj.fl[lane] & 2) {
          out_mean[tt] = qnan;
          if (out_var) out_var[tt] = qnan;
          if (out_status) out_status[tt] = GSM_ST_MISSING;
        } else {
          out_mean[tt] = singular ? qnan : mean;
          if (out_var) out_var[tt] = singular ? qnan : var;
          if (out_status) out_status[tt] = singular ? GSM_ST_SINGULAR : GSM_ST_OK;
        }
      }
    };

    if (EPIW && warp == WE) {
      // epilogue warp: job ji's Schur blocks arrive with BAR_EPI; the system warps may refill that Gx buffer (job
      // ji + 2) only after BAR_FREE0 + (ji & 1)
      for (int ji = 0; ji < niter; ++ji) {
        bar_sync(BAR_EPI, NSD);
        epilogue(ji);
        bar_arrive(BAR_FREE0 + (ji & 1), NSYS + 32);
      }
      return;
    }
    // prologue: start the shared-factor pipeline (virtual jobs -4..-1: the system warps prepare, this warp factors)
    for (int jv = -5; jv < 0; ++jv) {
      bar_sync(BAR_START, NTH);
      if (jv < -1) round(jv + 1, false, true);
      __threadfence_block();
      bar_sync(BAR_START, NTH);
    }
    for (int ji = 0; ji < niter; ++ji) {
      if (!EPIW && ji >= 1) bar_sync(BAR_EPI, NSD);   // Schur blocks of job ji-1 are in
      const int S_ = S.J[ji & (NBH - 1)].S;
      if (lane < SW) S.flag[ji & 3][lane] = 0;
      for (int r = 0; r < NBC - S_; ++r) {
        bar_sync(BAR_READY, NSD);
        if (r == 0) TR(0);
        round(ji, true, r == 0);
        bar_arrive(BAR_DONE, NSD);
        if (r == 0) TR(1);
        if (!EPIW && r == 0 && ji >= 1) epilogue(ji - 1);   // previous job (its Schur blocks are double-buffered)
        if (r == 0) TR(2);
      }
    }
    if (!EPIW) {
      bar_sync(BAR_EPI, NSD);
      epilogue(niter - 1);
    }
    return;
  }

  // ============================== system warps ==============================
  auto bar_sync_sys = [&]() { bar_sync(BAR_POOL, NSYS); };

  // ---- streaming the job descriptors (5 jobs ahead) and the pooled records (4 jobs ahead) --------------
  // (asynchronous global -> shared copies: no registers are held while the numeric phase runs)
  auto cp_async = [&](void* dst, const void* src, auto bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    if (decltype(bytes)::value == 4)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
    else
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  };
  auto L0 = [&](int q) __attribute__((always_inline)) {   // descriptor: one int per lane of warps 0..5
    if (warp < GSM_JOB_WORDS / 32) {
      int* dst = reinterpret_cast<int*>(&S.J[q & (NBH - 1)]) + warp * 32 + lane;
      if (q < niter)
        cp_async(dst, reinterpret_cast<const int*>(&jobs[blockIdx.x + (long long)q * gridDim.x]) + warp * 32 + lane,
                 std::integral_constant<int, 4>{});
      else
        *dst = 0;   // past the end: an empty descriptor
    }
    if (warp == 6 && lane < 4) S.hbad[q & (NBH - 1)][lane] = 0;
  };
  auto R0 = [&](int q) __attribute__((always_inline)) {   // pooled record of rank warp*32+lane (warps 0, 1)
    if (warp < 2 && q < niter) {
      const JobIdx& Jq = S.J[q & (NBH - 1)];
      const int r = warp * 32 + lane;
      if (r < Jq.P) {
        const double4* src = &b.rec[Jq.upos[r]];
        double4* dst = &S.prec[q & (NBJ - 1)][r];
        cp_async(dst, src, std::integral_constant<int, 16>{});
        cp_async(reinterpret_cast<char*>(dst) + 16, reinterpret_cast<const char*>(src) + 16, std::integral_constant<int, 16>{});
      }
    }
  };
  auto LR1 = [&]() __attribute__((always_inline)) {        // land both (visible after the next barrier)
    asm volatile("cp.async.wait_all;" ::: "memory");
  };

  // Q1: slot order, coordinates and right-hand-side rows of job q into buffer bf
  auto Q1 = [&](int q, int bf) __attribute__((always_inline)) {
    const int qb = q & (NBJ - 1);
    const JobIdx& Jq = S.J[q & (NBH - 1)];
    const int n = Jq.n[warp];
    const long long tt = Jq.tt[warp];
    const bool live = lane < n;
    double rx = 0.0, ry = 0.0, rz = 0.0, rw = 0.0;
    if (Jq.P >= 0) {
      const int id = Jq.sid[warp][lane];   // slot order from the descriptor (255: empty)
      if (live) {
        const double2 xy = *reinterpret_cast<const double2*>(&S.prec[qb][id]);
        const double2 zw = *reinterpret_cast<const double2*>(&S.prec[qb][id].z);
        rx = xy.x;
        ry = xy.y;
        rz = DIM3 ? zw.x : 0.0;
        rw = zw.y;
      }
    } else {
      // pool overflow: direct assembly from the neighbour records, ascending record position
      const double4 rec = q1_direct(b.rec, pos, tt, k, n, lane);
      rx = rec.x;
      ry = rec.y;
      rz = rec.z;
      rw = rec.w;
      S.px[bf][warp][lane] = rx;
      S.py[bf][warp][lane] = ry;
      if (DIM3) S.pz[bf][warp][lane] = rz;
    }
    double x_tx, x_ty, x_tz;
    target_xyz(Jq, warp, x_tx, x_ty, x_tz);
    const double dx = rx - x_tx, dy = ry - x_ty, dz = DIM3 ? rz - x_tz : 0.0;
    const double d2t = DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx);
    S.rhs[bf][warp][0][lane] = live ? cov_fast(cf, d2t) : 0.0;
    S.rhs[bf][warp][1][lane] = live ? (NCON == 0 ? rw - m.mean : rw) : 0.0;
    if (NCON > 0) {
      double ox = 0.0, oy = 0.0, oz = 0.0, isc = 1.0;
      if (centered) {
        double dmax = live ? d2t : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, gsm_shfl_xor(dmax, o));
        isc = dmax > 0.0 ? rsqrt(dmax) : 1.0;
        ox = x_tx;
        oy = x_ty;
        oz = x_tz;
      }
      const double sx = (rx - ox) * isc, sy = (ry - oy) * isc, sz = (rz - oz) * isc;
#pragma unroll
      for (int q2 = 0; q2 < NCON; ++q2) S.rhs[bf][warp][2 + q2][lane] = live ? gsm_drift(m, q2, sx, sy, sz) : 0.0;
    }
  };
  // Q4 (after the pool barrier): pooled covariances of job q, rows of the non-shared samples
  auto Q4 = [&](int q) __attribute__((always_inline)) {
    const int qb = q & (NBJ - 1);
    const JobIdx& Jq = S.J[q & (NBH - 1)];
    const int P = Jq.P, S8 = 8 * Jq.S;
    const double4* pr = S.prec[qb];
    if (!DIM3) {
      // 2-D patches (P - S8 ~ 14 rows): the full rows of the non-shared samples, 2-3 evaluations per thread;
      // exploiting the symmetry costs more index arithmetic than it saves here (measured)
      const int total = (P - S8) * P;    // (negative for overflow jobs: nothing to do)
      const float invP = 1.0f / (float)max(P, 1);
      auto rows = [&](auto kind_tag) __attribute__((always_inline)) {
        constexpr int KIND = decltype(kind_tag)::value;     // (the variogram kind is compiled in)
#pragma unroll 1
        for (int e = warp * 32 + lane; e < total; e += NSYS) {
          const int rr = (int)(((float)e + 0.5f) * invP);
          const int bc = e - rr * P;
          const int a = S8 + rr;
          double val = v.sill;
          if (a != bc) {
            const double2 axy = *reinterpret_cast<const double2*>(&pr[a]);
            const double2 bxy = *reinterpret_cast<const double2*>(&pr[bc]);
            const double dx = axy.x - bxy.x, dy = axy.y - bxy.y;
            val = cov_kind<KIND>(cf, fma(dy, dy, dx * dx));
          }
          S.PC[a * PLD + bc] = val;
        }
      };
      if (cf.kind == GSM_VARIO_SPHERICAL) rows(std::integral_constant<int, GSM_VARIO_SPHERICAL>{});
      else if (cf.kind == GSM_VARIO_GAUSSIAN) rows(std::integral_constant<int, GSM_VARIO_GAUSSIAN>{});
      else rows(std::integral_constant<int, GSM_VARIO_EXPONENTIAL>{});
    } else {
      // 3-D patches share less (P - S8 ~ 32 rows): columns b <= a only, the rest by symmetry.  Row a' = a - S8 has
      // S8 + a' + 1 entries; pairing row a' with row Pp-1-a' gives rows of constant width W, so (a, b) follows
      // from one float multiply.  (Round 2 tried rows interleaved over the warps with a plain column loop: a third of
      // the index arithmetic but unbalanced, 47.4 -> 48.4 ms on the north-star config.)  The variogram kind is compiled in.
      const int Pp = P - S8, W = 2 * S8 + Pp + 1, total = ((Pp + 1) >> 1) * W;
      const float invW = 1.0f / (float)W;
      auto rows = [&](auto kind_tag) __attribute__((always_inline)) {
        constexpr int KIND = decltype(kind_tag)::value;
#pragma unroll 1
        for (int e = warp * 32 + lane; e < total; e += NSYS) {
          const int rr = (int)(((float)e + 0.5f) * invW);
          const int cc = e - rr * W;
          const bool first = cc <= S8 + rr;
          const int ap = first ? rr : Pp - 1 - rr;
          const int bc = first ? cc : cc - (S8 + rr + 1);
          if (!first && ap == rr) continue;   // odd Pp: the middle row is its own partner
          const int a = S8 + ap;
          double val = v.sill;
          if (a != bc) {
            const double2 axy = *reinterpret_cast<const double2*>(&pr[a]);
            const double2 bxy = *reinterpret_cast<const double2*>(&pr[bc]);
            const double dx = axy.x - bxy.x, dy = axy.y - bxy.y, dz = pr[a].z - pr[bc].z;
            val = cov_kind<KIND>(cf, fma(dz, dz, fma(dy, dy, dx * dx)));
          }
          S.PC[a * PLD + bc] = val;
          if (bc >= S8) S.PC[bc * PLD + a] = val;
        }
      };
      if (cf.kind == GSM_VARIO_SPHERICAL) rows(std::integral_constant<int, GSM_VARIO_SPHERICAL>{});
      else if (cf.kind == GSM_VARIO_GAUSSIAN) rows(std::integral_constant<int, GSM_VARIO_GAUSSIAN>{});
      else rows(std::integral_constant<int, GSM_VARIO_EXPONENTIAL>{});
    }
  };
  // Shared-factor pipeline, system-warp side, run after the last round of job j:
  //  * warps 0..5: the six covariance blocks of the shared samples of job j+4 (its pool has just landed):
  //    (0,0) goes straight to the exchange lane of stage 0, the others wait in the slots their results will take
  //    ((1,0), (2,0), (2,1) in fL; (1,1), (2,2) in fW[1], fW[2]).
  //  * warps 5..7: the DMMA steps that connect the stages, using the inverse diagonal factors of this job's round:
  //    stage 1 of job j+3 (L10, L20, (1,1) - L10 L10') and stage 2 of job j+2 (L21, (2,2) - L20 L20' - L21 L21').
  auto FS = [&](int j) __attribute__((always_inline)) {
    if (SMAX < 1) return;
    if (warp < 6) {
      const int q = j + 4, qb = q & (NBJ - 1);
      const int I = warp < 1 ? 0 : (warp < 3 ? 1 : 2), Jc = warp - I * (I + 1) / 2;   // (0,0) (1,0) (1,1) (2,0) (2,1) (2,2)
      if (I < SMAX && stage_on(q, I)) {
        double* dst = (I == Jc) ? (I == 0 ? &S.exin[SW * EXS] : S.fW[qb][I]) : S.fL[qb][I * (I - 1) / 2 + Jc];
        put(dst, covs(q, I, Jc));
      }
    }
    if (SMAX >= 2 && (warp == 5 || warp == 6)) {
      const int q = j + 3, qb = q & (NBJ - 1);
      if (stage_on(q, 1)) {
        const double2 W0 = frag(S.fW[qb][0]);
        if (warp == 6) {
          double2 L10 = frag(S.fL[qb][0]);
          panel(L10, W0);
          put(S.fL[qb][0], L10);
          double2 A11 = frag(S.fW[qb][1]);
          upd(A11, L10, L10);
          put(&S.exin[(SW + 1) * EXS], A11);
        } else if (SMAX >= 3 && S.J[q & (NBH - 1)].S > 2) {
          double2 L20 = frag(S.fL[qb][1]);
          panel(L20, W0);
          put(S.fL[qb][1], L20);
        }
      }
    }
    if (SMAX >= 3 && warp == 7) {
      const int q = j + 2, qb = q & (NBJ - 1);
      if (stage_on(q, 2)) {
        const double2 W1 = frag(S.fW[qb][1]);
        const double2 L10 = frag(S.fL[qb][0]), L20 = frag(S.fL[qb][1]);
        double2 L21 = frag(S.fL[qb][2]);
        double2 A22 = frag(S.fW[qb][2]);
        upd(L21, L20, L10);
        upd(A22, L20, L20);
        panel(L21, W1);
        put(S.fL[qb][2], L21);
        upd(A22, L21, L21);
        put(&S.exin[(SW + 2) * EXS], A22);
      }
    }
  };

  // ---- job loop.  Jobs -5 .. -1 are virtual: they only stream descriptors / pools in and start the
  //      shared-factor pipeline (so every helper below has a single call site) --------------------------
  for (int ji = -5; ji < niter; ++ji) {
    const bool real = ji >= 0;
    const int jb = ji & (NBJ - 1), buf = ji & 1;
    const JobIdx& Jj = S.J[ji & (NBH - 1)];
    const int S_ = real ? Jj.S : 0;
    const int n = real ? Jj.n[warp] : 0;
    const bool pooled = real ? Jj.P >= 0 : true;
    bool q1_done = false;
    TR(0);
    L0(ji + 5);
    if (ji + 4 >= 0) R0(ji + 4);

    int irow[NBC], rbase[NBC];
    int2 icol[NBC];
    const bool fastg = pooled && n == KMAX;
    // gather block (I,J): I < NBC covariance block, I >= NBC right-hand-side rows
    auto gath = [&](int I, int J, auto fast) -> double2 {
      if (I >= NBC) {
        const int rrow = 8 * (I - NBC) + g;
        double2 val = make_double2(0.0, 0.0);
        if (rrow < NR) val = *reinterpret_cast<const double2*>(&S.rhs[buf][warp][rrow][8 * J + 2 * t]);
        return val;
      }
      if (decltype(fast)::value) return make_double2(S.PC[rbase[I] + icol[J].x], S.PC[rbase[I] + icol[J].y]);
      return gath_cold<DIM3>(S.PC, pooled, irow[I], icol[J].x, icol[J].y, 8 * I + g, 8 * J + 2 * t, n, S.px[buf][warp],
                             S.py[buf][warp], S.pz[buf][warp], &v);
    };
    auto post = [&](const double2& blk) {
      put(&S.exin[warp * EXS], blk);
      bar_arrive(BAR_READY, NSD);
    };
    auto waitW = [&]() -> double2 {
      bar_sync(BAR_DONE, NSD);
      return frag(&S.exout[warp * EXS]);
    };

    // ---- left-looking blocked Cholesky; block rows / columns below S_ come from the shared factor ----
    constexpr int KREL = NBC >= 2 ? NBC - 2 : 0;   // column after whose preparation PC is released
    double2 Lb[NROWS][NBC];
    double2 Gb[NRB][NRB];
    // column kb of the factorisation (kb is a constant after unrolling)
    auto column = [&](int kb, auto fast, auto sfix) __attribute__((always_inline)) {
      // Sx: the number of shared block columns, a compile-time constant on the specialised path
      const int Sx = decltype(sfix)::value >= 0 ? decltype(sfix)::value : S_;
      const bool shcol = kb < Sx;
      // prepare column kb: blocks below the diagonal, and the next diagonal block
      double2 C[NROWS];
#pragma unroll
      for (int i = kb + 1; i < NROWS; ++i) {
        if (i < NBC - 1 && i < Sx) {          // shared block: final already
          C[i] = frag(S.fL[jb][i * (i - 1) / 2 + kb]);
        } else {
          C[i] = gath(i, kb, fast);
#pragma unroll
          for (int j = 0; j < kb; ++j) upd(C[i], Lb[i][j], Lb[kb][j]);
        }
      }
      double2 Dn = make_double2(0.0, 0.0);
      const bool privnext = kb + 1 < NBC && kb + 1 >= Sx;
      if (privnext) {
        Dn = gath(kb + 1, kb + 1, fast);
#pragma unroll
        for (int j = 0; j < kb; ++j) upd(Dn, Lb[kb + 1][j], Lb[kb + 1][j]);
      }
      if (kb == KREL) TR(1);
      if (kb == KREL) bar_sync_sys();   // every gather of this job from the pooled matrix is done
      const double2 W = (kb < SMAX && shcol) ? frag(S.fW[jb][kb < 3 ? kb : 0]) : waitW();
      if (kb + 1 < NBC) {
        if (privnext) {
          panel(C[kb + 1], W);
          upd(Dn, C[kb + 1], C[kb + 1]);
          post(Dn);
        }
        Lb[kb + 1][kb] = C[kb + 1];
      }
#pragma unroll
      for (int i = (kb + 1 < NBC ? kb + 2 : kb + 1); i < NROWS; ++i) {
        if (!(i < NBC - 1 && i < Sx)) panel(C[i], W);
        Lb[i][kb] = C[i];
      }
#pragma unroll
      for (int r = 0; r < NRB; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) upd(Gb[r][q], Lb[NBC + r][kb], Lb[NBC + q][kb]);
    };

    if (real) {
#pragma unroll
      for (int I = 0; I < NBC; ++I) {
        const int ir = Jj.sid[warp][8 * I + g];
        const uchar2 ic = *reinterpret_cast<const uchar2*>(&Jj.sid[warp][8 * I + 2 * t]);
        irow[I] = ir == 255 ? -1 : ir;
        rbase[I] = irow[I] * PLD;
        icol[I] = make_int2(ic.x == 255 ? -1 : (int)ic.x, ic.y == 255 ? -1 : (int)ic.y);
      }
#pragma unroll
      for (int r = 0; r < NRB; ++r)
#pragma unroll
        for (int q = 0; q < NRB; ++q) Gb[r][q] = make_double2(0.0, 0.0);
      // every slot live and pooled (the common case): plain indexed loads, no calls on this path
      using SFull = std::integral_constant<int, SMAX>;
      using SNext = std::integral_constant<int, (S2SPEC >= 0 ? S2SPEC : 0)>;
      using SAny = std::integral_constant<int, -1>;
      if (fastg && SMAX > 0 && S_ == SMAX) {   // the common case, fully specialised: straight-line code
#pragma unroll
        for (int kb = 0; kb <= KREL; ++kb) column(kb, std::true_type{}, SFull{});
      } else if (fastg && S2SPEC >= 0 && S_ == S2SPEC) {   // the common case of 2 x 2 x 2 patches
        // two hand-offs: the next job's right-hand sides (Q1) fill the first wait, its pooled covariances
        // (which must follow this job's last gather from PC) the second
#pragma unroll
        for (int kb = 0; kb <= KREL; ++kb) {
          column(kb, std::true_type{}, SNext{});
          if (NRB == 1 && kb == S2SPEC - 1 && kb < KREL) {   // (measured slower with 16 right-hand-side rows)
            if (ji + 1 < niter) Q1(ji + 1, buf ^ 1);
            q1_done = true;
          }
        }
      } else if (fastg) {
        if (S_ == 0) post(gath(0, 0, std::true_type{}));
#pragma unroll
        for (int kb = 0; kb <= KREL; ++kb) column(kb, std::true_type{}, SAny{});
      } else {
        if (S_ == 0) post(gath(0, 0, std::false_type{}));
#pragma unroll
        for (int kb = 0; kb <= KREL; ++kb) column(kb, std::false_type{}, SAny{});
      }
    }

    // ---- slot (the diagonal warp works on the block just posted): land the streamed descriptor / pool,
    //      then the next job's slot order, right-hand sides and pooled covariances ----
    TR(3);
    LR1();
    TR(4);
    if (ji + 1 >= 0 && ji + 1 < niter) {
      if (!q1_done) Q1(ji + 1, buf ^ 1);
      TR(5);
      Q4(ji + 1);
    }
    TR(6);

    if (real) {
      using SFull = std::integral_constant<int, SMAX>;
      using SNext = std::integral_constant<int, (S2SPEC >= 0 ? S2SPEC : 0)>;
      using SAny = std::integral_constant<int, -1>;
      if (fastg && SMAX > 0 && S_ == SMAX) {
#pragma unroll
        for (int kb = KREL + 1; kb < NBC; ++kb) column(kb, std::true_type{}, SFull{});
      } else if (fastg && S2SPEC >= 0 && S_ == S2SPEC) {
#pragma unroll
        for (int kb = KREL + 1; kb < NBC; ++kb) column(kb, std::true_type{}, SNext{});
      } else if (fastg) {
#pragma unroll
        for (int kb = KREL + 1; kb < NBC; ++kb) column(kb, std::true_type{}, SAny{});
      } else {
#pragma unroll
        for (int kb = KREL + 1; kb < NBC; ++kb) column(kb, std::false_type{}, SAny{});
      }
      // ---- hand the trailing Schur blocks (-B'C^-1B) to the diagonal warp for the epilogue --------
      if (EPIW && ji >= 2) bar_sync(BAR_FREE0 + (ji & 1), NSYS + 32);   // the epilogue warp is done with this buffer
#pragma unroll
      for (int r = 0; r < NRB; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) put(&S.Gx[buf][(warp * NGB + r * (r + 1) / 2 + q) * EXS], Gb[r][q]);
      TR(7);
      bar_arrive(BAR_EPI, NSD);
      __syncwarp();
    }
    if (!real) bar_sync_sys();   // (real jobs: the last hand-off barrier already ordered the pool landing)
    FS(ji);
    if (!real) {   // the diagonal warp runs the shared stages of the virtual round
      __threadfence_block();
      bar_sync(BAR_START, NTH);
      bar_sync(BAR_START, NTH);
    }
    bar_sync_sys();        // next job's pooled covariances, descriptor and pool are complete
    TR(8);
  }
}

template <int NCON, bool DIM3, int NBC>
int launch_shr(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, const TargetsDev& tg, const JobIdx* jobs,
               long long ngroups, int k, int centered, const int* pos, double* mean, double* var,
               unsigned char* status) {
  const size_t smem = sizeof(ShrSmem<2 + NCON>);
  static_assert(sizeof(ShrSmem<2 + NCON>) <= 227 * 1024, "shared memory");
  const int per_sm = ((smem + 1024) * 2 <= 227 * 1024 && 2 + NCON <= 8) ? 2 : 1;
  const int blocks = (int)std::min<long long>(ngroups, 148LL * per_sm);
  auto kern = local_krige_shr_kernel<NCON, DIM3, NBC>;
  GSM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<blocks, NTH + ((2 + NCON > 8) ? 32 : 0), smem, ctx->stream>>>(ctx->bins, v, m, tg, make_cov(v), jobs, ngroups, k, centered, pos, mean, var,
                                           status);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

}  // namespace

// One translation unit per (DIM3, NBC) keeps the build parallel: GSM_SHR_DIM3 / GSM_SHR_NBC select it.
#ifndef GSM_SHR_DIM3
#error "compile through krige_local_shr_*.cu"
#endif
int GSM_SHR_ENTRY(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered, const TargetsDev& tg,
                  const JobIdx* jobs, long long ngroups, int k, const int* d_pos, double* d_mean, double* d_var,
                  unsigned char* d_status) {
#define GSM_LKS_CASE(N) \
  case N:               \
    return launch_shr<N, GSM_SHR_DIM3 != 0, GSM_SHR_NBC>(ctx, v, m, tg, jobs, ngroups, k, centered, d_pos, d_mean, \
                                                         d_var, d_status);
  switch (ncon) {
    GSM_LKS_CASE(0)
    GSM_LKS_CASE(1)
    GSM_LKS_CASE(2)
    GSM_LKS_CASE(3)
    GSM_LKS_CASE(4)
    GSM_LKS_CASE(6)
    GSM_LKS_CASE(10)
  }
#undef GSM_LKS_CASE
  return GSM_ERR_UNSUPPORTED;
}
