// Batched local Kriging, ONE WARP PER PATCH of 8 grid targets (k <= 32): the patch-level block elimination.
// Reference path replaced: per target of models.jl:162-184 -> fit (krig.jl:58-75), weights (krig.jl:316-339),
// krigmean (krig.jl:245-258), predictvar (krig.jl:264-293); constraints ordinary.jl:33-77, universal.jl:81-137.
//
// The shared-factor kernel (krige_local_shr.cuh) factors the samples common to the 8 systems of a patch once, but
// every system still solves its own panel against those shared columns and carries its own right-hand sides.  Here
// the whole patch is ONE augmented block matrix
//
//        rows                                   columns eliminated
//        [ shared samples        (8 S)       ]  <- the S leading block columns: Cholesky, once per patch
//        [ other pooled samples  (P - 8 S)   ]
//        [ c0 rows: one per target (8)       ]     right-hand sides ride along as extra ROWS
//        [ z row, drift rows     (1 + NCON)  ]     (shared by all 8 targets)
//
// After a right-looking blocked elimination of the S shared block columns (DMMA m8n8k4, blocks in the accumulator
// layout, lane-private shared-memory storage: no synchronisation), the trailing matrix T holds, for ALL targets at
// once, the Schur complement over the non-shared samples and every inner product with the right-hand sides.
// System w then only needs its own <= 16 private samples: a principal sub-matrix of T (gathered by pool rank), one
// or two 8 x 8 diagonal blocks (Cholesky + inverse thread-per-block, the 8 targets of the patch side by side), one
// panel product and the small Schur complement.  Per patch that is ~120 DMMA instead of 8 x 46, one evaluation of
// the pooled covariances instead of per-CTA pooling, and no hand-offs between warps at all: a warp never waits for
// another warp, latency is hidden by the other resident warps.
//
// Jobs the patch form does not cover (pool overflow, fewer than k - 16 common samples, pools that do not fit the
// block store) take the GENERIC per-target path in the same warp: the same fill + eliminate routines on the
// target's own k x k system.
//
// Mathematics (as in krige_local_pipe.cuh): C = L L', u = L^-1 c0, w = L^-1 z, V = L^-1 F; the trailing block of
// the augmented matrix after eliminating C is -B' C^-1 B, B = [c0 z F]: every inner product the prediction needs.
// Drift monomials are evaluated on coordinates centred at (one target of) the patch and scaled by the pool radius
// whenever the exponent set is a lower set (same polynomial space, well-conditioned Schur complement).
#include <algorithm>
#include <type_traits>

#include "gsm_internal.cuh"
#include "krige_cov_fast.cuh"
#include "krige_local_jobs.cuh"

namespace gsm_patch {

using gsm_cov::CovFast;
using gsm_cov::cov_fast;
using gsm_cov::fast_rsqrt;

struct PatchParams {
  BinsDev b;
  VarioDev v;
  ModelDev m;
  TargetsDev tg;
  CovFast cf;
  const JobIdx* jobs;
  long long ngroups;
  int k, nbc, centered;
  int drift_const;          // the only drift is the constant (OrdinaryKriging): no monomial evaluation
  int nbmax, nblk;          // block rows the block store can hold, blocks in the store (nbmax (nbmax + 1) / 2)
  int warp_bytes;           // shared memory per warp
  const int* pos;
  double* mean;
  double* var;
  unsigned char* status;
};

__device__ __forceinline__ void dmma(double2& d, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d.x), "+d"(d.y)
      : "d"(a), "d"(b));
}
// A -= X Y'   (blocks in the accumulator layout: lane (g,t) holds [g][2t], [g][2t+1])
__device__ __forceinline__ void upd(double2& A, const double2& X, const double2& Y) {
  dmma(A, -X.x, Y.x);
  dmma(A, -X.y, Y.y);
}
// X <- X W'
__device__ __forceinline__ void panel(double2& X, const double2& W) {
  double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
  dmma(a0, X.x, W.x);
  dmma(a1, X.y, W.y);
  X = make_double2(a0.x + a1.x, a0.y + a1.y);
}
__device__ __forceinline__ int bidx(int i, int j) { return i * (i + 1) / 2 + j; }   // lower block (i, j), j <= i

// Warp-cooperative inverse Cholesky factor of an 8x8 SPD block in the accumulator layout: Gaussian elimination
// without pivoting on [A | I] gives the unit-lower inverse M; scaling row g by 1/sqrt(pivot g) gives W = L^-1.
__device__ __forceinline__ double2 inv_chol_warp(double2 a, int g, int t, double thr, bool& bad) {
  double2 m = make_double2(g == 2 * t ? 1.0 : 0.0, g == 2 * t + 1 ? 1.0 : 0.0);
  double myrs = 1.0;
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const double ac = (j & 1) ? a.y : a.x;
    double d = gsm_shfl(ac, j * 4 + (j >> 1));            // pivot A[j][j]
    const double arj = gsm_shfl(ac, g * 4 + (j >> 1));    // A[g][j]
    const double pj0 = gsm_shfl(a.x, j * 4 + t), pj1 = gsm_shfl(a.y, j * 4 + t);   // pivot row of A
    const double mj0 = gsm_shfl(m.x, j * 4 + t), mj1 = gsm_shfl(m.y, j * 4 + t);   // pivot row of M
    if (!(d > thr)) {
      bad = true;
      d = 1.0;
    }
    const double rs = fast_rsqrt(d);
    if (g == j) myrs = rs;
    const double f = (g > j) ? -((arj * rs) * rs) : 0.0;
    a.x = fma(f, pj0, a.x);
    a.y = fma(f, pj1, a.y);
    m.x = fma(f, mj0, m.x);
    m.y = fma(f, mj1, m.y);
  }
  return make_double2(m.x * myrs, m.y * myrs);
}

// Thread-per-block Cholesky + inverse of an 8x8 SPD block, in place (row-major, lower triangle read, whole block
// written: W = L^-1, zeros above the diagonal).  Eight targets of a patch are factored side by side, one per lane.
#define GSM_LIDX(i, j) ((i) * ((i) + 1) / 2 + (j))
__device__ __forceinline__ bool diag_block_inplace(double* __restrict__ blk, double thr) {
  double a[36];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) a[GSM_LIDX(i, j)] = blk[i * 8 + j];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double d = a[GSM_LIDX(j, j)];
    if (!(d > thr)) {
      bad = true;
      d = 1.0;
    }
    const double rs = fast_rsqrt(d);
    a[GSM_LIDX(j, j)] = rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[GSM_LIDX(i, j)] *= rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i)
#pragma unroll
      for (int k = j + 1; k <= i; ++k) a[GSM_LIDX(i, k)] = fma(-a[GSM_LIDX(i, j)], a[GSM_LIDX(k, j)], a[GSM_LIDX(i, k)]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double x[8];
    x[j] = a[GSM_LIDX(j, j)];
    blk[j * 8 + j] = x[j];
#pragma unroll
    for (int i = j + 1; i < 8; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) s = fma(a[GSM_LIDX(i, k)], x[k], s);
      x[i] = -s * a[GSM_LIDX(i, i)];
      blk[i * 8 + j] = x[i];
      blk[j * 8 + i] = 0.0;
    }
  }
  return bad;
}
#undef GSM_LIDX

// sill - gamma(h) from a squared distance, variogram kind as a compile-time constant (krige_cov_fast.cuh has the
// run-time form).  d2 = 0 is an exact hit (covariance = sill); the clamp keeps 1/sqrt finite for it.
template <int KIND>
__device__ __forceinline__ double cov_eval(const CovFast& c, double d2) {
  // d2 >= 0: an integer max on the high word clamps it to >= ~1e-290 (one instruction)
  const double d2c = __hiloint2double(max(__double2hiint(d2), 0x03e00000), __double2loint(d2));
  double v;
  if (KIND == GSM_VARIO_SPHERICAL) {
    const double t2 = d2c * c.inv_r2;
    const double y = fast_rsqrt(d2c);
    const double t = (d2c * c.inv_r) * y;
    v = fma(t, fma(c.k2, t2, c.k1), c.c1);
    v = (d2c < c.range2) ? v : c.cfar;
  } else if (KIND == GSM_VARIO_GAUSSIAN) {
    const double t2 = d2c * c.inv_r2;
    v = c.sill - (c.cs * (1.0 - exp(-3.0 * t2)) + (c.sill - c.c1));
  } else {
    const double y = fast_rsqrt(d2c);
    const double t = (d2c * c.inv_r) * y;
    v = c.sill - (c.cs * (1.0 - exp(-3.0 * t)) + (c.sill - c.c1));
  }
  return (d2 > 0.0) ? v : c.sill;
}

__device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }

// NCON: constraints (drift monomials); NW: warps that share one patch (1: no block-level synchronisation at all;
// 2: block rows / members are dealt to the two warps by parity, halving the latency of a patch at the price of
// duplicated serial work -- measured slower at C3, kept as the "patch2" developer option)
template <int NCON, bool DIM3, int NW>
__global__ void __launch_bounds__(32 * NW, NW == 1 ? 10 : 8)
local_krige_patch_kernel(const __grid_constant__ PatchParams p) {
  auto cta_sync = [&]() {
    if (NW > 1) __syncthreads();
    else __syncwarp();
  };
  auto mine = [&](int x, int hh) -> bool { return NW == 1 || (x & 1) == hh; };
  constexpr int STEP2 = NW * (NW + 1) / 2;     // block (i + NW, j) = block (i, j) + NW * i + STEP2
  constexpr int NR = 2 + NCON;                 // right-hand-side rows of one system: c0, z, drifts
  constexpr int NRB = NR > 8 ? 2 : 1;          // ... in 8-row blocks
  constexpr int NGB = NRB * (NRB + 1) / 2;
  constexpr int NZR = 1 + NCON;                // rows shared by the patch: z, drifts
  constexpr int NZB = NZR > 8 ? 2 : 1;
  constexpr int GS = NR * NR;                  // trailing Schur entries handed to the epilogue, per target
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double4* prec = reinterpret_cast<double4*>(smem_raw);           // pooled records by rank (64)
  double* blk = reinterpret_cast<double*>(prec + GSM_JOB_PMAX);   // block store, 64 doubles per 8x8 block
  double* gout = blk + (size_t)p.nblk * 64;                       // 8 x GS
  double4* srec = reinterpret_cast<double4*>(blk + (size_t)(p.nblk - 2) * 64);   // generic path: own records (32),
                                                                                 // in the last two blocks of the store

  const int lane = threadIdx.x & 31;
  const int h = NW == 1 ? 0 : (int)(threadIdx.x >> 5);   // which half of the patch this warp takes (NW = 2)
  const int g = lane >> 2, t = lane & 3;
  const int fo = g * 8 + 2 * t;                // this lane's two elements of every block
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const VarioDev& v = p.v;
  const ModelDev& m = p.m;
  const CovFast& cf = p.cf;
  const double thr = GSM_PIVOT_RTOL * v.sill;
  const int nbc = p.nbc;

  auto ld2 = [&](const double* q) -> double2 { return *reinterpret_cast<const double2*>(q); };
  auto st2 = [&](double* q, const double2& val) { *reinterpret_cast<double2*>(q) = val; };
  auto d2of = [&](double ax, double ay, double az, double bx, double by, double bz) -> double {
    const double dx = ax - bx, dy = ay - by, dz = DIM3 ? az - bz : 0.0;
    return DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx);
  };
  // drift monomial n of a point in the drift frame (OrdinaryKriging: the constant)
  auto driftv = [&](int n, double x, double y, double z) -> double {
    return p.drift_const ? 1.0 : gsm_drift(m, n, x, y, z);
  };
  // Right-looking blocked elimination of the leading `ncols` block columns of a lower block matrix with NBt block
  // rows (block (i, j) at (tri(i) + j) * 64).  Every lane only touches its own two elements of every block: no
  // synchronisation inside.  Running pointers: block (i + 1, j) = block (i, j) + (i + 1).
  auto eliminate = [&](int ncols, int NBt) -> bool {
    bool bad = false;
    for (int kb = 0; kb < ncols; ++kb) {
      double* dk = blk + (tri(kb) + kb) * 64 + fo;
      const double2 W = inv_chol_warp(ld2(dk), g, t, thr, bad);     // (both warps, redundantly)
      // block rows are dealt to the two warps by parity: warp h owns rows i = h (mod 2)
      {
        const int i0 = NW == 1 ? kb + 1 : kb + 1 + ((kb + 1 + h) & 1);
        double* pl = blk + (tri(i0) + kb) * 64 + fo;
        int i = i0;
        for (; i + NW < NBt; i += 2 * NW) {
          const int d1 = (NW * i + STEP2) * 64;
          double2 X0 = ld2(pl), X1 = ld2(pl + d1);
          panel(X0, W);
          panel(X1, W);
          st2(pl, X0);
          st2(pl + d1, X1);
          pl += d1 + (NW * (i + NW) + STEP2) * 64;
        }
        if (i < NBt) {
          double2 X = ld2(pl);
          panel(X, W);
          st2(pl, X);
        }
      }
      cta_sync();
      double* lj = dk + (kb + 1) * 64;
      for (int j = kb + 1; j < NBt; ++j) {
        const double2 Lj = ld2(lj);
        const double nx = -Lj.x, ny = -Lj.y;
        const int i0 = NW == 1 ? j : j + ((j + h) & 1);
        double* a = blk + (tri(i0) + j) * 64 + fo;
        const double* li = blk + (tri(i0) + kb) * 64 + fo;
        int i = i0;
        for (; i + NW < NBt; i += 2 * NW) {            // two independent blocks per step
          const int d1 = (NW * i + STEP2) * 64;
          double2 A0 = ld2(a), A1 = ld2(a + d1);
          const double2 L0 = ld2(li), L1 = ld2(li + d1);
          dmma(A0, L0.x, nx);
          dmma(A1, L1.x, nx);
          dmma(A0, L0.y, ny);
          dmma(A1, L1.y, ny);
          st2(a, A0);
          st2(a + d1, A1);
          const int d2 = d1 + (NW * (i + NW) + STEP2) * 64;
          a += d2;
          li += d2;
        }
        if (i < NBt) {
          double2 A = ld2(a);
          const double2 Li = ld2(li);
          dmma(A, Li.x, nx);
          dmma(A, Li.y, ny);
          st2(a, A);
        }
        lj += (j + 1) * 64;
      }
      cta_sync();
    }
    return bad;
  };

  for (long long job = blockIdx.x; job < p.ngroups; job += gridDim.x) {
    const JobIdx& J = p.jobs[job];
    const int P = J.P, S = J.S;
    int n_l = 0, fl_l = 0, tt_l = -1;
    double tx = 0.0, ty = 0.0, tz = 0.0;       // lanes 0..7: the target of member `lane`
    if (lane < 8) {
      n_l = J.n[lane];
      fl_l = J.fl[lane];
      tt_l = J.tt[lane];
      if (fl_l & 1) {
        if (p.tg.kind == GSM_TARGETS_GRID) {
          tx = __dadd_rn(__dadd_rn(p.tg.origin[0], __dmul_rn((double)J.ci[lane][0], p.tg.spacing[0])), __dmul_rn(p.tg.spacing[0], 0.5));
          ty = __dadd_rn(__dadd_rn(p.tg.origin[1], __dmul_rn((double)J.ci[lane][1], p.tg.spacing[1])), __dmul_rn(p.tg.spacing[1], 0.5));
          tz = __dadd_rn(__dadd_rn(p.tg.origin[2], __dmul_rn((double)J.ci[lane][2], p.tg.spacing[2])), __dmul_rn(p.tg.spacing[2], 0.5));
        } else {
          gsm_target_coords(p.tg, p.tg.first + tt_l, tx, ty, tz);
        }
      }
    }
    const bool act_l = lane < 8 && (fl_l & 1) && !(fl_l & 2) && n_l > 0;
    const unsigned actmask = __ballot_sync(0xffffffffu, act_l);
    // per-member results of the solve, kept by lane w < 8 for the epilogue
    bool sing_l = false;
    double ecx = 0.0, ecy = 0.0, ecz = 0.0, eisc = 1.0;   // drift frame used for member `lane`

    const int NO = P >= 0 ? (P - 8 * S + 7) >> 3 : 0;
    const int NS = S + NO, NB = NS + 1 + NZB;
    const int npb = nbc - S;                                          // private block columns per system
    const int per_target = npb <= 1 ? 1 : 1 + NRB + NGB;              // scratch blocks per target in stage C
    const int ndead = S * (S + 1) / 2 + (NB - S) * S;                 // blocks of the eliminated columns
    const bool fast = P >= 0 && S >= 1 && npb <= 2 && NB <= p.nbmax && ndead >= per_target && actmask != 0u;

    if (actmask != 0u && P >= 0) {
      for (int r = threadIdx.x; r < GSM_JOB_PMAX; r += 32 * NW)
        prec[r] = r < P ? p.b.rec[J.upos[r]] : make_double4(0.0, 0.0, 0.0, 0.0);
    }
    cta_sync();

    if (fast) {
      // ---------------- drift frame of the patch: centre = first active target, scale = pool radius ----------
      double cx = 0.0, cy = 0.0, cz = 0.0, isc = 1.0;
      if (NCON > 0 && p.centered) {
        const int w0 = __ffs(actmask) - 1;
        cx = __shfl_sync(0xffffffffu, tx, w0);
        cy = __shfl_sync(0xffffffffu, ty, w0);
        cz = __shfl_sync(0xffffffffu, tz, w0);
        double dmax = 0.0;
        for (int r = lane; r < P; r += 32) dmax = fmax(dmax, d2of(prec[r].x, prec[r].y, prec[r].z, cx, cy, cz));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, gsm_shfl_xor(dmax, o));
        isc = dmax > 0.0 ? rsqrt(dmax) : 1.0;
      }
      ecx = cx, ecy = cy, ecz = cz, eisc = isc;
      // target of member g (rows of the c0 block row)
      const double gx = __shfl_sync(0xffffffffu, tx, g), gy = __shfl_sync(0xffffffffu, ty, g),
                   gz = __shfl_sync(0xffffffffu, tz, g);
      const bool gact = (actmask >> g) & 1u;

      // ---------------- fill the lower block matrix (variogram kind compiled in) ------------------------------
      auto fill = [&](auto kind_tag) __attribute__((always_inline)) {
        constexpr int KIND = decltype(kind_tag)::value;
        int turn = h;                                  // blocks are dealt to the two warps alternately
        double* rowp = blk + fo;                       // this lane's slot of block (i, 0)
        // rows of pooled samples.  Off-diagonal blocks (j < i) never have padded columns (8 j + 8 <= 8 i < P).
        for (int i = 0; i < NS; ++i) {
          const int r = 8 * i + g;
          const double2 rxy = ld2(&prec[r].x);
          const double rz = DIM3 ? prec[r].z : 0.0;
          const bool redge = 8 * i + 8 > P;            // (warp-uniform) the last row block may hold padding rows
          const double4* pc = prec + 2 * t;
          double* dst = rowp;
          // (this warp's blocks of the row: every second one)
          int j = NW == 1 ? 0 : turn;
          pc += 8 * j;
          dst += 64 * j;
          for (; j + NW < i; j += 2 * NW) {            // two blocks at a time: four independent chains
            const double2 a0 = ld2(&pc[0].x), a1 = ld2(&pc[1].x);
            const double2 b0 = ld2(&pc[8 * NW].x), b1 = ld2(&pc[8 * NW + 1].x);
            const double az0 = DIM3 ? pc[0].z : 0.0, az1 = DIM3 ? pc[1].z : 0.0;
            const double bz0 = DIM3 ? pc[8 * NW].z : 0.0, bz1 = DIM3 ? pc[8 * NW + 1].z : 0.0;
            double e0 = cov_eval<KIND>(cf, d2of(rxy.x, rxy.y, rz, a0.x, a0.y, az0));
            double e1 = cov_eval<KIND>(cf, d2of(rxy.x, rxy.y, rz, a1.x, a1.y, az1));
            double f0 = cov_eval<KIND>(cf, d2of(rxy.x, rxy.y, rz, b0.x, b0.y, bz0));
            double f1 = cov_eval<KIND>(cf, d2of(rxy.x, rxy.y, rz, b1.x, b1.y, bz1));
            if (redge && r >= P) e0 = e1 = f0 = f1 = 0.0;
            st2(dst, make_double2(e0, e1));
            st2(dst + 64 * NW, make_double2(f0, f1));
            pc += 16 * NW;
            dst += 128 * NW;
          }
          for (; j < i; j += NW) {
            const double2 a0 = ld2(&pc[0].x), a1 = ld2(&pc[1].x);
            double e0 = cov_eval<KIND>(cf, d2of(rxy.x, rxy.y, rz, a0.x, a0.y, DIM3 ? pc[0].z : 0.0));
            double e1 = cov_eval<KIND>(cf, d2of(rxy.x, rxy.y, rz, a1.x, a1.y, DIM3 ? pc[1].z : 0.0));
            if (redge && r >= P) e0 = e1 = 0.0;
            st2(dst, make_double2(e0, e1));
            pc += 8 * NW;
            dst += 64 * NW;
          }
          turn = (turn + i + 1) & 1;                   // i + 1 blocks in this row (the diagonal one included)
          if (j == i) {
            const int c0 = r - g + 2 * t;             // 8 i + 2 t
            const double2 a0 = ld2(&pc[0].x), a1 = ld2(&pc[1].x);
            double e0 = cov_eval<KIND>(cf, d2of(rxy.x, rxy.y, rz, a0.x, a0.y, DIM3 ? pc[0].z : 0.0));
            double e1 = cov_eval<KIND>(cf, d2of(rxy.x, rxy.y, rz, a1.x, a1.y, DIM3 ? pc[1].z : 0.0));
            if (redge) {                               // padding: sill * identity
              e0 = (r < P && c0 < P) ? e0 : 0.0;
              e1 = (r < P && c0 + 1 < P) ? e1 : 0.0;
            }
            e0 = g == 2 * t ? v.sill : e0;
            e1 = g == 2 * t + 1 ? v.sill : e1;
            st2(dst, make_double2(e0, e1));
          }
          rowp += (i + 1) * 64;
        }
        // c0 rows: row g = target of member g
        {
          const int j0 = NW == 1 ? 0 : turn;
          const double4* pc = prec + 2 * t + 8 * j0;
          double* dst = rowp + 64 * j0;
          int j = j0;
          for (; j + NW < NS; j += 2 * NW) {           // two blocks at a time (neither is the last, padded one)
            const double2 a0 = ld2(&pc[0].x), a1 = ld2(&pc[1].x);
            const double2 b0 = ld2(&pc[8 * NW].x), b1 = ld2(&pc[8 * NW + 1].x);
            const double az0 = DIM3 ? pc[0].z : 0.0, az1 = DIM3 ? pc[1].z : 0.0;
            const double bz0 = DIM3 ? pc[8 * NW].z : 0.0, bz1 = DIM3 ? pc[8 * NW + 1].z : 0.0;
            double e0 = cov_eval<KIND>(cf, d2of(gx, gy, gz, a0.x, a0.y, az0));
            double e1 = cov_eval<KIND>(cf, d2of(gx, gy, gz, a1.x, a1.y, az1));
            double f0 = cov_eval<KIND>(cf, d2of(gx, gy, gz, b0.x, b0.y, bz0));
            double f1 = cov_eval<KIND>(cf, d2of(gx, gy, gz, b1.x, b1.y, bz1));
            const int c0 = 8 * (j + NW) + 2 * t;
            st2(dst, make_double2(gact ? e0 : 0.0, gact ? e1 : 0.0));
            st2(dst + 64 * NW, make_double2((gact && c0 < P) ? f0 : 0.0, (gact && c0 + 1 < P) ? f1 : 0.0));
            pc += 16 * NW;
            dst += 128 * NW;
          }
          for (; j < NS; j += NW) {
            const double2 a0 = ld2(&pc[0].x), a1 = ld2(&pc[1].x);
            double e0 = cov_eval<KIND>(cf, d2of(gx, gy, gz, a0.x, a0.y, DIM3 ? pc[0].z : 0.0));
            double e1 = cov_eval<KIND>(cf, d2of(gx, gy, gz, a1.x, a1.y, DIM3 ? pc[1].z : 0.0));
            const int c0 = 8 * j + 2 * t;
            e0 = (gact && c0 < P) ? e0 : 0.0;
            e1 = (gact && c0 + 1 < P) ? e1 : 0.0;
            st2(dst, make_double2(e0, e1));
            pc += 8 * NW;
            dst += 64 * NW;
          }
          if (h == 0) st2(rowp + 64 * NS, make_double2(0.0, 0.0));   // corner (c0 rows, c0 rows)
          rowp += (NS + 1) * 64;
        }
      };
      if (cf.kind == GSM_VARIO_SPHERICAL) fill(std::integral_constant<int, GSM_VARIO_SPHERICAL>{});
      else if (cf.kind == GSM_VARIO_GAUSSIAN) fill(std::integral_constant<int, GSM_VARIO_GAUSSIAN>{});
      else fill(std::integral_constant<int, GSM_VARIO_EXPONENTIAL>{});
      // rows shared by the patch: z (minus the SimpleKriging mean), drift monomials
      {
        double* rowp = blk + (size_t)tri(NS + 1) * 64 + fo;
#pragma unroll
        for (int zb = 0; zb < NZB; ++zb) {
          const int zr = 8 * zb + g;
          const double4* pc = prec + 2 * t;
          double* dst = rowp;
          for (int j = 0; j < NS; ++j) {
            if (!mine(j, h)) {
              pc += 8;
              dst += 64;
              continue;
            }
            const int c0 = 8 * j + 2 * t;
            double2 val = make_double2(0.0, 0.0);
            if (zr == 0) {
              val.x = c0 < P ? (NCON == 0 ? pc[0].w - m.mean : pc[0].w) : 0.0;
              val.y = c0 + 1 < P ? (NCON == 0 ? pc[1].w - m.mean : pc[1].w) : 0.0;
            } else if (zr <= NCON) {
              val.x = c0 < P ? driftv(zr - 1, (pc[0].x - cx) * isc, (pc[0].y - cy) * isc, (pc[0].z - cz) * isc) : 0.0;
              val.y = c0 + 1 < P ? driftv(zr - 1, (pc[1].x - cx) * isc, (pc[1].y - cy) * isc, (pc[1].z - cz) * isc) : 0.0;
            }
            st2(dst, val);
            pc += 8;
            dst += 64;
          }
          for (int j = NS; j <= NS + 1 + zb; ++j) {    // corners
            if (h == 0) st2(dst, make_double2(0.0, 0.0));
            dst += 64;
          }
          rowp += (NS + 2 + zb) * 64;
        }
      }
      cta_sync();
      // ---------------- eliminate the shared block columns ---------------------------------------------------
      const bool shared_bad = eliminate(S, NB);

      // ---------------- per-target part --------------------------------------------------------------------
      // Private slots of every member, by pool rank: lane L = 4 w + q holds slots 2q, 2q+1 of private block column
      // 0 (and of column 1 when npb = 2) of member w, each packed as  rowpart | colpart << 16  with
      //   rowpart(r) = tri(r >> 3) * 64 + (r & 7) * 8,   colpart(r) = (r >> 3) * 64 + (r & 7):
      // T(hi, lo) = blk[rowpart(hi) + colpart(lo)].  0xffffffff: empty slot (fewer than k neighbours).
      const int s8 = 8 * S;
      unsigned pk[2][2] = {{0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu}};
      {
        const int w_ = lane >> 2, q_ = lane & 3;
        const int nw = __shfl_sync(0xffffffffu, n_l, w_);
        if ((actmask >> w_) & 1u) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c < npb) {
              const int slot = s8 + 8 * c + 2 * q_;
              const uchar2 pr = *reinterpret_cast<const uchar2*>(&J.sid[w_][slot]);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int r = h ? pr.y : pr.x;
                if (slot + h < nw && r != 255)
                  pk[c][h] = (unsigned)(tri(r >> 3) * 64 + (r & 7) * 8) | ((unsigned)((r >> 3) * 64 + (r & 7)) << 16);
              }
            }
          }
        }
      }
      // scratch: the blocks of the eliminated columns are dead now; lane L knows the offset of dead block #L
      int dofs;
      {
        int s = lane, j = 0, cnt = NB;
        while (s >= cnt && j < S) {
          s -= cnt;
          ++j;
          --cnt;
        }
        dofs = (tri(j + s) + j) * 64;      // (only read for L < ndead)
      }
      __syncwarp();
      // offsets of the right-hand-side rows / corner entries of this lane (row a = 8 rb + g), job-invariant part
      const int trow = tri(NS) * 64;       // block row of the c0 rows
      int rbase[NRB], gbx[NGB], gby[NGB], gwx[NGB], gwy[NGB];
#pragma unroll
      for (int rb = 0; rb < NRB; ++rb) {
        const int a = 8 * rb + g;
        rbase[rb] = a == 0 ? trow : tri(NS + 1 + ((a - 1) >> 3)) * 64 + ((a - 1) & 7) * 8;   // + 8 w for a = 0
      }
#pragma unroll
      for (int ra = 0; ra < NRB; ++ra)
#pragma unroll
        for (int rb = 0; rb <= ra; ++rb) {
          const int a = 8 * ra + g, q = ra * (ra + 1) / 2 + rb;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int b_ = 8 * rb + 2 * t + h;
            int base = -1, wm = 0;          // entry = blk[base + wm * w]; -1: not needed
            if (a < NR && b_ <= a) {
              if (a == 0) base = trow + NS * 64, wm = 9;
              else if (b_ == 0) base = (tri(NS + 1 + ((a - 1) >> 3)) + NS) * 64 + ((a - 1) & 7) * 8, wm = 1;
              else base = (tri(NS + 1 + ((a - 1) >> 3)) + NS + 1 + ((b_ - 1) >> 3)) * 64 + ((a - 1) & 7) * 8 + ((b_ - 1) & 7);
            }
            if (h == 0) gbx[q] = base, gwx[q] = wm;
            else gby[q] = base, gwy[q] = wm;
          }
        }
      // element (row slot, column slot) of T from packed ranks; rows and columns of one block column ascend in rank
      auto Tat = [&](unsigned rowpk, unsigned colpk, bool row_hi) -> double {
        const unsigned hi = row_hi ? rowpk : colpk, lo = row_hi ? colpk : rowpk;
        return blk[(hi & 0xffffu) + (lo >> 16)];
      };
      const int bt = min(8, ndead / per_target);     // targets per batch
      for (int w0 = 0; w0 < 8; w0 += bt) {
        const int wend = min(8, w0 + bt);
        // -- first private diagonal block of every target of the batch
        for (int w = w0; w < wend; ++w) {
          if (!((actmask >> w) & 1u) || !mine(w, h)) continue;     // (NW = 2: members are dealt to the two warps by parity)
          const int so = __shfl_sync(0xffffffffu, dofs, (w - w0) * per_target);
          const unsigned ra_ = __shfl_sync(0xffffffffu, pk[0][0], 4 * w + (g >> 1)), rb_ = __shfl_sync(0xffffffffu, pk[0][1], 4 * w + (g >> 1));
          const unsigned rowpk = (g & 1) ? rb_ : ra_;
          const unsigned c0pk = __shfl_sync(0xffffffffu, pk[0][0], 4 * w + t), c1pk = __shfl_sync(0xffffffffu, pk[0][1], 4 * w + t);
          double2 val;
          val.x = ((int)(rowpk | c0pk) >= 0) ? Tat(rowpk, c0pk, g >= 2 * t) : (g == 2 * t ? v.sill : 0.0);
          val.y = ((int)(rowpk | c1pk) >= 0) ? Tat(rowpk, c1pk, g >= 2 * t + 1) : (g == 2 * t + 1 ? v.sill : 0.0);
          st2(blk + so + fo, val);
        }
        __syncwarp();
        bool pbad = false;
        {
          const int so = __shfl_sync(0xffffffffu, dofs, (max(lane, w0) - w0) * per_target & 31);
          if (lane >= w0 && lane < wend && act_l && mine(lane, h)) pbad = diag_block_inplace(blk + so, thr);
        }
        __syncwarp();
        if (npb >= 2) {
          // -- second private block column: L10, D1, and the right-hand sides carried past column 0
          for (int w = w0; w < wend; ++w) {
            if (!((actmask >> w) & 1u) || !mine(w, h)) continue;
            const int sb = (w - w0) * per_target;
            const int so = __shfl_sync(0xffffffffu, dofs, sb);
            const double2 W0 = ld2(blk + so + fo);
            const unsigned ra_ = __shfl_sync(0xffffffffu, pk[1][0], 4 * w + (g >> 1)), rb_ = __shfl_sync(0xffffffffu, pk[1][1], 4 * w + (g >> 1));
            const unsigned row1 = (g & 1) ? rb_ : ra_;
            const unsigned c00 = __shfl_sync(0xffffffffu, pk[0][0], 4 * w + t), c01 = __shfl_sync(0xffffffffu, pk[0][1], 4 * w + t);
            const unsigned c10 = __shfl_sync(0xffffffffu, pk[1][0], 4 * w + t), c11 = __shfl_sync(0xffffffffu, pk[1][1], 4 * w + t);
            double2 L10, D1;
            L10.x = ((int)(row1 | c00) >= 0) ? Tat(row1, c00, true) : 0.0;
            L10.y = ((int)(row1 | c01) >= 0) ? Tat(row1, c01, true) : 0.0;
            D1.x = ((int)(row1 | c10) >= 0) ? Tat(row1, c10, g >= 2 * t) : (g == 2 * t ? v.sill : 0.0);
            D1.y = ((int)(row1 | c11) >= 0) ? Tat(row1, c11, g >= 2 * t + 1) : (g == 2 * t + 1 ? v.sill : 0.0);
            panel(L10, W0);
            upd(D1, L10, L10);
            double2 R0[NRB], R1[NRB];
#pragma unroll
            for (int rb = 0; rb < NRB; ++rb) {
              const int a = 8 * rb + g;
              const int base = rbase[rb] + (a == 0 ? 8 * w : 0);
              const bool rl = a < NR;
              R0[rb] = make_double2((rl && (int)c00 >= 0) ? blk[base + (c00 >> 16)] : 0.0, (rl && (int)c01 >= 0) ? blk[base + (c01 >> 16)] : 0.0);
              R1[rb] = make_double2((rl && (int)c10 >= 0) ? blk[base + (c10 >> 16)] : 0.0, (rl && (int)c11 >= 0) ? blk[base + (c11 >> 16)] : 0.0);
              panel(R0[rb], W0);
              upd(R1[rb], R0[rb], L10);
            }
            // (every lane only reads / writes its own two elements of the slot: W0 may be overwritten now)
            st2(blk + so + fo, D1);
#pragma unroll
            for (int rb = 0; rb < NRB; ++rb) st2(blk + __shfl_sync(0xffffffffu, dofs, sb + 1 + rb) + fo, R1[rb]);
#pragma unroll
            for (int ra = 0; ra < NRB; ++ra)
#pragma unroll
              for (int rb = 0; rb <= ra; ++rb) {
                const int q = ra * (ra + 1) / 2 + rb;
                double2 G = make_double2(gbx[q] >= 0 ? blk[gbx[q] + gwx[q] * w] : 0.0, gby[q] >= 0 ? blk[gby[q] + gwy[q] * w] : 0.0);
                upd(G, R0[ra], R0[rb]);
                st2(blk + __shfl_sync(0xffffffffu, dofs, sb + 1 + NRB + q) + fo, G);
              }
          }
          __syncwarp();
          {
            const int so = __shfl_sync(0xffffffffu, dofs, (max(lane, w0) - w0) * per_target & 31);
            if (lane >= w0 && lane < wend && act_l && mine(lane, h)) pbad |= diag_block_inplace(blk + so, thr);
          }
          __syncwarp();
        }
        sing_l |= pbad;
        // -- last private block column: panel of the right-hand sides and the trailing Schur entries
        for (int w = w0; w < wend; ++w) {
          if (!((actmask >> w) & 1u) || !mine(w, h)) continue;
          const int sb = (w - w0) * per_target;
          const double2 Wl = ld2(blk + __shfl_sync(0xffffffffu, dofs, sb) + fo);
          double2 R[NRB];
          if (npb <= 1) {
            const unsigned c0pk = __shfl_sync(0xffffffffu, pk[0][0], 4 * w + t), c1pk = __shfl_sync(0xffffffffu, pk[0][1], 4 * w + t);
#pragma unroll
            for (int rb = 0; rb < NRB; ++rb) {
              const int a = 8 * rb + g;
              const int base = rbase[rb] + (a == 0 ? 8 * w : 0);
              const bool rl = a < NR;
              R[rb] = make_double2((rl && (int)c0pk >= 0) ? blk[base + (c0pk >> 16)] : 0.0, (rl && (int)c1pk >= 0) ? blk[base + (c1pk >> 16)] : 0.0);
            }
          } else {
#pragma unroll
            for (int rb = 0; rb < NRB; ++rb) R[rb] = ld2(blk + __shfl_sync(0xffffffffu, dofs, sb + 1 + rb) + fo);
          }
#pragma unroll
          for (int rb = 0; rb < NRB; ++rb) panel(R[rb], Wl);
#pragma unroll
          for (int ra = 0; ra < NRB; ++ra)
#pragma unroll
            for (int rb = 0; rb <= ra; ++rb) {
              const int q = ra * (ra + 1) / 2 + rb;
              const int a = 8 * ra + g, b0 = 8 * rb + 2 * t, b1 = b0 + 1;
              double2 G;
              if (npb <= 1)
                G = make_double2(gbx[q] >= 0 ? blk[gbx[q] + gwx[q] * w] : 0.0, gby[q] >= 0 ? blk[gby[q] + gwy[q] * w] : 0.0);
              else
                G = ld2(blk + __shfl_sync(0xffffffffu, dofs, sb + 1 + NRB + q) + fo);
              upd(G, R[ra], R[rb]);
              if (a < NR && b0 < NR) gout[w * GS + a * NR + b0] = G.x;
              if (a < NR && b1 < NR) gout[w * GS + a * NR + b1] = G.y;
            }
        }
        cta_sync();     // (the next batch reuses the scratch slots, possibly from the other warp)
      }
      sing_l |= shared_bad;
    } else if (actmask != 0u) {
      // ---------------- generic path: every active target on its own k x k system ----------------------------
      const int NBt = nbc + NRB;
      for (int w = 0; w < 8; ++w) {
        if (!((actmask >> w) & 1u)) continue;
        const int nw = __shfl_sync(0xffffffffu, n_l, w);
        const long long ttw = __shfl_sync(0xffffffffu, tt_l, w);
        const double wx = __shfl_sync(0xffffffffu, tx, w), wy = __shfl_sync(0xffffffffu, ty, w),
                     wz = __shfl_sync(0xffffffffu, tz, w);
        double4 rec = make_double4(0.0, 0.0, 0.0, 0.0);
        if (P >= 0) {
          if (lane < nw) rec = prec[J.sid[w][lane]];
        } else {
          int q = 0x7fffffff;
          if (lane < nw) q = p.pos[ttw * (long long)p.k + lane];
          q = gsm_warp_sort_int(q, lane);        // canonical order: ascending record position
          if (lane < nw) rec = p.b.rec[q];
        }
        cta_sync();       // (the previous target's blocks, srec among them, are no longer read)
        if (h == 0) srec[lane] = rec;
        cta_sync();
        double cx = 0.0, cy = 0.0, cz = 0.0, isc = 1.0;
        if (NCON > 0 && p.centered) {
          double dmax = lane < nw ? d2of(rec.x, rec.y, rec.z, wx, wy, wz) : 0.0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, gsm_shfl_xor(dmax, o));
          isc = dmax > 0.0 ? rsqrt(dmax) : 1.0;
          cx = wx, cy = wy, cz = wz;
        }
        {
          int i = 0, j = 0;
          double4 rr = srec[g];
          const int nblocks = NBt * (NBt + 1) / 2;
          for (int bi = 0; bi < nblocks; ++bi) {
            const int c0 = 8 * j + 2 * t, c1 = c0 + 1;
            double2 val = make_double2(0.0, 0.0);
            if (j < nbc) {
              const double4 a0 = srec[c0], a1 = srec[c1];
              if (i < nbc) {
                const int r = 8 * i + g;
                const double e0 = cov_fast(cf, d2of(rr.x, rr.y, rr.z, a0.x, a0.y, a0.z));
                const double e1 = cov_fast(cf, d2of(rr.x, rr.y, rr.z, a1.x, a1.y, a1.z));
                val.x = r == c0 ? v.sill : ((r < nw && c0 < nw) ? e0 : 0.0);
                val.y = r == c1 ? v.sill : ((r < nw && c1 < nw) ? e1 : 0.0);
              } else {
                const int a = 8 * (i - nbc) + g;
                if (a == 0) {
                  val.x = c0 < nw ? cov_fast(cf, d2of(wx, wy, wz, a0.x, a0.y, a0.z)) : 0.0;
                  val.y = c1 < nw ? cov_fast(cf, d2of(wx, wy, wz, a1.x, a1.y, a1.z)) : 0.0;
                } else if (a == 1) {
                  val.x = c0 < nw ? (NCON == 0 ? a0.w - m.mean : a0.w) : 0.0;
                  val.y = c1 < nw ? (NCON == 0 ? a1.w - m.mean : a1.w) : 0.0;
                } else if (a < NR) {
                  val.x = c0 < nw ? driftv(a - 2, (a0.x - cx) * isc, (a0.y - cy) * isc, (a0.z - cz) * isc) : 0.0;
                  val.y = c1 < nw ? driftv(a - 2, (a1.x - cx) * isc, (a1.y - cy) * isc, (a1.z - cz) * isc) : 0.0;
                }
              }
            }
            if (mine(bi, h)) st2(blk + bi * 64 + fo, val);
            if (++j > i) {
              ++i;
              j = 0;
              if (i < nbc) rr = srec[8 * i + g];
            }
          }
        }
        cta_sync();
        const bool bad = eliminate(nbc, NBt);
#pragma unroll
        for (int ra = 0; ra < NRB; ++ra)
#pragma unroll
          for (int rb = 0; rb <= ra; ++rb) {
            const double2 G = ld2(blk + bidx(nbc + ra, nbc + rb) * 64 + fo);
            const int a = 8 * ra + g, b0 = 8 * rb + 2 * t, b1 = b0 + 1;
            if (a < NR && b0 < NR) gout[w * GS + a * NR + b0] = G.x;
            if (a < NR && b1 < NR) gout[w * GS + a * NR + b1] = G.y;
          }
        if (lane == w) {
          sing_l = bad;
          ecx = cx, ecy = cy, ecz = cz, eisc = isc;
        }
        __syncwarp();
      }
    }
    __syncwarp();

    // ---------------- epilogue: thread per target (mean, variance, status) -------------------------------------
    if (lane < 8 && (fl_l & 1) && mine(lane, h)) {
      if (!act_l) {
        p.mean[tt_l] = qnan;
        if (p.var) p.var[tt_l] = qnan;
        if (p.status) p.status[tt_l] = GSM_ST_MISSING;
      } else {
        const double* Gw = gout + lane * GS;
        auto Gm = [&](int a, int b_) -> double { return Gw[a * NR + b_]; };   // a >= b_: -(B' C^-1 B)[a][b_]
        bool singular = sing_l;
        const double uu = -Gm(0, 0), wu = -Gm(1, 0);
        double mean, var;
        if (NCON == 0) {
          mean = m.mean + wu;
          var = v.sill - uu;
        } else {
          double Gs[NCON > 0 ? NCON * NCON : 1], gq[NCON > 0 ? NCON : 1], hw[NCON > 0 ? NCON : 1];
#pragma unroll
          for (int q = 0; q < NCON; ++q) {
            gq[q] = -Gm(2 + q, 0);
            hw[q] = -Gm(2 + q, 1);
#pragma unroll
            for (int q2 = 0; q2 <= q; ++q2) Gs[q * NCON + q2] = -Gm(2 + q, 2 + q2);
          }
          {
            // f0: the drift monomials at the target, in the frame the sample monomials were evaluated in
            const double fx = (tx - ecx) * eisc, fy = (ty - ecy) * eisc, fz = (tz - ecz) * eisc;
#pragma unroll
            for (int q = 0; q < NCON; ++q) gq[q] -= driftv(q, fx, fy, fz);
          }
          double gd0[NCON > 0 ? NCON : 1];   // original diagonal: the singularity test is relative to it
#pragma unroll
          for (int q = 0; q < NCON; ++q) gd0[q] = Gs[q * NCON + q];
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
          var = v.sill - uu + gnu;
        }
        var += 1e-10;  // krig.jl:278-279
        p.mean[tt_l] = singular ? qnan : mean;
        if (p.var) p.var[tt_l] = singular ? qnan : var;
        if (p.status) p.status[tt_l] = singular ? GSM_ST_SINGULAR : GSM_ST_OK;
      }
    }
    cta_sync();     // the next job overwrites the pool and the block store
  }
}

template <int NCON, bool DIM3, int NW = 1>
int launch_patch(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, const TargetsDev& tg, const JobIdx* jobs,
                 long long ngroups, int k, int nbc, int centered, const int* pos, double* mean, double* var,
                 unsigned char* status) {
  constexpr int NR = 2 + NCON, NRB = NR > 8 ? 2 : 1;
  PatchParams p{};
  p.b = ctx->bins;
  p.v = v;
  p.m = m;
  p.tg = tg;
  p.cf = gsm_cov::make_cov(v);
  p.jobs = jobs;
  p.ngroups = ngroups;
  p.k = k;
  p.nbc = nbc;
  p.centered = centered;
  p.drift_const = (NCON == 1 && m.exps[0][0] == 0 && m.exps[0][1] == 0 && m.exps[0][2] == 0) ? 1 : 0;
  // block rows of the store: the generic path needs nbc + NRB; the patch form S + ceil((P - 8 S) / 8) + 1 + NZB.
  // 2-D patches pool ~38 samples (5 block rows), 3-D ~48 (6): one spare row each (developer option "patch_nbmax").
  int nbmax = ctx->opt_patch_nbmax > 0 ? ctx->opt_patch_nbmax : (ctx->dim >= 3 ? 9 : 8) + (NRB - 1);
  nbmax = std::max(nbmax, nbc + NRB);
  nbmax = std::min(nbmax, 8 + 1 + 2);
  p.nbmax = nbmax;
  p.nblk = nbmax * (nbmax + 1) / 2;
  p.warp_bytes = (int)(GSM_JOB_PMAX * sizeof(double4) + (size_t)p.nblk * 64 * sizeof(double) + 8 * NR * NR * sizeof(double));
  p.pos = pos;
  p.mean = mean;
  p.var = var;
  p.status = status;
  auto kern = local_krige_patch_kernel<NCON, DIM3, NW>;
  GSM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.warp_bytes));
  int per_sm = 0;
  GSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * NW, (size_t)p.warp_bytes));
  per_sm = std::max(per_sm, 1);
  const int blocks = (int)std::min<long long>(ngroups, 148LL * per_sm * 4);
  kern<<<blocks, 32 * NW, p.warp_bytes, ctx->stream>>>(p);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

}  // namespace gsm_patch
