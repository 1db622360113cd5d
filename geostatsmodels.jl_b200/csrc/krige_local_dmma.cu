// Batched local Kriging, tensor-path version (k <= 32, at most 8 right-hand-side rows: SK, OK and UK
// with up to 6 drift monomials).  Same mathematics as krige_local.cu (see its header for the block
// elimination of the reference's saddle-point system, krig.jl:58-75 / 316-339 / 245-293) but mapped on
// the FP64 DMMA pipe:
//
//   * one warp per system; the augmented matrix  [C ; B']  (B = right-hand sides c0, z, drift columns,
//     stored as 8 extra ROWS) lives in registers as 8x8 blocks in the mma.m8n8k4 accumulator layout
//     (lane (g,t) holds (row g, cols 2t, 2t+1) of every block);
//   * blocked right-looking Cholesky over 4 block columns.  With W = inv(L_kk) the panel solve
//     L_ik = A_ik W' and the trailing update A_ij -= L_ik L_jk' are both plain DMMA products whose
//     operands ARE accumulator-layout registers (the k index is summed in the order 0,2,4,6 then
//     1,3,5,7 -- any order is a valid contraction), so no shuffles or shared memory are needed;
//   * the RHS rows ride through the same updates: after the last block column the trailing 8x8 block
//     holds -B'C^-1B, i.e. every inner product the prediction needs (u'u, w'u, V'u, V'w, V'V);
//   * the only serial work -- Cholesky + inverse of the four 8x8 diagonal blocks -- is batched across
//     the systems of a CTA: each system warp posts its block to shared memory and a dedicated warp
//     factors one block per lane (thread-per-block, registers, fully unrolled).
#include <algorithm>

#include "gsm_internal.cuh"

#if 0   // superseded by the trace in krige_local_pipe.cuh
// Developer-only phase trace (make trace): clock64 stamps of CTA 0 for a few group iterations.
__device__ long long* g_trace = nullptr;
extern "C" __attribute__((visibility("default"))) int gsm_debug_set_trace(long long* p) {
  return cudaMemcpyToSymbol(g_trace, &p, sizeof(p)) == cudaSuccess ? 0 : 1;
}
#define TR(slot)                                                                                      \
  do {                                                                                                \
    if (g_trace && blockIdx.x == 0 && lane == 0 && grp >= 20LL * gridDim.x && grp < 24LL * gridDim.x) \
      g_trace[(((grp / gridDim.x) - 20) * 9 + warp) * 40 + (slot)] = clock64();                       \
  } while (0)
#else
#define TR(slot) do {} while (0)
#endif

namespace {

constexpr int SW = 8;                      // system warps per CTA
constexpr int STHREADS = (SW + 1) * 32;    // + 1 diagonal-block warp
constexpr int EXS = 66;                    // padded stride (doubles) of one exchanged 8x8 block

__device__ __forceinline__ void dmma(double2& d, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d.x), "+d"(d.y)
               : "d"(a), "d"(b));
}

// Named barriers (ids 1, 2) for the post / wait hand-off between system warps and the diagonal warp:
// the poster arrives without waiting, the consumer syncs.
constexpr int BAR_READY = 1, BAR_DONE = 2;
__device__ __forceinline__ void bar_arrive(int id) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(STHREADS) : "memory");
}
__device__ __forceinline__ void bar_sync(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(STHREADS) : "memory");
}
__device__ __forceinline__ void bar_sync_sys(int id) {   // the SW system warps only
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(SW * 32) : "memory");
}

// 1/sqrt(d), d > 0: hardware seed (rel. error ~2^-22) + one third-order step -> full double precision
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-(d * y), y, 1.0);
  const double p = e * fma(e, 0.375, 0.5);
  return fma(y, p, y);
}

// covariance sill - gamma(h) from a squared distance, fast path (no IEEE sqrt / division)
struct CovFast {
  int kind;
  double sill, c1, k1, k2, inv_r, inv_r2, cs, range2, cfar;
};
__device__ __forceinline__ CovFast make_cov(const VarioDev& v) {
  CovFast c;
  c.kind = v.kind;
  c.sill = v.sill;
  c.c1 = v.sill - v.nadd;       // value as h -> 0+
  c.cs = v.cs;
  c.k1 = -1.5 * v.cs;
  c.k2 = 0.5 * v.cs;
  c.inv_r = v.inv_range;
  c.inv_r2 = v.inv_range * v.inv_range;
  c.range2 = v.range * v.range;
  c.cfar = v.sill - (v.cs + v.nadd);   // h >= range (spherical)
  return c;
}
__device__ __forceinline__ double cov_fast(const CovFast& c, double d2) {
  if (c.kind == GSM_VARIO_SPHERICAL) {
    const double t2 = d2 * c.inv_r2;
    const double y = fast_rsqrt(fmax(d2, 1e-300));
    const double t = (d2 * c.inv_r) * y;               // h / r
    double v = fma(t, fma(c.k2, t2, c.k1), c.c1);      // c1 - cs (1.5 t - 0.5 t^3)
    v = (d2 < c.range2) ? v : c.cfar;
    return (d2 > 0.0) ? v : c.sill;
  } else if (c.kind == GSM_VARIO_GAUSSIAN) {
    const double t2 = d2 * c.inv_r2;
    const double v = c.sill - (c.cs * (1.0 - exp(-3.0 * t2)) + (c.sill - c.c1));
    return (d2 > 0.0) ? v : c.sill;
  } else {
    const double y = fast_rsqrt(fmax(d2, 1e-300));
    const double t = (d2 * c.inv_r) * y;
    const double v = c.sill - (c.cs * (1.0 - exp(-3.0 * t)) + (c.sill - c.c1));
    return (d2 > 0.0) ? v : c.sill;
  }
}

// ---- thread-per-block Cholesky + inverse of an 8x8 SPD block ---------------------------------
// in : lower triangle read from blk[r*8 + c];  out: inv(L) written to out[r*8 + c] (lower; the strictly
// upper part of `out` is zeroed once at kernel start and never touched).  Returns true if a pivot was
// not positive.
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
    a[LIDX(j, j)] = rs;  // keep 1 / L_jj
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[LIDX(i, j)] *= rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i)
#pragma unroll
      for (int k = j + 1; k <= i; ++k) a[LIDX(i, k)] = fma(-a[LIDX(i, j)], a[LIDX(k, j)], a[LIDX(i, k)]);
  }
  // X = inv(L), column by column: X_jj = 1/L_jj, X_ij = -(sum_{k=j}^{i-1} L_ik X_kj) / L_ii
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
    }
  }
  return bad;
}

struct SysSmem {
  double px[32], py[32], pz[32];
  double rhs[8][32];
};

// Covariance pool shared by the SW systems of a group: consecutive targets share most neighbours, so
// the pairwise covariances are evaluated once per distinct sample pair of the group and every system
// gathers its 32x32 block from the pooled matrix (bit-identical values: cov is a function of the pair).
constexpr int PMAX = 64;          // pool capacity (distinct samples per group); larger unions fall back
constexpr int PLD = PMAX + 1;     // leading dimension of the pooled matrix (odd: spreads banks)
constexpr int HSIZE = 512;        // open-addressing table for de-duplicating neighbour positions
constexpr int BAR_POOL = 3;       // named barrier among the system warps only
constexpr int SYS_THREADS = SW * 32;

struct KernelSmem {
  double exin[SW * EXS];
  double exout[SW * EXS];
  SysSmem sys[SW];
  double PC[PMAX * PLD];
  double ppx[PMAX], ppy[PMAX], ppz[PMAX];
  int hkey[HSIZE];
  int hid[HSIZE];
  int ids[SW][32];
  int flag[SW];
  int pcount;
};

template <int NCON, bool DIM3>
__global__ void __launch_bounds__(STHREADS, 2)
local_krige_dmma_kernel(BinsDev b, VarioDev v, ModelDev m, TargetsDev tg, int k, int minn, int centered,
                        const int* __restrict__ pos, const int* __restrict__ cnt, double* __restrict__ out_mean,
                        double* __restrict__ out_var, unsigned char* __restrict__ out_status) {
  constexpr int NR = 2 + NCON;
  static_assert(NR <= 8, "at most 8 right-hand-side rows");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KernelSmem& S = *reinterpret_cast<KernelSmem*>(smem_raw);
  double* s_exin = S.exin;
  double* s_exout = S.exout;
  SysSmem* s_sys = S.sys;
  int* s_flag = S.flag;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const long long ngroups = (tg.count + SW - 1) / SW;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  for (int e = threadIdx.x; e < SW * EXS; e += STHREADS) s_exout[e] = 0.0;
  for (int e = threadIdx.x; e < HSIZE; e += STHREADS) S.hkey[e] = -1;
  for (int e = threadIdx.x; e < PMAX; e += STHREADS) S.ppx[e] = S.ppy[e] = S.ppz[e] = 0.0;
  if (threadIdx.x == 0) S.pcount = 0;
  __syncthreads();

  if (warp == SW) {
    // ---------------- diagonal-block warp: one block per lane ----------------
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb) {
        TR(kb * 3 + 0);
        bar_sync(BAR_READY);
        TR(kb * 3 + 1);
        if (lane < SW) {
          if (diag_block(&s_exin[lane * EXS], &s_exout[lane * EXS], GSM_PIVOT_RTOL * v.sill)) s_flag[lane] = 1;
        }
        bar_arrive(BAR_DONE);
        TR(kb * 3 + 2);
      }
    }
    return;
  }

  // -------------------- system warps --------------------
  const CovFast cf = make_cov(v);
  SysSmem& sm = s_sys[warp];
  for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    TR(0);
    const long long tt = grp * SW + warp;
    const bool valid = tt < tg.count;
    int n = valid ? cnt[tt] : 0;
    const bool missing = (n < minn) || (n <= 0);
    if (missing) n = 0;
    double tx = 0, ty = 0, tz = 0;
    if (valid) gsm_target_coords(tg, tg.first + tt, tx, ty, tz);

    // ---- gather neighbours, right-hand-side rows -------------------------------------------
    const bool live = lane < n;
    double4 r = make_double4(0.0, 0.0, 0.0, 0.0);
    int mypos = -1;
    if (lane < k && valid) mypos = pos[tt * (long long)k + lane];
    {
      const int key = gsm_warp_sort_int(mypos < 0 ? 0x7fffffff : mypos, lane);   // canonical order
      mypos = key == 0x7fffffff ? -1 : key;
    }
    if (live) r = b.rec[mypos];
    sm.px[lane] = r.x;
    sm.py[lane] = r.y;
    if (DIM3) sm.pz[lane] = r.z;
    {
      const double dx = r.x - tx, dy = r.y - ty, dz = DIM3 ? r.z - tz : 0.0;
      const double d2t = DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx);
      sm.rhs[0][lane] = live ? cov_fast(cf, d2t) : 0.0;
      sm.rhs[1][lane] = live ? (NCON == 0 ? r.w - m.mean : r.w) : 0.0;
      if (NCON > 0) {
        double ox = 0.0, oy = 0.0, oz = 0.0, isc = 1.0;
        if (centered) {
          double dmax = live ? d2t : 0.0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, gsm_shfl_xor(dmax, o));
          isc = dmax > 0.0 ? rsqrt(dmax) : 1.0;
          ox = tx;
          oy = ty;
          oz = tz;
        }
        const double sx = (r.x - ox) * isc, sy = (r.y - oy) * isc, sz = (r.z - oz) * isc;
#pragma unroll
        for (int q = 0; q < NCON; ++q) sm.rhs[2 + q][lane] = live ? gsm_drift(m, q, sx, sy, sz) : 0.0;
      }
    }
    if (lane == 0) s_flag[warp] = 0;
    TR(1);

    // ---- pool the distinct neighbours of the group and their pairwise covariances ------------
    int myslot = 0;
    bool owner = false;
    if (live) {
      myslot = (int)(((unsigned)mypos * 0x9E3779B1u) >> 23);
      for (;;) {
        const int old = atomicCAS(&S.hkey[myslot], -1, mypos);
        if (old == -1) {
          owner = true;
          break;
        }
        if (old == mypos) break;
        myslot = (myslot + 1) & (HSIZE - 1);
      }
    }
    bar_sync_sys(BAR_POOL);
    TR(2);
    if (owner) {
      const int id = atomicAdd(&S.pcount, 1);
      S.hid[myslot] = id;
      if (id < PMAX) {
        S.ppx[id] = r.x;
        S.ppy[id] = r.y;
        S.ppz[id] = r.z;
      }
    }
    bar_sync_sys(BAR_POOL);
    TR(3);
    const int P = S.pcount;
    const bool pooled = P <= PMAX;
    S.ids[warp][lane] = live ? S.hid[myslot] : -1;
    if (pooled) {
      // lower triangle (with diagonal) of the P x P pooled matrix, folded into a (Pe/2) x (Pe+1) rectangle
      // (Pe = P rounded up to even; a dummy last entry is computed and never read)
      const int Pe = P + (P & 1), Wd = Pe + 1, npairs = (Pe >> 1) * Wd;
      const float invW = 1.0f / (float)Wd;
      for (int e = warp * 32 + lane; e < npairs; e += SYS_THREADS) {
        const int rr = (int)(((float)e + 0.5f) * invW);
        const int cc = e - rr * Wd;
        const int a = (cc <= rr) ? rr : Pe - 1 - rr;
        const int bcol = (cc <= rr) ? cc : cc - rr - 1;
        double val = v.sill;
        if (a != bcol) {
          const double dx = S.ppx[a] - S.ppx[bcol], dy = S.ppy[a] - S.ppy[bcol];
          const double dz = DIM3 ? S.ppz[a] - S.ppz[bcol] : 0.0;
          val = cov_fast(cf, DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx));
        }
        S.PC[a * PLD + bcol] = val;
        S.PC[bcol * PLD + a] = val;
      }
    }
    TR(4);
    bar_sync_sys(BAR_POOL);
    TR(5);
    // every lookup of this group is done: release the table for the next group
    if (live) S.hkey[myslot] = -1;
    if (warp == 0 && lane == 0) S.pcount = 0;

    // ---- assemble the augmented matrix in accumulator layout ---------------------------------
    // Block (0,0) first: it is posted to the diagonal warp at once so its factorisation overlaps
    // the assembly of the other 14 blocks.
    double2 A[5][5];
    auto assemble = [&](int I, int J) {
      if (pooled) {
        const int row = 8 * I + g, col = 8 * J + 2 * t;
        const int ir = S.ids[warp][row];
        const int2 ic = *reinterpret_cast<const int2*>(&S.ids[warp][col]);
        double c0 = (row == col) ? 1.0 : 0.0, c1 = (row == col + 1) ? 1.0 : 0.0;
        if (ir >= 0 && ic.x >= 0) c0 = S.PC[ir * PLD + ic.x];
        if (ir >= 0 && ic.y >= 0) c1 = S.PC[ir * PLD + ic.y];
        return make_double2(c0, c1);
      }
      const int row = 8 * I + g;
      const double xr = sm.px[row], yr = sm.py[row], zr = DIM3 ? sm.pz[row] : 0.0;
      const bool rowlive = row < n;
      const int col = 8 * J + 2 * t;
      const double2 xc = *reinterpret_cast<const double2*>(&sm.px[col]);
      const double2 yc = *reinterpret_cast<const double2*>(&sm.py[col]);
      double2 zc = make_double2(0.0, 0.0);
      if (DIM3) zc = *reinterpret_cast<const double2*>(&sm.pz[col]);
      double dx = xr - xc.x, dy = yr - yc.x, dz = zr - zc.x;
      double c0 = cov_fast(cf, DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx));
      dx = xr - xc.y, dy = yr - yc.y, dz = zr - zc.y;
      double c1 = cov_fast(cf, DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx));
      c0 = (row == col) ? (rowlive ? v.sill : 1.0) : ((rowlive && col < n) ? c0 : 0.0);
      c1 = (row == col + 1) ? (rowlive ? v.sill : 1.0) : ((rowlive && col + 1 < n) ? c1 : 0.0);
      return make_double2(c0, c1);
    };
    A[0][0] = assemble(0, 0);
    *reinterpret_cast<double2*>(&s_exin[warp * EXS + g * 8 + 2 * t]) = A[0][0];
    bar_arrive(BAR_READY);
    TR(6);
#pragma unroll
    for (int I = 1; I < 4; ++I)
#pragma unroll
      for (int J = 0; J <= I; ++J) A[I][J] = assemble(I, J);
#pragma unroll
    for (int J = 0; J < 4; ++J) {
      double2 val = make_double2(0.0, 0.0);
      if (g < NR) val = *reinterpret_cast<const double2*>(&sm.rhs[g][8 * J + 2 * t]);
      A[4][J] = val;
    }
    A[4][4] = make_double2(0.0, 0.0);
    TR(7);

    // ---- blocked Cholesky with the RHS rows riding along ------------------------------------
    // Look-ahead: at step kb the next diagonal block (kb+1,kb+1) is finished and posted first, the
    // rest of the trailing update runs while the diagonal warp factors it.
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      TR(8 + kb * 4);
      bar_sync(BAR_DONE);
      TR(9 + kb * 4);
      const double2 W = *reinterpret_cast<const double2*>(&s_exout[warp * EXS + g * 8 + 2 * t]);
      // Each product is two k-chunks (accumulator slots x and y).  All x-chunks are issued before the
      // y-chunks so consecutive DMMAs never wait on each other's accumulator.
      {   // first the row that produces the next diagonal block
        double2 acc = make_double2(0.0, 0.0), acc2 = make_double2(0.0, 0.0);
        dmma(acc, A[kb + 1][kb].x, W.x);
        dmma(acc2, A[kb + 1][kb].y, W.y);
        A[kb + 1][kb] = make_double2(acc.x + acc2.x, acc.y + acc2.y);   // L_{kb+1,kb} = A_{kb+1,kb} W'
        double2 u = make_double2(0.0, 0.0);
        dmma(A[kb + 1][kb + 1], -A[kb + 1][kb].x, A[kb + 1][kb].x);
        dmma(u, -A[kb + 1][kb].y, A[kb + 1][kb].y);
        A[kb + 1][kb + 1].x += u.x;
        A[kb + 1][kb + 1].y += u.y;
      }
      if (kb < 3) {
        *reinterpret_cast<double2*>(&s_exin[warp * EXS + g * 8 + 2 * t]) = A[kb + 1][kb + 1];
        bar_arrive(BAR_READY);
      }
      TR(10 + kb * 4);
      {   // remaining panel rows: L_ik = A_ik W'
        double2 acc[5];
#pragma unroll
        for (int i = kb + 2; i < 5; ++i) {
          acc[i] = make_double2(0.0, 0.0);
          dmma(acc[i], A[i][kb].x, W.x);
        }
#pragma unroll
        for (int i = kb + 2; i < 5; ++i) {
          dmma(acc[i], A[i][kb].y, W.y);
          A[i][kb] = acc[i];
        }
      }
      // remaining trailing update: A_ij -= L_ik L_jk'
#pragma unroll
      for (int i = kb + 2; i < 5; ++i)
#pragma unroll
        for (int j = kb + 1; j <= i; ++j) dmma(A[i][j], -A[i][kb].x, A[j][kb].x);
#pragma unroll
      for (int i = kb + 2; i < 5; ++i)
#pragma unroll
        for (int j = kb + 1; j <= i; ++j) dmma(A[i][j], -A[i][kb].y, A[j][kb].y);
      TR(11 + kb * 4);
    }

    // ---- epilogue: the trailing block is -B'C^-1B ---------------------------------------------
    *reinterpret_cast<double2*>(&s_exin[warp * EXS + g * 8 + 2 * t]) = A[4][4];
    __syncwarp();
    if (lane == 0 && valid) {
      const double* Gm = &s_exin[warp * EXS];   // Gm[a*8+b] = -G[a][b]
      bool singular = s_flag[warp] != 0;
      const double uu = -Gm[0], wu = -Gm[8];
      const double c0 = v.sill;
      double mean, var;
      if (NCON == 0) {
        mean = m.mean + wu;
        var = c0 - uu;
      } else {
        double G[NCON > 0 ? NCON * NCON : 1], gq[NCON > 0 ? NCON : 1], hw[NCON > 0 ? NCON : 1];
        double ox = 0.0, oy = 0.0, oz = 0.0, isc = 1.0;
        // f0: drift at the target in the same (possibly centred) coordinates as the samples
#pragma unroll
        for (int p = 0; p < NCON; ++p) {
          gq[p] = -Gm[(2 + p) * 8 + 0];
          hw[p] = -Gm[(2 + p) * 8 + 1];
#pragma unroll
          for (int q = 0; q <= p; ++q) G[p * NCON + q] = -Gm[(2 + p) * 8 + 2 + q];
        }
        if (centered) {
          // centred coordinates: the target is the origin -> only the constant monomial is 1
#pragma unroll
          for (int p = 0; p < NCON; ++p)
            gq[p] -= (m.exps[p][0] == 0 && m.exps[p][1] == 0 && m.exps[p][2] == 0) ? 1.0 : 0.0;
        } else {
#pragma unroll
          for (int p = 0; p < NCON; ++p) gq[p] -= gsm_drift(m, p, (tx - ox) * isc, (ty - oy) * isc, (tz - oz) * isc);
        }
#pragma unroll
        double gd0[NCON > 0 ? NCON : 1];   // original diagonal: the singularity test is relative to it
        for (int p0_ = 0; p0_ < NCON; ++p0_) gd0[p0_] = G[p0_ * NCON + p0_];
        for (int j = 0; j < NCON; ++j) {
          double d = G[j * NCON + j];
          if (!(d > GSM_SCHUR_RTOL * gd0[j])) {
            singular = true;
            d = 1.0;
          }
          const double rs = rsqrt(d);
          G[j * NCON + j] = rs;
#pragma unroll
          for (int i = j + 1; i < NCON; ++i) G[i * NCON + j] *= rs;
#pragma unroll
          for (int i = j + 1; i < NCON; ++i)
#pragma unroll
            for (int q = j + 1; q <= i; ++q) G[i * NCON + q] = fma(-G[i * NCON + j], G[q * NCON + j], G[i * NCON + q]);
        }
        double y[NCON > 0 ? NCON : 1], hy[NCON > 0 ? NCON : 1];
#pragma unroll
        for (int i = 0; i < NCON; ++i) {
          double s1 = gq[i], s2 = hw[i];
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
      if (missing) {
        out_mean[tt] = qnan;
        if (out_var) out_var[tt] = qnan;
        if (out_status) out_status[tt] = GSM_ST_MISSING;
      } else {
        out_mean[tt] = singular ? qnan : mean;
        if (out_var) out_var[tt] = singular ? qnan : var;
        if (out_status) out_status[tt] = singular ? GSM_ST_SINGULAR : GSM_ST_OK;
      }
    }
    __syncwarp();
    TR(24);
  }
}

template <int NCON>
int launch_dmma(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, const TargetsDev& tg, int k, int minn,
                int centered, const int* pos, const int* cnt, double* mean, double* var, unsigned char* status) {
  const long long ngroups = (tg.count + SW - 1) / SW;
  const int blocks = (int)std::min<long long>(ngroups, 148LL * 2);
  const size_t smem = sizeof(KernelSmem);
  if (ctx->dim >= 3) {
    GSM_CUDA(ctx, cudaFuncSetAttribute(local_krige_dmma_kernel<NCON, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    local_krige_dmma_kernel<NCON, true><<<blocks, STHREADS, smem, ctx->stream>>>(ctx->bins, v, m, tg, k, minn, centered,
                                                                                pos, cnt, mean, var, status);
  } else {
    GSM_CUDA(ctx, cudaFuncSetAttribute(local_krige_dmma_kernel<NCON, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    local_krige_dmma_kernel<NCON, false><<<blocks, STHREADS, smem, ctx->stream>>>(ctx->bins, v, m, tg, k, minn, centered,
                                                                                 pos, cnt, mean, var, status);
  }
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

}  // namespace

// returns GSM_ERR_UNSUPPORTED (without touching the error message) when the shape is not covered
int gsm_launch_local_krige_dmma(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered,
                                const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                                double* d_mean, double* d_var, unsigned char* d_status) {
  if (k > 32 || ncon > 6) return GSM_ERR_UNSUPPORTED;
#define GSM_LKD_CASE(N) \
  case N:               \
    return launch_dmma<N>(ctx, v, m, tg, k, minneighbors, centered, d_pos, d_cnt, d_mean, d_var, d_status);
  switch (ncon) {
    GSM_LKD_CASE(0)
    GSM_LKD_CASE(1)
    GSM_LKD_CASE(2)
    GSM_LKD_CASE(3)
    GSM_LKD_CASE(4)
    GSM_LKD_CASE(5)
    GSM_LKD_CASE(6)
  }
#undef GSM_LKD_CASE
  return GSM_ERR_UNSUPPORTED;
}
