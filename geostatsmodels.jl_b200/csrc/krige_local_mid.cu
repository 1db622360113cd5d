// Local Kriging for 32 < k <= 64 neighbours (config C5 sweeps k up to 64): one WARP per system, the augmented
// matrix [C ; B'] in shared memory as 8x8 blocks in the mma.m8n8k4 fragment layout, blocked left-looking
// Cholesky on the FP64 DMMA path.  Same mathematics as the register-resident kernels (block elimination of the
// reference's saddle system, krig.jl:58-75 / 316-339 / 245-293): the right-hand sides (c0, z, drift monomials) are
// carried as extra block rows, so after the last block column the trailing block(s) are -B'C^-1B.
// Differences from krige_local_shr.cuh / krige_local_pipe.cuh, which keep a 32 x 32 system in registers:
//   * 36 + 8 blocks do not fit a warp's registers, so blocks live in shared memory (22.5 KB per system) and only
//     the accumulator of the block being formed is in registers; the block loops are rolled (compact code);
//   * TWO warps per system (the 22.5 KB per system bound the occupancy, not the warps) that meet at one named
//     barrier per column: warp A runs the serial chain one column ahead (block row kb+1 -> next diagonal block ->
//     its inverse factor, Gaussian elimination on [A | I] with shuffles inside the warp), warp B does every other
//     block row of the column and the Schur blocks; the assembly is split by parity;
//   * covariances are evaluated directly (no pooling across neighbouring targets).
// One warp per system took 34 ms per 2^20 systems at k = 64, two with look-ahead take 26 ms (the chain of eight
// diagonal-block inversions per system is what is left).
// It replaces the CTA-per-system SIMT kernel (krige_local_big.cu, kept as GSM_LOCAL_IMPL=big) as the k > 32 path.
#include <algorithm>

#include <type_traits>

#include "gsm_internal.cuh"
#include "krige_cov_fast.cuh"

namespace {

constexpr int MW = 4;                // systems per CTA (two warps each)
constexpr int MKMAX = GSM_MAX_NEIGHBORS;   // 64
constexpr int MNBC = MKMAX / 8;      // 8 block columns at most

__device__ __forceinline__ void dmma_m(double2& d, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d.x), "+d"(d.y)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ double rsqrt_m(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-(d * y), y, 1.0);
  const double p = e * fma(e, 0.375, 0.5);
  return fma(y, p, y);
}

// sill - gamma(h) from a squared distance with the variogram kind as a compile-time constant (krige_cov_fast.cuh holds
// the run-time form; same arithmetic)
template <int KIND>
__device__ __forceinline__ double cov_kind_mid(const gsm_cov::CovFast& c, double d2) {
  const double d2c = __hiloint2double(max(__double2hiint(d2), 0x03e00000), __double2loint(d2));
  double v;
  if (KIND == GSM_VARIO_SPHERICAL) {
    const double t2 = d2 * c.inv_r2;
    const double y = gsm_cov::fast_rsqrt(d2c);
    const double t = (d2 * c.inv_r) * y;
    v = fma(t, fma(c.k2, t2, c.k1), c.c1);
    v = (d2 < c.range2) ? v : c.cfar;
  } else if (KIND == GSM_VARIO_GAUSSIAN) {
    const double t2 = d2 * c.inv_r2;
    v = c.sill - (c.cs * (1.0 - exp(-3.0 * t2)) + (c.sill - c.c1));
  } else {
    const double y = gsm_cov::fast_rsqrt(d2c);
    const double t = (d2 * c.inv_r) * y;
    v = c.sill - (c.cs * (1.0 - exp(-3.0 * t)) + (c.sill - c.c1));
  }
  return (d2 > 0.0) ? v : c.sill;
}

// Warp-cooperative inverse Cholesky factor of an 8x8 SPD block in fragment layout (lane (g,t) owns A[g][2t],
// A[g][2t+1]): Gaussian elimination without pivoting on [A | I] gives the unit-lower inverse M; scaling row g by
// 1/sqrt(pivot g) turns it into W = L^-1 with A = L L'.
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
    const double rs = rsqrt_m(d);
    if (g == j) myrs = rs;
    const double f = (g > j) ? -((arj * rs) * rs) : 0.0;
    a.x = fma(f, pj0, a.x);
    a.y = fma(f, pj1, a.y);
    m.x = fma(f, mj0, m.x);
    m.y = fma(f, mj1, m.y);
  }
  return make_double2(m.x * myrs, m.y * myrs);
}

// blocks of one system: lower triangle of the covariance part, then NRB right-hand-side block rows
__host__ __device__ constexpr int mid_blocks(int nrb) { return MNBC * (MNBC + 1) / 2 + nrb * MNBC; }

template <int NRB>
struct alignas(16) MidWarp {
  double blk[mid_blocks(NRB)][64];
  double px[MKMAX], py[MKMAX], pz[MKMAX], pv[MKMAX];
  double G[NRB * (NRB + 1) / 2][64];
  double Wb[2][64];                   // inverse diagonal factors, published one column ahead
  int sp[MKMAX];
  int badflag;
};

template <int NRB>
__global__ void __launch_bounds__(MW * 64)
local_krige_mid_kernel(const __grid_constant__ BinsDev b, const __grid_constant__ VarioDev v,
                       const __grid_constant__ ModelDev m, const __grid_constant__ gsm_cov::CovFast cf, int ncon,
                       int centered, const __grid_constant__ TargetsDev tg, const __grid_constant__ BlockDev bk, int k,
                       int minn,
                       const int* __restrict__ pos,
                       const int* __restrict__ cnt, double* __restrict__ out_mean, double* __restrict__ out_var,
                       unsigned char* __restrict__ out_status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int sys = warp >> 1, h = warp & 1;             // system of this CTA, half of the pair
  MidWarp<NRB>& S = reinterpret_cast<MidWarp<NRB>*>(smem_raw)[sys];
  auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + sys) : "memory"); };
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const int nr = 2 + ncon;
  constexpr int NGB = NRB * (NRB + 1) / 2;

  auto lowidx = [](int I, int J) { return I * (I + 1) / 2 + J; };   // J <= I < nbc
  auto frag = [&](const double* src) -> double2 { return *reinterpret_cast<const double2*>(&src[g * 8 + 2 * t]); };
  auto put = [&](double* dst, const double2& val) { *reinterpret_cast<double2*>(&dst[g * 8 + 2 * t]) = val; };
  auto upd = [&](double2& A, const double2& X, const double2& Y) {   // A -= X Y'
    dmma_m(A, -X.x, Y.x);
    dmma_m(A, -X.y, Y.y);
  };
  auto panel = [&](double2& X, const double2& W) {                   // X <- X W'
    double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
    dmma_m(a0, X.x, W.x);
    dmma_m(a1, X.y, W.y);
    X = make_double2(a0.x + a1.x, a0.y + a1.y);
  };

  const long long nsys = (long long)gridDim.x * MW;
  for (long long tt = blockIdx.x * (long long)MW + sys; tt < tg.count; tt += nsys) {
    int n = cnt[tt];
    if (n < minn || n <= 0) {
      if (lane == 0 && h == 0) {
        out_mean[tt] = qnan;
        if (out_var) out_var[tt] = qnan;
        if (out_status) out_status[tt] = GSM_ST_MISSING;
      }
      continue;
    }
    n = min(n, min(k, MKMAX));
    const int nbc = (n + 7) >> 3;                       // block columns of this system
    double tx, ty, tz;
    gsm_target_coords(tg, tg.first + tt, tx, ty, tz);

    // ---- canonical neighbour order: ascending record position (rank by counting; warp h takes entries 32h..) ----
    pair_sync();                                        // the previous system of this pair is finished
    const int e0 = 32 * h + lane;
    const int p0 = e0 < n ? pos[tt * (long long)k + e0] : 0x7fffffff;
    S.sp[e0] = p0;
    pair_sync();
    int r0 = 0;
    for (int j = 0; j < n; ++j) r0 += S.sp[j] < p0 ? 1 : 0;
    if (e0 < n) {
      const double4 r = b.rec[p0];
      S.px[r0] = r.x, S.py[r0] = r.y, S.pz[r0] = r.z, S.pv[r0] = r.w;
    }
    pair_sync();
    // drift scaling: centre at the target, scale by the neighbourhood radius (lower-set exponent tables)
    double ox = 0.0, oy = 0.0, oz = 0.0, isc = 1.0;
    if (centered) {
      double dmax = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double dx = S.px[i] - tx, dy = S.py[i] - ty, dz = S.pz[i] - tz;
        dmax = fmax(dmax, dx * dx + dy * dy + dz * dz);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, gsm_shfl_xor(dmax, o));
      ox = tx, oy = ty, oz = tz;
      isc = dmax > 0.0 ? rsqrt(dmax) : 1.0;
    }

    // ---- assemble the blocks (fragment layout); empty slots are identity rows / columns ----
    int bcount = 0;
    const bool fullsys = n == 8 * nbc && !v.general;
    if (fullsys) {
      // every slot live, one of the three isotropic families (the common case): row coordinates hoisted per block row,
      // the two columns of a lane loaded as pairs, no padding tests, the variogram kind compiled in
      auto fast = [&](auto kind_tag) __attribute__((always_inline)) {
        constexpr int KIND = decltype(kind_tag)::value;
        for (int I = 0; I < nbc; ++I) {
          const double rx = S.px[8 * I + g], ry = S.py[8 * I + g], rz = S.pz[8 * I + g];
          for (int J = 0; J <= I; ++J) {
            if (((bcount++) & 1) != h) continue;
            const double2 cx = *reinterpret_cast<const double2*>(&S.px[8 * J + 2 * t]);
            const double2 cy = *reinterpret_cast<const double2*>(&S.py[8 * J + 2 * t]);
            const double2 cz = *reinterpret_cast<const double2*>(&S.pz[8 * J + 2 * t]);
            const double dx0 = rx - cx.x, dy0 = ry - cy.x, dz0 = rz - cz.x, dx1 = rx - cx.y, dy1 = ry - cy.y, dz1 = rz - cz.y;
            double e0 = cov_kind_mid<KIND>(cf, fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0)));
            double e1 = cov_kind_mid<KIND>(cf, fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1)));
            if (I == J) {
              e0 = g == 2 * t ? v.sill : e0;
              e1 = g == 2 * t + 1 ? v.sill : e1;
            }
            put(S.blk[lowidx(I, J)], make_double2(e0, e1));
          }
        }
      };
      if (cf.kind == GSM_VARIO_SPHERICAL) fast(std::integral_constant<int, GSM_VARIO_SPHERICAL>{});
      else if (cf.kind == GSM_VARIO_GAUSSIAN) fast(std::integral_constant<int, GSM_VARIO_GAUSSIAN>{});
      else fast(std::integral_constant<int, GSM_VARIO_EXPONENTIAL>{});
    }
    for (int I = 0; I < (fullsys ? 0 : nbc); ++I)
      for (int J = 0; J <= I; ++J) {
        if (((bcount++) & 1) != h) continue;
        const int row = 8 * I + g, c0 = 8 * J + 2 * t;
        double e[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int col = c0 + s;
          double val = (row == col) ? v.sill : 0.0;   // padding: sill * identity
          if (row < n && col < n) {
            val = v.sill;
            if (row != col) {
              const double dx = S.px[row] - S.px[col], dy = S.py[row] - S.py[col], dz = S.pz[row] - S.pz[col];
              val = v.general ? gsm_cov_general(v, dx, dy, dz) : gsm_cov::cov_fast(cf, fma(dz, dz, fma(dy, dy, dx * dx)));
            }
          }
          e[s] = val;
        }
        put(S.blk[lowidx(I, J)], make_double2(e[0], e[1]));
      }
    const int rbase = MNBC * (MNBC + 1) / 2;            // right-hand-side blocks: rbase + R * MNBC + J
    for (int R = 0; R < NRB; ++R)
      for (int J = 0; J < nbc; ++J) {
        if (((bcount++) & 1) != h) continue;
        const int q = 8 * R + g, c0 = 8 * J + 2 * t;
        double e[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int c = c0 + s;
          double val = 0.0;
          if (q < nr && c < n) {
            if (q == 0) {
              const double dx = S.px[c] - tx, dy = S.py[c] - ty, dz = S.pz[c] - tz;
              if (bk.nq > 0) val = gsm_cov_block(v, bk, dx, dy, dz);   // change of support: mean over the block's points
              else val = v.general ? gsm_cov_general(v, dx, dy, dz) : gsm_cov::cov_fast(cf, fma(dz, dz, fma(dy, dy, dx * dx)));
            } else if (q == 1) {
              val = (m.kind == GSM_KRIG_SIMPLE) ? S.pv[c] - m.mean : S.pv[c];
            } else {
              val = gsm_drift(m, q - 2, (S.px[c] - ox) * isc, (S.py[c] - oy) * isc, (S.pz[c] - oz) * isc);
            }
          }
          e[s] = val;
        }
        put(S.blk[rbase + R * MNBC + J], make_double2(e[0], e[1]));
      }
    pair_sync();

    // ---- left-looking blocked Cholesky with LOOK-AHEAD, right-hand-side rows included; blocks become L ----
    // The chain "block row kb+1 of column kb -> diagonal block kb+1 -> its inverse factor" is the serial part:
    // warp 0 of the pair (A) runs it one column ahead and publishes W in shared memory, while warp 1 (B) does every
    // other block row of column kb (the right-hand sides included) and accumulates the Schur blocks.
    bool bad = false;
    double2 Gacc[NGB];
#pragma unroll
    for (int q = 0; q < NGB; ++q) Gacc[q] = make_double2(0.0, 0.0);
    if (h == 0) put(S.Wb[0], inv_chol_warp(frag(S.blk[lowidx(0, 0)]), g, t, GSM_PIVOT_RTOL * v.sill, bad));
    pair_sync();
#pragma unroll 1
    for (int kb = 0; kb < nbc; ++kb) {
      const double2 W = frag(S.Wb[kb & 1]);
      if (h == 0) {
        if (kb + 1 < nbc) {
          // row kb+1 of this column, then the next diagonal block and its inverse factor
          double* bi = S.blk[lowidx(kb + 1, kb)];
          double2 C = frag(bi);
          double2 D = frag(S.blk[lowidx(kb + 1, kb + 1)]);
#pragma unroll 1
          for (int j = 0; j < kb; ++j) {
            const double2 X = frag(S.blk[lowidx(kb + 1, j)]);
            upd(C, X, frag(S.blk[lowidx(kb, j)]));
            upd(D, X, X);
          }
          panel(C, W);
          put(bi, C);
          upd(D, C, C);
          put(S.Wb[(kb + 1) & 1], inv_chol_warp(D, g, t, GSM_PIVOT_RTOL * v.sill, bad));
        }
      } else {
#pragma unroll 1
        for (int i = kb + 2; i < nbc + NRB; ++i) {
          const bool rhs = i >= nbc;
          double* bi = rhs ? S.blk[rbase + (i - nbc) * MNBC + kb] : S.blk[lowidx(i, kb)];
          double2 C = frag(bi);
#pragma unroll 1
          for (int j = 0; j < kb; ++j) {
            const double2 X = frag(rhs ? S.blk[rbase + (i - nbc) * MNBC + j] : S.blk[lowidx(i, j)]);
            upd(C, X, frag(S.blk[lowidx(kb, j)]));
          }
          panel(C, W);
          put(bi, C);
        }
        if (kb + 1 >= nbc) {
          // (the first right-hand-side row is block row kb+1 of the last column: nobody looks ahead there)
#pragma unroll 1
          for (int i = kb + 1; i < min(kb + 2, nbc + NRB); ++i) {
            double* bi = S.blk[rbase + (i - nbc) * MNBC + kb];
            double2 C = frag(bi);
#pragma unroll 1
            for (int j = 0; j < kb; ++j) upd(C, frag(S.blk[rbase + (i - nbc) * MNBC + j]), frag(S.blk[lowidx(kb, j)]));
            panel(C, W);
            put(bi, C);
          }
        }
        // Schur blocks: G[r][q] -= L_Rr,kb L_Rq,kb'
#pragma unroll
        for (int r = 0; r < NRB; ++r)
#pragma unroll
          for (int q = 0; q <= r; ++q) {
            const double2 X = frag(S.blk[rbase + r * MNBC + kb]), Y = frag(S.blk[rbase + q * MNBC + kb]);
            upd(Gacc[r * (r + 1) / 2 + q], X, Y);
          }
      }
      pair_sync();
    }
    if (h == 0) {
      if (lane == 0) S.badflag = bad ? 1 : 0;
    } else {
#pragma unroll
      for (int q = 0; q < NGB; ++q) put(S.G[q], Gacc[q]);
    }
    pair_sync();
    if (h != 0) continue;                               // the epilogue belongs to warp 0 of the pair
    const unsigned anybad = S.badflag;
    __syncwarp();

    // ---- epilogue (one lane): -G[a][b], a >= b, stored as 8x8 blocks (0,0), (1,0), (1,1) ----
    if (lane == 0) {
      bool singular = anybad != 0;
      auto Gm = [&](int a, int bb) -> double {
        const int br = a >> 3, bc = bb >> 3;
        return -S.G[br * (br + 1) / 2 + bc][(a & 7) * 8 + (bb & 7)];
      };
      const double uu = Gm(0, 0), wu = Gm(1, 0);
      const double c0 = bk.c0;   // covzero (krig.jl:295-301): the sill for a point target
      double mean, var;
      if (ncon == 0) {
        mean = m.mean + wu;
        var = c0 - uu;
      } else {
        double Gs[GSM_MAX_DRIFT * GSM_MAX_DRIFT], gq[GSM_MAX_DRIFT], hw[GSM_MAX_DRIFT], y[GSM_MAX_DRIFT], hy[GSM_MAX_DRIFT];
        for (int p = 0; p < ncon; ++p) {
          const double f0 = (m.kind == GSM_KRIG_ORDINARY) ? 1.0
                                                           : gsm_drift(m, p, (tx - ox) * isc, (ty - oy) * isc, (tz - oz) * isc);
          gq[p] = Gm(2 + p, 0) - f0;
          hw[p] = Gm(2 + p, 1);
          for (int q = 0; q <= p; ++q) Gs[p * ncon + q] = Gm(2 + p, 2 + q);
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
      out_mean[tt] = singular ? qnan : mean;
      if (out_var) out_var[tt] = singular ? qnan : var;
      if (out_status) out_status[tt] = singular ? GSM_ST_SINGULAR : GSM_ST_OK;
    }
  }
}

template <int NRB>
int launch_mid(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered, const TargetsDev& tg, int k,
               int minn, const int* pos, const int* cnt, double* mean, double* var, unsigned char* status) {
  const size_t smem = sizeof(MidWarp<NRB>) * MW;   // per system; two warps work on each
  auto kern = local_krige_mid_kernel<NRB>;
  GSM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = (int)std::max<size_t>(1, (227 * 1024) / (smem + 1024));
  const long long want = (tg.count + MW - 1) / MW;
  const int blocks = (int)std::min<long long>(want, 148LL * per_sm);
  kern<<<blocks, MW * 64, smem, ctx->stream>>>(ctx->bins, v, m, gsm_cov::make_cov(v), ncon, centered, tg, ctx->block_cur, k, minn, pos, cnt,
                                               mean, var, status);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

}  // namespace

int gsm_launch_local_krige_mid(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered,
                               const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                               double* d_mean, double* d_var, unsigned char* d_status) {
  if (k > MKMAX || k < 1 || ncon > GSM_MAX_DRIFT) return GSM_ERR_UNSUPPORTED;
  if (tg.count <= 0) return GSM_OK;
  if (2 + ncon > 8)
    return launch_mid<2>(ctx, v, m, ncon, centered, tg, k, minneighbors, d_pos, d_cnt, d_mean, d_var, d_status);
  return launch_mid<1>(ctx, v, m, ncon, centered, tg, k, minneighbors, d_pos, d_cnt, d_mean, d_var, d_status);
}
