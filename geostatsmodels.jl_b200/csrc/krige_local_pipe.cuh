// Batched local Kriging on the FP64 DMMA path, software-pipelined version (k <= 32, <= 8 RHS rows).
// Mathematics and register layout as in krige_local_dmma.cu (reference: fit krig.jl:58-75, weights
// krig.jl:316-339, krigmean krig.jl:245-258, predictvar krig.jl:264-293 per target of models.jl:162-184).
// What changes is the schedule, driven by the phase trace of the first DMMA kernel:
//
//   * LEFT-LOOKING blocked Cholesky with LAZY assembly: a block is gathered from the pooled covariance
//     matrix right before its first use, so only block (0,0) precedes the first hand-off to the
//     diagonal warp and at most 11 blocks (22 doubles) are live instead of 15.
//   * The whole pre-phase of the NEXT group (neighbour loads, covariance to the target, right-hand-side
//     rows, hash de-duplication, pooled pair covariances) is issued inside the four wait slots of the
//     CURRENT group's factorisation, so its global-load and barrier latencies are hidden.
//   * Pool of up to 72 distinct samples per group (full 72 x 73 matrix, 42 KB), single-buffered:
//     every gather of the current group precedes the slot in which the next pool is evaluated.
#include <algorithm>

#include "gsm_internal.cuh"

#if defined(GSM_TRACE) && GSM_PIPE_DIM3 == 0 && GSM_PIPE_NBC == 4   // developer phase trace: 2-D, k <= 32 unit only
__device__ long long* g_trace2 = nullptr;
extern "C" __attribute__((visibility("default"))) int gsm_debug_set_trace2(long long* p) {
  return cudaMemcpyToSymbol(g_trace2, &p, sizeof(p)) == cudaSuccess ? 0 : 1;
}
#define TR(slot)                                                                                       \
  do {                                                                                                 \
    if (g_trace2 && blockIdx.x == 0 && lane == 0 && grp >= 20LL * gridDim.x && grp < 24LL * gridDim.x) \
      g_trace2[(((grp / gridDim.x) - 20) * 9 + warp) * 40 + (slot)] = clock64();                       \
  } while (0)
#else
#define TR(slot) do {} while (0)
#endif

namespace {

constexpr int SW = 8;                      // system warps per CTA
constexpr int NTH = (SW + 1) * 32;         // + 1 diagonal-block warp
constexpr int EXS = 66;                    // padded stride (doubles) of one exchanged 8x8 block
constexpr int PMAX = 72;                   // pool capacity (distinct samples per group)
constexpr int PLD = PMAX + 1;              // leading dimension of the pooled matrix (odd: spreads banks)
constexpr int BAR_EPI = 4;                 // system warps -> diagonal warp: Schur block ready for the epilogue
constexpr int HSIZE = 512;                 // open-addressing table for de-duplicating positions
constexpr int BAR_READY = 1, BAR_DONE = 2, BAR_POOL = 3;

__device__ __forceinline__ void dmma(double2& d, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d.x), "+d"(d.y)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void bar_arrive(int id) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(NTH) : "memory");
}
__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NTH) : "memory"); }
__device__ __forceinline__ void bar_sync_sys_all() { asm volatile("bar.sync %0, %1;" ::"n"(5), "n"(NTH) : "memory"); }
__device__ __forceinline__ void bar_sync_sys() {
  asm volatile("bar.sync %0, %1;" ::"n"(BAR_POOL), "n"(SW * 32) : "memory");
}

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
__device__ __forceinline__ CovFast make_cov(const VarioDev& v) {
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
__device__ __forceinline__ double cov_fast(const CovFast& c, double d2) {
  if (c.kind == GSM_VARIO_SPHERICAL) {
    const double t2 = d2 * c.inv_r2;
    const double y = fast_rsqrt(__hiloint2double(max(__double2hiint(d2), 0x03e00000), __double2loint(d2)));
    const double t = (d2 * c.inv_r) * y;
    double v = fma(t, fma(c.k2, t2, c.k1), c.c1);
    v = (d2 < c.range2) ? v : c.cfar;
    return (d2 > 0.0) ? v : c.sill;
  } else if (c.kind == GSM_VARIO_GAUSSIAN) {
    const double t2 = d2 * c.inv_r2;
    const double v = c.sill - (c.cs * (1.0 - exp(-3.0 * t2)) + (c.sill - c.c1));
    return (d2 > 0.0) ? v : c.sill;
  } else {
    const double y = fast_rsqrt(__hiloint2double(max(__double2hiint(d2), 0x03e00000), __double2loint(d2)));
    const double t = (d2 * c.inv_r) * y;
    const double v = c.sill - (c.cs * (1.0 - exp(-3.0 * t)) + (c.sill - c.c1));
    return (d2 > 0.0) ? v : c.sill;
  }
}

// thread-per-block Cholesky + inverse of an 8x8 SPD block (see krige_local_dmma.cu)
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
    }
  }
  return bad;
}

template <int NR>
struct alignas(16) PipeSmem {
  static constexpr int NRB = NR > 8 ? 2 : 1;            // right-hand-side block rows
  static constexpr int NGB = NRB * (NRB + 1) / 2;       // Schur blocks handed to the epilogue
  double exin[SW * EXS];
  double exout[SW * EXS];
  double px[2][SW][32], py[2][SW][32], pz[2][SW][32];   // neighbour coordinates (direct-assembly fallback)
  double rhs[2][SW][NR][32];                            // right-hand-side rows: c0, z, drifts
  double PC[PMAX * PLD];                                // pooled covariances, full symmetric
  double Gx[SW * NGB * EXS];                            // trailing Schur blocks handed to the epilogue
  double tgt[2][SW][4];                                 // target coordinates of the groups in flight
  long long e_tt[SW];                                   // epilogue hand-off: target index, flags
  int e_flags[SW];
  double ppx[PMAX], ppy[PMAX], ppz[PMAX];
  int hkey[HSIZE];
  int hid[HSIZE];
  int ids[2][SW][32];
  int flag[SW];
  int pctr;
};

// NCON: constraints (drift monomials); NBC: 8-wide block columns of the covariance block (k <= 8 NBC)
template <int NCON, bool DIM3, int NBC>
__global__ void __launch_bounds__(NTH, (2 + NCON > 8) ? 1 : 2)
local_krige_pipe_kernel(BinsDev b, VarioDev v, ModelDev m, TargetsDev tg, int k, int minn, int centered,
                        const int* __restrict__ pos, const int* __restrict__ cnt, double* __restrict__ out_mean,
                        double* __restrict__ out_var, unsigned char* __restrict__ out_status) {
  constexpr int NR = 2 + NCON;
  constexpr int NRB = NR > 8 ? 2 : 1;       // right-hand-side block rows
  constexpr int NGB = NRB * (NRB + 1) / 2;
  constexpr int NROWS = NBC + NRB;          // block rows of the augmented matrix
  constexpr int KMAX = 8 * NBC;             // neighbour slots
  static_assert(NR <= 16 && NBC >= 1 && NBC <= 4, "unsupported shape");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PipeSmem<NR>& S = *reinterpret_cast<PipeSmem<NR>*>(smem_raw);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const long long ngroups = (tg.count + SW - 1) / SW;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  for (int e = threadIdx.x; e < SW * EXS; e += NTH) S.exout[e] = 0.0;
  for (int e = threadIdx.x; e < HSIZE; e += NTH) S.hkey[e] = -1;
  for (int e = threadIdx.x; e < PMAX; e += NTH) S.ppx[e] = S.ppy[e] = S.ppz[e] = 0.0;
  if (threadIdx.x == 0) S.pctr = 0;
  __syncthreads();

  if (warp == SW) {
    // ---------------- diagonal-block warp: one lane per system ----------------
    // Besides the four diagonal blocks it does the per-system scalar work thread-per-system: target
    // coordinates of the next group and the Schur-complement epilogue (mean, variance, status).
    auto coords_for = [&](long long grp, int bf) {
      if (lane < SW) {
        const long long tt = grp * SW + lane;
        double x = 0.0, y = 0.0, z = 0.0;
        if (grp < ngroups && tt < tg.count) gsm_target_coords(tg, tg.first + tt, x, y, z);
        S.tgt[bf][lane][0] = x;
        S.tgt[bf][lane][1] = y;
        S.tgt[bf][lane][2] = z;
      }
    };
    int dbuf = 0;
    coords_for(blockIdx.x, 0);
    __threadfence_block();
    bar_sync_sys_all();   // prologue hand-off of the first group's coordinates
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, dbuf ^= 1) {
      coords_for(grp + gridDim.x, dbuf ^ 1);
      if (lane < SW) S.flag[lane] = 0;
#pragma unroll 1
      for (int kb = 0; kb < NBC; ++kb) {
        TR(kb * 3 + 0);
        bar_sync(BAR_READY);
        TR(kb * 3 + 1);
        if (lane < SW) {
          if (diag_block(&S.exin[lane * EXS], &S.exout[lane * EXS], GSM_PIVOT_RTOL * v.sill)) S.flag[lane] = 1;
        }
        bar_arrive(BAR_DONE);
        TR(kb * 3 + 2);
      }
      bar_sync(BAR_EPI);
      if (lane < SW && (S.e_flags[lane] & 1)) {
        // -G[a][b], a >= b, stored as 8x8 blocks (0,0), (1,0), (1,1)
        auto Gm = [&](int a, int bb) -> double {
          const int br = a >> 3, bc = bb >> 3;
          return S.Gx[(lane * NGB + br * (br + 1) / 2 + bc) * EXS + (a & 7) * 8 + (bb & 7)];
        };
        const long long tt = S.e_tt[lane];
        bool singular = S.flag[lane] != 0;
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
            const double tx = S.tgt[dbuf][lane][0], ty = S.tgt[dbuf][lane][1], tz = S.tgt[dbuf][lane][2];
#pragma unroll
            for (int p = 0; p < NCON; ++p) gq[p] -= gsm_drift(m, p, tx, ty, tz);
          }
          double gd0[NCON > 0 ? NCON : 1];   // original diagonal: the singularity test is relative to it
#pragma unroll
          for (int p0_ = 0; p0_ < NCON; ++p0_) gd0[p0_] = Gs[p0_ * NCON + p0_];
#pragma unroll
          for (int j = 0; j < NCON; ++j) {
            double d = Gs[j * NCON + j];
            if (!(d > GSM_SCHUR_RTOL * gd0[j])) {
              singular = true;
              d = 1.0;
            }
            const double rs = rsqrt(d);
            Gs[j * NCON + j] = rs;
#pragma unroll
            for (int i = j + 1; i < NCON; ++i) Gs[i * NCON + j] *= rs;
#pragma unroll
            for (int i = j + 1; i < NCON; ++i)
#pragma unroll
              for (int q = j + 1; q <= i; ++q)
                Gs[i * NCON + q] = fma(-Gs[i * NCON + j], Gs[q * NCON + j], Gs[i * NCON + q]);
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
        if (S.e_flags[lane] & 2) {
          out_mean[tt] = qnan;
          if (out_var) out_var[tt] = qnan;
          if (out_status) out_status[tt] = GSM_ST_MISSING;
        } else {
          out_mean[tt] = singular ? qnan : mean;
          if (out_var) out_var[tt] = singular ? qnan : var;
          if (out_status) out_status[tt] = singular ? GSM_ST_SINGULAR : GSM_ST_OK;
        }
      }
    }
    return;
  }

  // ============================== system warps ==============================
  const CovFast cf = make_cov(v);

  // ---- state of a group: "c" = current (being factored), "x" = next (being prepared) ----------
  long long c_tt = 0;
  int c_n = 0, c_P = 0;
  bool c_valid = false, c_missing = true;
  double c_tx = 0, c_ty = 0, c_tz = 0;

  long long x_tt = 0;
  int x_n = 0, x_pos = -1, x_slot = 0, x_P = 0;
  bool x_valid = false, x_missing = true, x_owner = false;
  double x_tx = 0, x_ty = 0, x_tz = 0;
  double4 x_r = make_double4(0.0, 0.0, 0.0, 0.0);

  // P0: issue the index loads of group `grp`
  auto P0 = [&](long long grp) {
    x_tt = grp * SW + warp;
    x_valid = grp < ngroups && x_tt < tg.count;
    x_pos = -1;
    x_n = 0;
    if (x_valid) {
      x_n = cnt[x_tt];
      if (lane < k && lane < KMAX) x_pos = pos[x_tt * (long long)k + lane];
    }
  };
  // canonical neighbour order: ascending record position, unused slots (-1) last
  auto Psort = [&]() {
    const int key = gsm_warp_sort_int(x_pos < 0 ? 0x7fffffff : x_pos, lane);
    x_pos = key == 0x7fffffff ? -1 : key;
  };
  // P1: neighbour records
  auto P1 = [&]() {
    Psort();
    x_missing = (x_n < minn) || (x_n <= 0);
    if (x_missing) x_n = 0;
    x_r = make_double4(0.0, 0.0, 0.0, 0.0);
    if (lane < x_n) x_r = b.rec[x_pos];
  };
  // P2+P3: coordinates / right-hand-side rows into buffer `bf`, hash insert, pool ids for owners
  auto P23 = [&](int bf) {
    const bool live = lane < x_n;
    // (computed here, not read from the diagonal warp's S.tgt: with one or two block columns this phase can run
    //  before the diagonal warp has finished the previous group's epilogue and written the next coordinates)
    x_tx = x_ty = x_tz = 0.0;
    if (x_valid) gsm_target_coords(tg, tg.first + x_tt, x_tx, x_ty, x_tz);
    S.px[bf][warp][lane] = x_r.x;
    S.py[bf][warp][lane] = x_r.y;
    if (DIM3) S.pz[bf][warp][lane] = x_r.z;
    const double dx = x_r.x - x_tx, dy = x_r.y - x_ty, dz = DIM3 ? x_r.z - x_tz : 0.0;
    const double d2t = DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx);
    S.rhs[bf][warp][0][lane] = live ? cov_fast(cf, d2t) : 0.0;
    S.rhs[bf][warp][1][lane] = live ? (NCON == 0 ? x_r.w - m.mean : x_r.w) : 0.0;
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
      const double sx = (x_r.x - ox) * isc, sy = (x_r.y - oy) * isc, sz = (x_r.z - oz) * isc;
#pragma unroll
      for (int q = 0; q < NCON; ++q) S.rhs[bf][warp][2 + q][lane] = live ? gsm_drift(m, q, sx, sy, sz) : 0.0;
    }
    x_owner = false;
    x_slot = 0;
    if (live) {
      x_slot = (int)(((unsigned)x_pos * 0x9E3779B1u) >> 23);
      for (;;) {
        const int old = atomicCAS(&S.hkey[x_slot], -1, x_pos);
        if (old == -1) {
          x_owner = true;
          break;
        }
        if (old == x_pos) break;
        x_slot = (x_slot + 1) & (HSIZE - 1);
      }
      if (x_owner) {
        const int id = atomicAdd(&S.pctr, 1);
        S.hid[x_slot] = id;
        if (id < PMAX) {
          S.ppx[id] = x_r.x;
          S.ppy[id] = x_r.y;
          S.ppz[id] = x_r.z;
        }
      }
    }
  };
  // P4 (after the pool barrier): pool ids of every neighbour, pooled pair covariances
  auto P4 = [&](int bf) {
    x_P = S.pctr;
    S.ids[bf][warp][lane] = (lane < x_n) ? S.hid[x_slot] : -1;
    if (x_P <= PMAX) {
      const int Pe = x_P + (x_P & 1), Wd = Pe + 1, npairs = (Pe >> 1) * Wd;
      const float invW = 1.0f / (float)Wd;
#pragma unroll 2
      for (int e = warp * 32 + lane; e < npairs; e += SW * 32) {
        const int rr = (int)(((float)e + 0.5f) * invW);
        const int cc = e - rr * Wd;
        const int a = (cc <= rr) ? rr : Pe - 1 - rr;
        const int bc = (cc <= rr) ? cc : cc - rr - 1;
        double val = v.sill;
        if (a != bc) {
          const double dx = S.ppx[a] - S.ppx[bc], dy = S.ppy[a] - S.ppy[bc];
          const double dz = DIM3 ? S.ppz[a] - S.ppz[bc] : 0.0;
          val = cov_fast(cf, DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx));
        }
        if (a < PMAX) {
          S.PC[a * PLD + bc] = val;
          S.PC[bc * PLD + a] = val;
        }
      }
    }
  };
  // after the second pool barrier: release the hash table
  auto P5 = [&]() {
    if (lane < x_n) S.hkey[x_slot] = -1;
    if (warp == 0 && lane == 0) S.pctr = 0;
  };

  // ---- prologue: prepare the first group --------------------------------------------------
  bar_sync_sys_all();   // first group's target coordinates (diagonal warp)
  P0(blockIdx.x);
  P1();
  P23(0);
  bar_sync_sys();
  P4(0);
  bar_sync_sys();
  P5();
  if (NBC <= 2) bar_sync_sys();

  int buf = 0;
  for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, buf ^= 1) {
    // next -> current
    c_tt = x_tt;
    c_n = x_n;
    c_P = x_P;
    c_valid = x_valid;
    c_missing = x_missing;
    c_tx = x_tx;
    c_ty = x_ty;
    c_tz = x_tz;
    const bool pooled = c_P <= PMAX;
    const int n = c_n;

    int irow[NBC], rbase[NBC];
    int2 icol[NBC];
#pragma unroll
    for (int I = 0; I < NBC; ++I) {
      irow[I] = S.ids[buf][warp][8 * I + g];
      rbase[I] = irow[I] * PLD;
      icol[I] = *reinterpret_cast<const int2*>(&S.ids[buf][warp][8 * I + 2 * t]);
    }
    const bool fastg = pooled && n == KMAX;
    // gather block (I,J): I < NBC covariance block, I >= NBC right-hand-side rows
    auto gath = [&](int I, int J) -> double2 {
      if (I >= NBC) {
        const int rrow = 8 * (I - NBC) + g;
        double2 val = make_double2(0.0, 0.0);
        if (rrow < NR) val = *reinterpret_cast<const double2*>(&S.rhs[buf][warp][rrow][8 * J + 2 * t]);
        return val;
      }
      if (fastg) {   // every neighbour slot is live and pooled: plain indexed loads
        return make_double2(S.PC[rbase[I] + icol[J].x], S.PC[rbase[I] + icol[J].y]);
      }
      const int row = 8 * I + g, col = 8 * J + 2 * t;
      double c0 = (row == col) ? v.sill : 0.0, c1 = (row == col + 1) ? v.sill : 0.0;   // padding: sill * identity
      if (pooled) {
        const int ir = irow[I], ix = icol[J].x, iy = icol[J].y;
        if (ir >= 0 && ix >= 0) c0 = S.PC[ir * PLD + ix];
        if (ir >= 0 && iy >= 0) c1 = S.PC[ir * PLD + iy];
      } else {
        const double xr = S.px[buf][warp][row], yr = S.py[buf][warp][row], zr = DIM3 ? S.pz[buf][warp][row] : 0.0;
        const double2 xc = *reinterpret_cast<const double2*>(&S.px[buf][warp][col]);
        const double2 yc = *reinterpret_cast<const double2*>(&S.py[buf][warp][col]);
        double2 zc = make_double2(0.0, 0.0);
        if (DIM3) zc = *reinterpret_cast<const double2*>(&S.pz[buf][warp][col]);
        const bool rowlive = row < n;
        double dx = xr - xc.x, dy = yr - yc.x, dz = zr - zc.x;
        const double e0 = cov_fast(cf, DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx));
        dx = xr - xc.y, dy = yr - yc.y, dz = zr - zc.y;
        const double e1 = cov_fast(cf, DIM3 ? fma(dz, dz, fma(dy, dy, dx * dx)) : fma(dy, dy, dx * dx));
        c0 = (row == col) ? v.sill : ((rowlive && col < n) ? e0 : 0.0);
        c1 = (row == col + 1) ? v.sill : ((rowlive && col + 1 < n) ? e1 : 0.0);
      }
      return make_double2(c0, c1);
    };
    auto post = [&](const double2& blk) {
      *reinterpret_cast<double2*>(&S.exin[warp * EXS + g * 8 + 2 * t]) = blk;
      bar_arrive(BAR_READY);
    };
    auto waitW = [&]() -> double2 {
      bar_sync(BAR_DONE);
      return *reinterpret_cast<const double2*>(&S.exout[warp * EXS + g * 8 + 2 * t]);
    };
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

    // ---- left-looking blocked Cholesky; the next group's pre-phase fills the wait slots ----------
    // Schedule of the next-group phases (P0 index loads, P1 records, P23 RHS rows + hash insert,
    // pool barrier, P4 pooled covariances).  The pool barrier / P4 must follow the last gather from the
    // pooled matrix, i.e. the diagonal block of the last covariance column.
    constexpr int KREL = NBC >= 2 ? NBC - 2 : 0;   // column after whose preparation PC is released
    double2 Lb[NROWS][NBC];
    double2 Gb[NRB][NRB];
#pragma unroll
    for (int r = 0; r < NRB; ++r)
#pragma unroll
      for (int q = 0; q < NRB; ++q) Gb[r][q] = make_double2(0.0, 0.0);

    TR(0);
    {
      double2 A00 = gath(0, 0);
      post(A00);
    }
    TR(1);
    P0(grp + gridDim.x);
    if (NBC <= 2) {
      P1();
      P23(buf ^ 1);
    }
#pragma unroll
    for (int kb = 0; kb < NBC; ++kb) {
      // prepare column kb: blocks below the diagonal, and the next diagonal block
      double2 C[NROWS];
#pragma unroll
      for (int i = kb + 1; i < NROWS; ++i) {
        C[i] = gath(i, kb);
#pragma unroll
        for (int j = 0; j < kb; ++j) upd(C[i], Lb[i][j], Lb[kb][j]);
      }
      double2 Dn = make_double2(0.0, 0.0);
      if (kb + 1 < NBC) {
        Dn = gath(kb + 1, kb + 1);
#pragma unroll
        for (int j = 0; j < kb; ++j) upd(Dn, Lb[kb + 1][j], Lb[kb + 1][j]);
      }
      TR(2 + 4 * kb);
      if (kb == KREL) bar_sync_sys();   // next group's hash inserts complete; last PC gather of this group done
      const double2 W = waitW();
      TR(3 + 4 * kb);
      if (kb + 1 < NBC) {
        panel(C[kb + 1], W);
        Lb[kb + 1][kb] = C[kb + 1];
        upd(Dn, C[kb + 1], C[kb + 1]);
        post(Dn);
      }
      TR(4 + 4 * kb);
      // wait-slot work for the next group
      if (NBC >= 3 && kb == 0) P1();
      if ((NBC == 4 && kb == 1) || (NBC == 3 && kb == 0)) P23(buf ^ 1);
      if (kb == KREL) P4(buf ^ 1);
#pragma unroll
      for (int i = (kb + 1 < NBC ? kb + 2 : kb + 1); i < NROWS; ++i) {
        panel(C[i], W);
        Lb[i][kb] = C[i];
      }
#pragma unroll
      for (int r = 0; r < NRB; ++r)
#pragma unroll
        for (int q = 0; q <= r; ++q) upd(Gb[r][q], Lb[NBC + r][kb], Lb[NBC + q][kb]);
      TR(5 + 4 * kb);
    }

    // ---- hand the trailing Schur blocks (-B'C^-1B) to the diagonal warp for the epilogue --------
#pragma unroll
    for (int r = 0; r < NRB; ++r)
#pragma unroll
      for (int q = 0; q <= r; ++q)
        *reinterpret_cast<double2*>(&S.Gx[(warp * NGB + r * (r + 1) / 2 + q) * EXS + g * 8 + 2 * t]) = Gb[r][q];
    if (lane == 0) {
      S.e_tt[warp] = c_tt;
      S.e_flags[warp] = (c_valid ? 1 : 0) | (c_missing ? 2 : 0);
    }
    bar_arrive(BAR_EPI);
    __syncwarp();
    TR(18);
    bar_sync_sys();        // next group's pooled covariances are complete
    P5();
    // With one or two block columns the next group's hash inserts (P23) follow without a CTA-wide barrier in
    // between (with four, the first hand-off lies in between): every warp must have released its entries first.
    if (NBC <= 2) bar_sync_sys();
    TR(19);
  }
}

template <int NCON, bool DIM3, int NBC>
int launch_pipe(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, const TargetsDev& tg, int k, int minn,
                int centered, const int* pos, const int* cnt, double* mean, double* var, unsigned char* status) {
  const long long ngroups = (tg.count + SW - 1) / SW;
  const int per_sm = (sizeof(PipeSmem<2 + NCON>) * 2 <= 220 * 1024) ? 2 : 1;
  const int blocks = (int)std::min<long long>(ngroups, 148LL * per_sm);
  const size_t smem = sizeof(PipeSmem<2 + NCON>);
  auto kern = local_krige_pipe_kernel<NCON, DIM3, NBC>;
  GSM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<blocks, NTH, smem, ctx->stream>>>(ctx->bins, v, m, tg, k, minn, centered, pos, cnt, mean, var, status);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

}  // namespace

// One translation unit per (DIM3, NBC) keeps the build parallel: GSM_PIPE_DIM3 / GSM_PIPE_NBC select it.
#ifndef GSM_PIPE_DIM3
#error "compile through krige_local_pipe_*.cu"
#endif
int GSM_PIPE_ENTRY(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered, const TargetsDev& tg,
                   int k, int minneighbors, const int* d_pos, const int* d_cnt, double* d_mean, double* d_var,
                   unsigned char* d_status) {
#define GSM_LKP_CASE(N) \
  case N:               \
    return launch_pipe<N, GSM_PIPE_DIM3 != 0, GSM_PIPE_NBC>(ctx, v, m, tg, k, minneighbors, centered, d_pos, d_cnt, \
                                                            d_mean, d_var, d_status);
  switch (ncon) {
    GSM_LKP_CASE(0)
    GSM_LKP_CASE(1)
    GSM_LKP_CASE(2)
    GSM_LKP_CASE(3)
    GSM_LKP_CASE(4)
    GSM_LKP_CASE(6)
    GSM_LKP_CASE(10)
  }
#undef GSM_LKP_CASE
  return GSM_ERR_UNSUPPORTED;
}
