// Exact k-nearest-neighbour search over the grid bins: the B200 replacement for Meshes `search!`
// on a KNearestSearch / KBallSearch (reference src/models.jl:153-164).
//
// One thread walks KNN_RUN consecutive targets.  A max-heap of the best k candidates lives in shared
// memory with layout [slot][thread] so every lane owns one bank pair and heap traffic is conflict-free
// whatever slot a lane touches.
//   cold start : cells are visited in Chebyshev rings around the target's cell until the current k-th
//                distance is strictly below the distance to the nearest unvisited cell face;
//   warm start : the previous target's k neighbours are k valid candidates for the next target, so the
//                largest of their recomputed distances bounds the new k-th distance; only the cells
//                overlapping that ball are scanned (cell binning is monotone in each coordinate, so the
//                cell range of [t-R, t+R] contains every sample within R).
// Total order: ascending (d2, original row); d2 = ((dx*dx+dy*dy)+dz*dz) with every op rounded once
// (gsm_d2_key) so the selected SET is bit-identical to the oracle's brute-force lexsort.
// Output per target: positions into the cell-sorted record array in heap order (a deterministic
// function of the inputs); sort_lists_kernel orders them ascending when the caller asks for lists.
#include "gsm_internal.cuh"
#include "idw_weight.cuh"

namespace {

constexpr int KNN_THREADS = 128;
constexpr int KNN_RUN = 1;   // consecutive targets per thread (warm start only pays when lanes stay spatially coherent)

// Max-heap of the best k candidates, [slot][thread] in shared memory.  Keys are the squared distances ROUNDED TO
// FLOAT (8 bytes per slot instead of 12: six CTAs per SM instead of four, and this kernel's speed is proportional to
// its occupancy).  Rounding is monotone, so different float keys order like the exact distances; equal float keys
// fall back to the exact double distances recomputed from the records, then to the sample rows -- the comparison is
// the exact total order (d2, row) in every case, and the selected set stays bit-identical to the oracle's.
struct Heap {
  float* d;   // [k][KNN_THREADS] squared distance, rounded to nearest float
  int* p;     // [k][KNN_THREADS] positions into rec
  int k;
  __device__ __forceinline__ float& D(int i) { return d[i * KNN_THREADS]; }
  __device__ __forceinline__ int& P(int i) { return p[i * KNN_THREADS]; }
};
struct TieCtx {   // what the exact comparison needs
  const double4* __restrict__ rec;
  const int* __restrict__ sidx;
  double tx, ty, tz;
};
// (out of line and with by-value arguments: the rare path must not pin the context into local memory)
__device__ __forceinline__ bool key_less_exact(int pa, int pb, const double4* __restrict__ rec, const int* __restrict__ sidx,
                                            double tx, double ty, double tz) {
  const double4 ra = rec[pa], rb = rec[pb];
  const double da = gsm_d2_key(ra.x, ra.y, ra.z, tx, ty, tz), db = gsm_d2_key(rb.x, rb.y, rb.z, tx, ty, tz);
  if (da != db) return da < db;
  return sidx[pa] < sidx[pb];
}
// (d2a, rowa) < (d2b, rowb)?
__device__ __forceinline__ bool key_less(float fa, int pa, float fb, int pb, const TieCtx& c) {
  if (fa != fb) return fa < fb;
  return key_less_exact(pa, pb, c.rec, c.sidx, c.tx, c.ty, c.tz);
}
// upper bound of the exact squared distance behind a float key (round-to-nearest: d2 <= next float above)
__device__ __forceinline__ double key_upper(float f) { return (double)__uint_as_float(__float_as_uint(f) + 1u); }

__device__ __forceinline__ void sift_down(Heap& hp, int i, int n, float d, int p, const TieCtx& tc) {
  // place (d,p) at the hole i of a max-heap of size n.  Positions are only loaded when an entry moves (or on a float
  // tie): two key loads and one position load per level.
  for (;;) {
    int c = 2 * i + 1;
    if (c >= n) break;
    float dc = hp.D(c);
    if (c + 1 < n) {
      const float dr = hp.D(c + 1);
      const bool right = (dc != dr) ? dc < dr : key_less_exact(hp.P(c), hp.P(c + 1), tc.rec, tc.sidx, tc.tx, tc.ty, tc.tz);
      if (right) {
        dc = dr;
        c = c + 1;
      }
    }
    const int pc = hp.P(c);
    if (!key_less(d, p, dc, pc, tc)) break;
    hp.D(i) = dc;
    hp.P(i) = pc;
    i = c;
  }
  hp.D(i) = d;
  hp.P(i) = p;
}

__device__ __forceinline__ void sift_up(Heap& hp, int i, float d, int p, const TieCtx& tc) {
  while (i > 0) {
    int par = (i - 1) >> 1;
    float dp = hp.D(par);
    int pp = hp.P(par);
    if (!key_less(dp, pp, d, p, tc)) break;
    hp.D(i) = dp;
    hp.P(i) = pp;
    i = par;
  }
  hp.D(i) = d;
  hp.P(i) = p;
}

// Local IDW fused into the search (gsm_idw with a neighbourhood, models.jl:162-184 with idw.jl:59-89): the weights
// only need the k selected (d2, value) pairs, which the heap already holds, so nothing but the prediction leaves.
struct KnnIdw {
  int on;                 // 0: emit the neighbour lists; 1: emit IDW predictions
  int mode, ei, minn;     // exponent class (gsm_idww::exp_mode), integer exponent, minneighbors
  double e;
  double* mean;
  unsigned char* status;
};

__global__ void __launch_bounds__(KNN_THREADS, 6)
knn_kernel(BinsDev b, TargetsDev tg, int k, double radius, int tiled, long long row0, int ntx, int nty, KnnIdw idw,
           int* __restrict__ out_pos, int* __restrict__ out_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Heap hp;
  hp.d = reinterpret_cast<float*>(smem_raw) + threadIdx.x;
  hp.p = reinterpret_cast<int*>(smem_raw + (size_t)k * KNN_THREADS * sizeof(float)) + threadIdx.x;
  hp.k = k;

  const int* __restrict__ cs = b.cell_start;
  const double4* __restrict__ rec = b.rec;
  const int* __restrict__ sidx = b.sidx;
  const double inf = 1e308 * 10.0;
  const int maxring = max(max(b.nc[0], b.nc[1]), b.nc[2]);

  // Thread -> target map.  Grid targets: a warp covers an 8 x 4 patch of the grid (CTA = 16 x 8), so its
  // lanes sit in the same one or two cells, walk the same rings and turn candidate loads into broadcasts.
  // 3-D grids spanning several planes: a 4 x 4 x 2 brick per warp (tiled == 2).  Point targets: consecutive indices.
  long long t0;
  if (tiled == 1) {
    const int tid = threadIdx.x, lx = tid & 7, ly = (tid >> 3) & 3, wq = tid >> 5;
    const long long bx = blockIdx.x % ntx, by = blockIdx.x / ntx;
    const long long x = bx * 16 + (wq & 1) * 8 + lx;
    const long long row = row0 + by * 8 + (wq >> 1) * 4 + ly;
    const long long lin = row * tg.dims[0] + x;
    t0 = (x < tg.dims[0] && lin >= tg.first) ? lin - tg.first : tg.count;   // out of range -> skipped below
  } else if (tiled == 2) {
    // 3-D grids: a warp covers a 4 x 4 x 2 brick (CTA = 8 x 8 x 2); row0 is the first plane pair of the slab
    const int tid = threadIdx.x, ln = tid & 31, wq = tid >> 5;
    const int lx = ln & 3, ly = (ln >> 2) & 3, lz = ln >> 4;
    const long long bx = blockIdx.x % ntx, r = blockIdx.x / ntx, by = r % nty, bz = r / nty;
    const long long x = bx * 8 + (wq & 1) * 4 + lx, y = by * 8 + (wq >> 1) * 4 + ly, z = (row0 + bz) * 2 + lz;
    const long long lin = x + tg.dims[0] * (y + tg.dims[1] * z);
    const bool in = x < tg.dims[0] && y < tg.dims[1] && z < tg.dims[2] && lin >= tg.first && lin < tg.first + tg.count;
    t0 = in ? lin - tg.first : tg.count;
  } else {
    t0 = (blockIdx.x * (long long)KNN_THREADS + threadIdx.x) * KNN_RUN;
  }
  int cnt = 0;              // heap fill (k when complete)

  // ---- per-target epilogue: NN value, fused local IDW, or the neighbour list (heap order) -------------------
  auto emit = [&](const long long t, const int cnt, const TieCtx& tie) {
    // the heap keeps float keys: whatever needs a distance recomputes the exact one from the record
    auto exact_d2 = [&](int q) -> double {
      const double4 r = rec[q];
      return gsm_d2_key(r.x, r.y, r.z, tie.tx, tie.ty, tie.tz);
    };
    if (idw.on && idw.mode < 0) {
      // ---- NN model (nn.jl:47-54): the nearest sample's value; among equidistant samples the lowest row ----
      int best = -1;
      for (int i = 0; i < cnt; ++i)
        if (best < 0 || key_less(hp.D(i), hp.P(i), hp.D(best), hp.P(best), tie)) best = i;
      idw.mean[t] = best >= 0 ? rec[hp.P(best)].w : __longlong_as_double(0x7ff8000000000000LL);
      if (idw.status) idw.status[t] = best >= 0 ? GSM_ST_OK : GSM_ST_MISSING;
      out_cnt[t] = cnt;
      return;
    }
    if (idw.on) {
      // ---- fused local IDW over the selected neighbours, in heap order (the order idw_local_kernel would see) ----
      double sw = 0.0, swz = 0.0, hitv = 0.0;
      int nkeep = 0, hitrow = 0x7fffffff;
      for (int i = 0; i < cnt; ++i) {
        const int q = hp.P(i);
        const double d2 = exact_d2(q);
        if (radius > 0.0 && !(__dsqrt_rn(d2) <= radius)) continue;
        nkeep++;
        const double val = rec[q].w;
        if (d2 > 0.0) {
          const double w = gsm_idww::idw_weight_rt(idw.mode, d2, idw.e, idw.ei);
          sw += w;
          swz = fma(w, val, swz);
        } else {   // exact hit: the FIRST such sample of the reference's ascending (d2, row) list = lowest row
          const int rw = sidx[q];
          if (rw < hitrow) {
            hitrow = rw;
            hitv = val;
          }
        }
      }
      out_cnt[t] = nkeep;
      if (nkeep < idw.minn || nkeep <= 0) {
        idw.mean[t] = __longlong_as_double(0x7ff8000000000000LL);
        if (idw.status) idw.status[t] = GSM_ST_MISSING;
      } else {
        idw.mean[t] = hitrow != 0x7fffffff ? hitv : swz / sw;
        if (idw.status) idw.status[t] = GSM_ST_OK;
      }
      return;
    }
    // ---- emit (heap order); KBallSearch keeps neighbours with distance <= radius (models.jl:158) ----
    int* o = out_pos + t * (long long)k;
    int nkeep = 0;
    if (radius > 0.0) {
      for (int i = 0; i < cnt; ++i)
        if (__dsqrt_rn(exact_d2(hp.P(i))) <= radius) o[nkeep++] = hp.P(i);
    } else {
      for (int i = 0; i < cnt; ++i) o[i] = hp.P(i);
      nkeep = cnt;
    }
    for (int i = nkeep; i < k; ++i) o[i] = -1;
    out_cnt[t] = nkeep;
  };

  if (tiled != 0) {
    // ---- grid targets: WARP-UNIFORM search.  The 32 targets of a warp form a compact brick, so their searches
    // visit nearly the same cells.  Instead of 32 divergent ring walks, the warp walks ONE ring sequence around
    // the brick's cell box; every candidate record is loaded once (a broadcast) and tested by all lanes against
    // their own heap.  Cells are skipped / x-runs clipped with the brick's bounding box and the LARGEST current
    // k-th distance of the warp (a candidate farther than that from the box is farther than every lane's k-th
    // distance from its target).  A lane stops changing once its own k-th distance is inside the covered box; the
    // warp stops when every lane has.
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool active = t0 < tg.count;
    const unsigned am = __ballot_sync(FULL, active);
    if (am == 0u) return;
    double tx = 0.0, ty = 0.0, tz = 0.0;
    if (active) gsm_target_coords(tg, tg.first + t0, tx, ty, tz);
    {   // idle lanes shadow an active one: they neither widen the brick nor delay the end of the search
      const int src = active ? lane : __ffs(am) - 1;
      tx = __shfl_sync(FULL, tx, src);
      ty = __shfl_sync(FULL, ty, src);
      tz = __shfl_sync(FULL, tz, src);
    }
    const double tc[3] = {tx, ty, tz};
    int lo[3], hi[3];
    double bmin[3], bmax[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      int c = (int)fmax(-1.0, fmin(floor((tc[d] - b.bbmin[d]) * b.inv_h), (double)b.nc[d]));
      c = min(max(c, 0), b.nc[d] - 1);
      lo[d] = __reduce_min_sync(FULL, c);
      hi[d] = __reduce_max_sync(FULL, c);
      double mn = tc[d], mx = tc[d];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(FULL, mn, o));
        mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
      }
      bmin[d] = mn;
      bmax[d] = mx;
    }
    int cnt = 0;              // heap fill: the same on every lane (all lanes see the same candidates)
    double worst = inf;
    // candidate records are staged through shared memory 32 at a time: one coalesced load per chunk instead of a
    // dependent global load per candidate
    double4* stage = reinterpret_cast<double4*>(smem_raw + (size_t)k * KNN_THREADS * (sizeof(float) + sizeof(int))) +
                     (threadIdx.x >> 5) * 32;
    const TieCtx tie{rec, sidx, tx, ty, tz};
    // `worst`: an upper bound of the exact k-th squared distance (next float above the root's key)
    auto consider = [&](int q, const double d2) {
      if (cnt < k) {
        sift_up(hp, cnt, __double2float_rn(d2), q, tie);
        cnt++;
        if (cnt == k) worst = key_upper(hp.D(0));
      } else if (d2 <= worst) {
        const float f = __double2float_rn(d2);
        if (key_less(f, q, hp.D(0), hp.P(0), tie)) {
          sift_down(hp, 0, k, f, q, tie);
          worst = key_upper(hp.D(0));
        }
      }
    };
    for (int ring = 0; ring <= maxring; ++ring) {
      double wmax = inf;
      if (cnt == k) {
        wmax = worst;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(FULL, wmax, o));
      }
      const int zlo = max(lo[2] - ring, 0), zhi = min(hi[2] + ring, b.nc[2] - 1);
      const int ylo = max(lo[1] - ring, 0), yhi = min(hi[1] + ring, b.nc[1] - 1);
      const int xlo = max(lo[0] - ring, 0), xhi = min(hi[0] + ring, b.nc[0] - 1);
      for (int cz = zlo; cz <= zhi; ++cz) {
        const bool zface = (cz == lo[2] - ring) || (cz == hi[2] + ring);
        for (int cy = ylo; cy <= yhi; ++cy) {
          const bool yface = (cy == lo[1] - ring) || (cy == hi[1] + ring);
          const int rowbase = b.nc[0] * (cy + b.nc[1] * cz);
          // on a z- or y-face of the shell the whole x-run is new; otherwise only its two end cells
          int nseg, segx[2][2];
          if (zface || yface || ring == 0) {
            nseg = 1;
            segx[0][0] = xlo;
            segx[0][1] = xhi;
          } else {
            nseg = 0;
            if (lo[0] - ring >= 0) {
              segx[nseg][0] = segx[nseg][1] = lo[0] - ring;
              nseg++;
            }
            if (hi[0] + ring <= b.nc[0] - 1) {
              segx[nseg][0] = segx[nseg][1] = hi[0] + ring;
              nseg++;
            }
          }
          int clo = 0, chi = b.nc[0] - 1;
          if (cnt == k) {
            const double ylo_ = b.bbmin[1] + (double)cy * b.h, zlo_ = b.bbmin[2] + (double)cz * b.h;
            const double gy = fmax(0.0, fmax(ylo_ - bmax[1], bmin[1] - (ylo_ + b.h))) - 1e-9 * b.h;
            const double gz = fmax(0.0, fmax(zlo_ - bmax[2], bmin[2] - (zlo_ + b.h))) - 1e-9 * b.h;
            const double gyz2 = (gy > 0.0 ? gy * gy : 0.0) + (gz > 0.0 ? gz * gz : 0.0);
            if (gyz2 > wmax) continue;
            const double w = __dsqrt_ru(wmax - gyz2) * (1.0 + 1e-12) + 1e-9 * b.h;
            clo = (int)fmax(-1.0, fmin(floor((bmin[0] - w - b.bbmin[0]) * b.inv_h), (double)b.nc[0]));
            chi = (int)fmax(-1.0, fmin(floor((bmax[0] + w - b.bbmin[0]) * b.inv_h), (double)b.nc[0]));
          }
          for (int sgi = 0; sgi < nseg; ++sgi) {
            const int x0 = max(segx[sgi][0], clo), x1 = min(segx[sgi][1], chi);
            if (x0 > x1) continue;
            const int beg = cs[rowbase + x0];
            const int end = cs[rowbase + x1 + 1];
            for (int q0 = beg; q0 < end; q0 += 32) {
              const int nq = min(32, end - q0);
              __syncwarp();
              if (lane < nq) stage[lane] = rec[q0 + lane];
              __syncwarp();
              int j = 0;
              for (; j < nq; ++j) {
                const double4 r = stage[j];
                consider(q0 + j, gsm_d2_key(r.x, r.y, r.z, tx, ty, tz));
              }
            }
          }
        }
      }
      // per lane: distance from its target to the nearest face of the covered box with unvisited cells behind it
      double reach = inf;
      bool any = false;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (lo[d] - ring > 0) {
          reach = fmin(reach, tc[d] - (b.bbmin[d] + (double)(lo[d] - ring) * b.h));
          any = true;
        }
        if (hi[d] + ring < b.nc[d] - 1) {
          reach = fmin(reach, (b.bbmin[d] + (double)(hi[d] + ring + 1) * b.h) - tc[d]);
          any = true;
        }
      }
      if (!any) break;  // every cell visited (uniform: lo / hi / ring are)
      bool done = false;
      if (cnt == k) {
        const double safe = reach - 1e-9 * b.h;   // a sample can sit a rounding error outside its nominal cell box
        done = safe > 0.0 && worst < safe * safe;
      }
      if (__all_sync(FULL, done)) break;
    }
    if (active) emit(t0, cnt, tie);
    return;
  }

  for (int run = 0; run < KNN_RUN; ++run) {
    const long long t = t0 + run;
    if (t >= tg.count) break;
    double tx, ty, tz;
    gsm_target_coords(tg, tg.first + t, tx, ty, tz);
    const double tc[3] = {tx, ty, tz};

    // (a warm start from the previous target's neighbour set was tried in round 1 and dropped: DESIGN.md section 9)
    cnt = 0;
    double worst = inf;   // an upper bound of the exact k-th squared distance (next float above the root's key)
    const TieCtx tie{rec, sidx, tx, ty, tz};

    auto consider = [&](int q) {
      const double4 r = rec[q];
      const double d2 = gsm_d2_key(r.x, r.y, r.z, tx, ty, tz);
      if (cnt < k) {
        sift_up(hp, cnt, __double2float_rn(d2), q, tie);
        cnt++;
        if (cnt == k) worst = key_upper(hp.D(0));
      } else if (d2 <= worst) {
        const float f = __double2float_rn(d2);
        if (key_less(f, q, hp.D(0), hp.P(0), tie)) {
          sift_down(hp, 0, k, f, q, tie);
          worst = key_upper(hp.D(0));
        }
      }
    };

    {
      // ---- cold: Chebyshev rings around the target's cell ----------------------------------------
      int c0[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        int c = (int)fmax(-1.0, fmin(floor((tc[d] - b.bbmin[d]) * b.inv_h), (double)b.nc[d]));
        c0[d] = min(max(c, 0), b.nc[d] - 1);
      }
      for (int ring = 0; ring <= maxring; ++ring) {
        const int zlo = max(c0[2] - ring, 0), zhi = min(c0[2] + ring, b.nc[2] - 1);
        const int ylo = max(c0[1] - ring, 0), yhi = min(c0[1] + ring, b.nc[1] - 1);
        const int xlo = max(c0[0] - ring, 0), xhi = min(c0[0] + ring, b.nc[0] - 1);
        for (int cz = zlo; cz <= zhi; ++cz) {
          const bool zface = (cz == c0[2] - ring) || (cz == c0[2] + ring);
          for (int cy = ylo; cy <= yhi; ++cy) {
            const bool yface = (cy == c0[1] - ring) || (cy == c0[1] + ring);
            const int rowbase = b.nc[0] * (cy + b.nc[1] * cz);
            // on a z- or y-face of the shell the whole x-run belongs to the ring; otherwise only its two ends
            int nseg, segx[2][2];
            if (zface || yface || ring == 0) {
              nseg = 1;
              segx[0][0] = xlo;
              segx[0][1] = xhi;
            } else {
              nseg = 0;
              if (c0[0] - ring >= 0) {
                segx[nseg][0] = segx[nseg][1] = c0[0] - ring;
                nseg++;
              }
              if (c0[0] + ring <= b.nc[0] - 1) {
                segx[nseg][0] = segx[nseg][1] = c0[0] + ring;
                nseg++;
              }
            }
            // once k candidates are held, only the part of the row inside the current k-th-distance ball
            // can matter: skip rows farther than that, clip the x-run to the ball's chord (cell binning
            // is monotone in x, so the clipped cell range still contains every sample of the chord)
            int clo = 0, chi = b.nc[0] - 1;
            if (cnt == k) {
              const double ylo_ = b.bbmin[1] + (double)cy * b.h, zlo_ = b.bbmin[2] + (double)cz * b.h;
              const double gy = fmax(0.0, fmax(ylo_ - ty, ty - (ylo_ + b.h))) - 1e-9 * b.h;
              const double gz = fmax(0.0, fmax(zlo_ - tz, tz - (zlo_ + b.h))) - 1e-9 * b.h;
              const double gyz2 = (gy > 0.0 ? gy * gy : 0.0) + (gz > 0.0 ? gz * gz : 0.0);
              if (gyz2 > worst) continue;
              const double w = __dsqrt_ru(worst - gyz2) * (1.0 + 1e-12) + 1e-9 * b.h;
              clo = (int)fmax(-1.0, fmin(floor((tx - w - b.bbmin[0]) * b.inv_h), (double)b.nc[0]));
              chi = (int)fmax(-1.0, fmin(floor((tx + w - b.bbmin[0]) * b.inv_h), (double)b.nc[0]));
            }
            for (int s = 0; s < nseg; ++s) {
              const int x0 = max(segx[s][0], clo), x1 = min(segx[s][1], chi);
              if (x0 > x1) continue;
              const int beg = cs[rowbase + x0];
              const int end = cs[rowbase + x1 + 1];
              for (int q = beg; q < end; ++q) consider(q);
            }
          }
        }
        // distance from the target to the nearest face that still has unvisited cells behind it
        double reach = inf;
        bool any = false;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          if (c0[d] - ring > 0) {
            reach = fmin(reach, tc[d] - (b.bbmin[d] + (double)(c0[d] - ring) * b.h));
            any = true;
          }
          if (c0[d] + ring < b.nc[d] - 1) {
            reach = fmin(reach, (b.bbmin[d] + (double)(c0[d] + ring + 1) * b.h) - tc[d]);
            any = true;
          }
        }
        if (!any) break;  // every cell visited
        if (cnt == k) {
          // conservative margin: a sample can sit a rounding error outside its nominal cell box
          double safe = reach - 1e-9 * b.h;
          if (safe > 0.0 && worst < safe * safe) break;
        }
      }
    }

    emit(t, cnt, tie);
  }
}

// Ascending (d2, row) order of every neighbour list, in place (only when the caller asks for lists).
__global__ void sort_lists_kernel(BinsDev b, TargetsDev tg, int k, int* __restrict__ pos, const int* __restrict__ cnt) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= tg.count) return;
  double tx, ty, tz;
  gsm_target_coords(tg, tg.first + t, tx, ty, tz);
  int* p = pos + t * (long long)k;
  const int n = cnt[t];
  double key[GSM_MAX_NEIGHBORS];
  int row[GSM_MAX_NEIGHBORS], pp[GSM_MAX_NEIGHBORS];
  for (int i = 0; i < n; ++i) {
    const int q = p[i];
    const double4 r = b.rec[q];
    const double d2 = gsm_d2_key(r.x, r.y, r.z, tx, ty, tz);
    const int rw = b.sidx[q];
    int j = i - 1;
    while (j >= 0 && (key[j] > d2 || (key[j] == d2 && row[j] > rw))) {
      key[j + 1] = key[j];
      row[j + 1] = row[j];
      pp[j + 1] = pp[j];
      --j;
    }
    key[j + 1] = d2;
    row[j + 1] = rw;
    pp[j + 1] = q;
  }
  for (int i = 0; i < n; ++i) p[i] = pp[i];
}

__global__ void pos_to_rows_kernel(const int* __restrict__ pos, long long total, const int* __restrict__ sidx,
                                   int* __restrict__ rows) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  int p = pos[i];
  rows[i] = p >= 0 ? sidx[p] : -1;
}

}  // namespace

static int launch_knn_impl(gsm_ctx* ctx, const TargetsDev& tg, int k, double radius, const KnnIdw& idw, int* d_pos,
                           int* d_cnt) {
  if (tg.count <= 0) return GSM_OK;
  size_t smem = (size_t)k * KNN_THREADS * (sizeof(float) + sizeof(int)) + (KNN_THREADS / 32) * 32 * sizeof(double4) +
                (size_t)ctx->opt_knn_pad_smem;   // (developer option: occupancy experiments)
  GSM_CUDA(ctx, cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = (tg.count + (long long)KNN_THREADS * KNN_RUN - 1) / ((long long)KNN_THREADS * KNN_RUN);
  int tiled = 0, ntx = 1;
  long long row0 = 0;
  int nty = 1;
  if (tg.kind == GSM_TARGETS_GRID && KNN_RUN == 1 && tg.dims[0] >= 8) {
    tiled = 1;
    row0 = tg.first / tg.dims[0];
    const long long rlast = (tg.first + tg.count - 1) / tg.dims[0];
    ntx = (int)((tg.dims[0] + 15) / 16);
    blocks = (long long)ntx * ((rlast - row0 + 1 + 7) / 8);
    const long long z_lo = row0 / tg.dims[1], z_hi = rlast / tg.dims[1];
    if (tg.dim >= 3 && tg.dims[2] > 1 && z_hi > z_lo && tg.dims[1] >= 4) {   // bricks pay once the slab spans planes
      tiled = 2;
      row0 = z_lo / 2;
      ntx = (int)((tg.dims[0] + 7) / 8);
      nty = (int)((tg.dims[1] + 7) / 8);
      blocks = (long long)ntx * nty * (z_hi / 2 - row0 + 1);
    }
  }
  knn_kernel<<<(unsigned)blocks, KNN_THREADS, smem, ctx->stream>>>(ctx->bins, tg, k, radius, tiled, row0, ntx, nty, idw,
                                                                  d_pos, d_cnt);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

int gsm_launch_knn(gsm_ctx* ctx, const TargetsDev& tg, int k, double radius, int* d_pos, int* d_cnt) {
  KnnIdw off{};
  return launch_knn_impl(ctx, tg, k, radius, off, d_pos, d_cnt);
}

int gsm_launch_knn_idw(gsm_ctx* ctx, const TargetsDev& tg, int k, double radius, double exponent, int minneighbors,
                       int* d_cnt, double* d_mean, unsigned char* d_status) {
  KnnIdw f{};
  f.on = 1;
  f.mode = gsm_idww::exp_mode(exponent, f.ei);
  f.e = exponent;
  f.minn = minneighbors;
  f.mean = d_mean;
  f.status = d_status;
  return launch_knn_impl(ctx, tg, k, radius, f, nullptr, d_cnt);
}

int gsm_launch_knn_nearest(gsm_ctx* ctx, const TargetsDev& tg, int* d_cnt, double* d_value) {
  KnnIdw f{};
  f.on = 1;
  f.mode = -1;
  f.minn = 1;
  f.mean = d_value;
  return launch_knn_impl(ctx, tg, 1, 0.0, f, nullptr, d_cnt);
}

namespace {
// NN over a neighbour list (fitpredictneigh with the NN model, models.jl:162-184 + nn.jl:47-54): the nearest, in the
// order (d2, row), of the selected samples that has a value; `missing` when fewer than minneighbors were found or none
// of them has a value.
__global__ void nn_local_kernel(BinsDev b, TargetsDev tg, int k, int minn, const int* __restrict__ pos,
                                const int* __restrict__ cnt, double* __restrict__ out, unsigned char* __restrict__ status) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= tg.count) return;
  const int n = cnt[t];
  double tx, ty, tz;
  gsm_target_coords(tg, tg.first + t, tx, ty, tz);
  const int* p = pos + t * (long long)k;
  double best = 0.0, val = __longlong_as_double(0x7ff8000000000000LL);
  int brow = 0x7fffffff;
  bool found = false;
  if (n >= minn)
    for (int i = 0; i < n; ++i) {
      const int q = p[i];
      const double4 r = b.rec[q];
      if (r.w != r.w) continue;   // no value
      const double d2 = gsm_d2_key(r.x, r.y, r.z, tx, ty, tz);
      const int rw = b.sidx[q];
      if (!found || d2 < best || (d2 == best && rw < brow)) {
        found = true;
        best = d2;
        brow = rw;
        val = r.w;
      }
    }
  out[t] = val;
  if (status) status[t] = found ? GSM_ST_OK : GSM_ST_MISSING;
}
}  // namespace

int gsm_launch_nn_local(gsm_ctx* ctx, const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                        double* d_val, unsigned char* d_status) {
  if (tg.count <= 0) return GSM_OK;
  nn_local_kernel<<<(unsigned)((tg.count + 127) / 128), 128, 0, ctx->stream>>>(ctx->bins, tg, k, minneighbors, d_pos, d_cnt,
                                                                             d_val, d_status);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

int gsm_launch_sort_lists(gsm_ctx* ctx, const TargetsDev& tg, int k, int* d_pos, const int* d_cnt) {
  if (tg.count <= 0) return GSM_OK;
  sort_lists_kernel<<<(unsigned)((tg.count + 127) / 128), 128, 0, ctx->stream>>>(ctx->bins, tg, k, d_pos, d_cnt);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

int gsm_launch_nbr_to_rows(gsm_ctx* ctx, const int* d_pos, long long total, int* d_rows) {
  if (total <= 0) return GSM_OK;
  pos_to_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_pos, total, ctx->bins.sidx, d_rows);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}
