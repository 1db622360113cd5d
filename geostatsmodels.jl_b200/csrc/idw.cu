// Inverse distance weighting (reference src/idw.jl:59-89): weights 1/d^e over all samples
// (fitpredictfull, models.jl:200-234) or over the k nearest (fitpredictneigh, models.jl:162-184).
// A zero distance returns the value of the FIRST such sample (idw.jl:68-69).
//
// Global form: one thread per target, sample tiles staged through shared memory with vectorised
// coalesced loads and consumed as warp-wide broadcasts.  N*M pair interactions on O(N+M) bytes: the
// kernel is FP64-pipe bound, not HBM bound.  Local form: O(k) work per target on gathered records.
#include "gsm_internal.cuh"
#include "idw_weight.cuh"

namespace {

constexpr int IDW_THREADS = 128;
constexpr int IDW_TILE = 512;

using gsm_idww::idw_weight;

// IDW_TPB targets per CTA, IDW_SPLIT threads per target: thread (target, q) takes the samples i = q mod IDW_SPLIT
// of every tile, so a 256 x 256 grid (config C1) still fills the SMs; the IDW_SPLIT partial sums of a target are
// added in a fixed order (reproducible).
constexpr int IDW_TPB = 64;
constexpr int IDW_SPLIT = 4;
static_assert(IDW_TPB * IDW_SPLIT == 2 * IDW_THREADS, "thread layout");

template <int EXPMODE, bool DIM3>
__global__ void __launch_bounds__(IDW_TPB * IDW_SPLIT)
idw_global_kernel(const double* __restrict__ coords, const double* __restrict__ values, int dim, int n,
                  TargetsDev tg, double e, int ei, double* __restrict__ out) {
  __shared__ __align__(16) double sx[IDW_TILE], sy[IDW_TILE], sz[IDW_TILE], sv[IDW_TILE];
  __shared__ double psw[IDW_SPLIT][IDW_TPB], pswz[IDW_SPLIT][IDW_TPB], phv[IDW_SPLIT][IDW_TPB];
  __shared__ int phit[IDW_SPLIT][IDW_TPB];
  const int tl = threadIdx.x % IDW_TPB, q = threadIdx.x / IDW_TPB;   // a warp shares q: tile reads stay broadcasts
  const long long t = blockIdx.x * (long long)IDW_TPB + tl;
  const bool active = t < tg.count;
  double tx = 0, ty = 0, tz = 0;
  if (active) gsm_target_coords(tg, tg.first + t, tx, ty, tz);
  double sw = 0.0, swz = 0.0;
  int hit = 0x7fffffff;
  double hitv = 0.0;
  for (int base = 0; base < n; base += IDW_TILE) {
    const int m = min(IDW_TILE, n - base);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += IDW_TPB * IDW_SPLIT) {
      const double* c = coords + (size_t)(base + i) * dim;
      sx[i] = c[0];
      sy[i] = dim > 1 ? c[1] : 0.0;
      sz[i] = dim > 2 ? c[2] : 0.0;
      sv[i] = values[base + i];
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int i = q; i < m; i += IDW_SPLIT) {
        const double dx = tx - sx[i], dy = ty - sy[i];
        double d2 = dx * dx + dy * dy;
        if (DIM3) {
          const double dz = tz - sz[i];
          d2 += dz * dz;
        }
        if (d2 > 0.0) {
          const double w = idw_weight<EXPMODE>(d2, e, ei);
          sw += w;
          swz = fma(w, sv[i], swz);
        } else if (hit == 0x7fffffff) {
          hit = base + i;
          hitv = sv[i];
        }
      }
    }
  }
  psw[q][tl] = sw;
  pswz[q][tl] = swz;
  phit[q][tl] = hit;
  phv[q][tl] = hitv;
  __syncthreads();
  if (q == 0 && active) {
    double a = 0.0, bsum = 0.0, hv = 0.0;
    int h = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < IDW_SPLIT; ++r) {
      a += psw[r][tl];
      bsum += pswz[r][tl];
      if (phit[r][tl] < h) {   // the FIRST zero-distance sample (idw.jl:68-69) = the lowest index over the slices
        h = phit[r][tl];
        hv = phv[r][tl];
      }
    }
    out[t] = (h != 0x7fffffff) ? hv : bsum / a;
  }
}

template <int EXPMODE>
__global__ void __launch_bounds__(IDW_THREADS)
idw_local_kernel(BinsDev b, TargetsDev tg, int k, int minn, double e, int ei, const int* __restrict__ pos,
                 const int* __restrict__ cnt, double* __restrict__ out, unsigned char* __restrict__ status) {
  const long long t = blockIdx.x * (long long)IDW_THREADS + threadIdx.x;
  if (t >= tg.count) return;
  const int n = cnt[t];
  if (n < minn || n <= 0) {
    out[t] = __longlong_as_double(0x7ff8000000000000LL);
    if (status) status[t] = GSM_ST_MISSING;
    return;
  }
  double tx, ty, tz;
  gsm_target_coords(tg, tg.first + t, tx, ty, tz);
  double sw = 0.0, swz = 0.0, hitv = 0.0;
  bool hit = false;
  int hitrow = 0x7fffffff;
  const int* p = pos + t * (long long)k;
  for (int i = 0; i < n; ++i) {
    const double4 r = b.rec[p[i]];
    const double dx = tx - r.x, dy = ty - r.y, dz = tz - r.z;
    const double d2 = dx * dx + dy * dy + dz * dz;
    if (d2 > 0.0) {
      const double w = idw_weight<EXPMODE>(d2, e, ei);
      sw += w;
      swz = fma(w, r.w, swz);
    } else {  // exact hit: the FIRST such sample of the reference's ascending (d2,row) list = lowest row
      const int rw = b.sidx[p[i]];
      if (rw < hitrow) {
        hitrow = rw;
        hit = true;
        hitv = r.w;
      }
    }
  }
  out[t] = hit ? hitv : swz / sw;
  if (status) status[t] = GSM_ST_OK;
}

using gsm_idww::exp_mode;

}  // namespace

int gsm_launch_idw_global(gsm_ctx* ctx, double exponent, const TargetsDev& tg, double* d_mean) {
  if (tg.count <= 0) return GSM_OK;
  int ei;
  const int mode = exp_mode(exponent, ei);
  const unsigned blocks = (unsigned)((tg.count + IDW_TPB - 1) / IDW_TPB);
  const double* c = (const double*)ctx->coords.p;
  const double* v = (const double*)ctx->values.p;
  const int n = (int)ctx->n, dim = ctx->dim;
  switch (mode) {
    case 1: if (dim > 2) idw_global_kernel<1, true><<<blocks, IDW_TPB * IDW_SPLIT, 0, ctx->stream>>>(c, v, dim, n, tg, exponent, ei, d_mean);
      else idw_global_kernel<1, false><<<blocks, IDW_TPB * IDW_SPLIT, 0, ctx->stream>>>(c, v, dim, n, tg, exponent, ei, d_mean);
      break;
    case 2: if (dim > 2) idw_global_kernel<2, true><<<blocks, IDW_TPB * IDW_SPLIT, 0, ctx->stream>>>(c, v, dim, n, tg, exponent, ei, d_mean);
      else idw_global_kernel<2, false><<<blocks, IDW_TPB * IDW_SPLIT, 0, ctx->stream>>>(c, v, dim, n, tg, exponent, ei, d_mean);
      break;
    case 3: if (dim > 2) idw_global_kernel<3, true><<<blocks, IDW_TPB * IDW_SPLIT, 0, ctx->stream>>>(c, v, dim, n, tg, exponent, ei, d_mean);
      else idw_global_kernel<3, false><<<blocks, IDW_TPB * IDW_SPLIT, 0, ctx->stream>>>(c, v, dim, n, tg, exponent, ei, d_mean);
      break;
    default: if (dim > 2) idw_global_kernel<0, true><<<blocks, IDW_TPB * IDW_SPLIT, 0, ctx->stream>>>(c, v, dim, n, tg, exponent, ei, d_mean);
      else idw_global_kernel<0, false><<<blocks, IDW_TPB * IDW_SPLIT, 0, ctx->stream>>>(c, v, dim, n, tg, exponent, ei, d_mean);
      break;
  }
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}

int gsm_launch_idw_local(gsm_ctx* ctx, double exponent, const TargetsDev& tg, int k, int minneighbors,
                         const int* d_pos, const int* d_cnt, double* d_mean, unsigned char* d_status) {
  if (tg.count <= 0) return GSM_OK;
  int ei;
  const int mode = exp_mode(exponent, ei);
  const unsigned blocks = (unsigned)((tg.count + IDW_THREADS - 1) / IDW_THREADS);
  switch (mode) {
    case 1: idw_local_kernel<1><<<blocks, IDW_THREADS, 0, ctx->stream>>>(ctx->bins, tg, k, minneighbors, exponent, ei, d_pos, d_cnt, d_mean, d_status); break;
    case 2: idw_local_kernel<2><<<blocks, IDW_THREADS, 0, ctx->stream>>>(ctx->bins, tg, k, minneighbors, exponent, ei, d_pos, d_cnt, d_mean, d_status); break;
    case 3: idw_local_kernel<3><<<blocks, IDW_THREADS, 0, ctx->stream>>>(ctx->bins, tg, k, minneighbors, exponent, ei, d_pos, d_cnt, d_mean, d_status); break;
    default: idw_local_kernel<0><<<blocks, IDW_THREADS, 0, ctx->stream>>>(ctx->bins, tg, k, minneighbors, exponent, ei, d_pos, d_cnt, d_mean, d_status); break;
  }
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  return GSM_OK;
}
