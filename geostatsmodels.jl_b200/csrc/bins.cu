// Grid-bin index over the sample coordinates: the B200 analogue of the KNearestSearch construction in
// fitpredictneigh (reference src/models.jl:153-159).  Samples are counting-sorted by cell into packed
// double4 records {x, y, z, value} so a neighbour gather touches one 32-byte record per sample.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>

#include <cstdlib>

#include "gsm_internal.cuh"

namespace {

// order-preserving map double -> signed 64-bit (for atomicMin/Max on coordinates)
__device__ __forceinline__ long long enc(double d) {
  long long b = __double_as_longlong(d);
  return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
__host__ __device__ inline double dec(long long b) {
  long long r = b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
#ifdef __CUDA_ARCH__
  return __longlong_as_double(r);
#else
  double d;
  memcpy(&d, &r, sizeof(d));
  return d;
#endif
}

// red[0..2] = min, red[3..5] = max (encoded), red[6] = sum of values (double bits via atomicAdd)
__global__ void bbox_kernel(const double* __restrict__ coords, const double* __restrict__ values, int dim,
                            long long n, long long* __restrict__ red, double* __restrict__ vsum) {
  long long mn[3], mx[3];
  for (int d = 0; d < 3; ++d) {
    mn[d] = 0x7fffffffffffffffLL;
    mx[d] = -0x7fffffffffffffffLL - 1;
  }
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    for (int d = 0; d < dim; ++d) {
      long long e = enc(coords[i * dim + d]);
      mn[d] = min(mn[d], e);
      mx[d] = max(mx[d], e);
    }
    s += values[i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    for (int d = 0; d < 3; ++d) {
      mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if ((threadIdx.x & 31) == 0) {
    for (int d = 0; d < dim; ++d) {
      atomicMin(&red[d], mn[d]);
      atomicMax(&red[3 + d], mx[d]);
    }
    atomicAdd(vsum, s);
  }
}

__device__ __forceinline__ int cell_coord(double x, double lo, double inv_h, int nc) {
  int c = (int)floor((x - lo) * inv_h);
  return min(max(c, 0), nc - 1);
}

__global__ void count_kernel(const double* __restrict__ coords, int dim, int n, BinsDev b, int* __restrict__ cell_of,
                             int* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx = cell_coord(coords[(size_t)i * dim], b.bbmin[0], b.inv_h, b.nc[0]);
  int cy = dim > 1 ? cell_coord(coords[(size_t)i * dim + 1], b.bbmin[1], b.inv_h, b.nc[1]) : 0;
  int cz = dim > 2 ? cell_coord(coords[(size_t)i * dim + 2], b.bbmin[2], b.inv_h, b.nc[2]) : 0;
  int c = cx + b.nc[0] * (cy + b.nc[1] * cz);
  cell_of[i] = c;
  atomicAdd(&counts[c], 1);
}

// Deterministic scatter: within a cell, records keep ascending original row order.  One thread per
// cell walks nothing; instead every sample computes its rank among same-cell samples with a smaller
// row by claiming slots in row order: we launch over samples in row order and use the fact that
// atomicAdd returns a unique slot; a final per-cell insertion sort (cells hold a handful of samples)
// restores row order so the layout -- and therefore every downstream rounding -- is reproducible.
__global__ void scatter_kernel(const double* __restrict__ coords, const double* __restrict__ values, int dim, int n,
                               const int* __restrict__ cell_of, const int* __restrict__ cell_start,
                               int* __restrict__ fill, double4* __restrict__ rec, int* __restrict__ sidx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cell_of[i];
  int slot = cell_start[c] + atomicAdd(&fill[c], 1);
  double4 r;
  r.x = coords[(size_t)i * dim];
  r.y = dim > 1 ? coords[(size_t)i * dim + 1] : 0.0;
  r.z = dim > 2 ? coords[(size_t)i * dim + 2] : 0.0;
  r.w = values[i];
  rec[slot] = r;
  sidx[slot] = i;
}

__global__ void sort_cells_kernel(int ncells, const int* __restrict__ cell_start, double4* __restrict__ rec,
                                  int* __restrict__ sidx) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  int b = cell_start[c], e = cell_start[c + 1];
  for (int i = b + 1; i < e; ++i) {
    int key = sidx[i];
    double4 r = rec[i];
    int j = i - 1;
    while (j >= b && sidx[j] > key) {
      sidx[j + 1] = sidx[j];
      rec[j + 1] = rec[j];
      --j;
    }
    sidx[j + 1] = key;
    rec[j + 1] = r;
  }
}

}  // namespace

// Average samples per cell the bins aim for.  Tuned on B200 (kNN kernel).
static double target_occupancy(const gsm_ctx* ctx, int dim) {
  if (ctx->opt_bin_occupancy > 0.0) return ctx->opt_bin_occupancy;   // developer option "bin_occupancy" (tuning)
  return dim >= 3 ? 2.0 : 3.0;   // (round-2 sweep with the warp-uniform search: flat optima, tools/sweep_bins.py)
}

int gsm_build_bins(gsm_ctx* ctx) {
  const int dim = ctx->dim;
  const long long n = ctx->n;
  cudaStream_t st = ctx->stream;
  if (n > 0x7fffffffLL) {
    gsm_set_error(ctx, "more than 2^31-1 samples are not supported");
    return GSM_ERR_UNSUPPORTED;
  }
  GSM_CHECK(gsm_reserve(ctx, ctx->red, 8 * sizeof(long long)));
  long long init[8];
  for (int d = 0; d < 3; ++d) {
    init[d] = 0x7fffffffffffffffLL;
    init[3 + d] = -0x7fffffffffffffffLL - 1;
  }
  init[6] = 0;  // value sum (double 0.0 bits)
  init[7] = 0;
  GSM_CUDA(ctx, cudaMemcpyAsync(ctx->red.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
  int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  bbox_kernel<<<blocks, 256, 0, st>>>((const double*)ctx->coords.p, (const double*)ctx->values.p, dim, n,
                                      (long long*)ctx->red.p, (double*)((long long*)ctx->red.p + 6));
  ctx->tim.launches++;
  long long red[8];
  GSM_CUDA(ctx, cudaMemcpyAsync(red, ctx->red.p, sizeof(red), cudaMemcpyDeviceToHost, st));
  GSM_CUDA(ctx, cudaStreamSynchronize(st));

  BinsDev& b = ctx->bins;
  double ext[3] = {0, 0, 0};
  double vol = 1.0;
  int live = 0;
  for (int d = 0; d < 3; ++d) {
    if (d < dim) {
      b.bbmin[d] = dec(red[d]);
      ext[d] = dec(red[3 + d]) - b.bbmin[d];
      if (ext[d] > 0) {
        vol *= ext[d];
        live++;
      }
    } else {
      b.bbmin[d] = 0.0;
    }
  }
  double vsum;
  memcpy(&vsum, &red[6], sizeof(double));
  ctx->sample_mean = vsum / (double)n;

  double h;
  if (live == 0) {
    h = 1.0;
  } else {
    h = std::pow(target_occupancy(ctx, live) * vol / (double)n, 1.0 / live);
  }
  // cap the total number of cells (memory) by growing h if needed
  const double max_cells = 64.0 * 1024 * 1024;
  for (;;) {
    double cells = 1.0;
    for (int d = 0; d < dim; ++d) cells *= std::max(1.0, std::ceil(ext[d] / h));
    if (cells <= max_cells) break;
    h *= 1.26;
  }
  b.h = h;
  b.inv_h = 1.0 / h;
  long long ncells = 1;
  for (int d = 0; d < 3; ++d) {
    b.nc[d] = d < dim ? (int)std::max(1.0, std::ceil(ext[d] / h)) : 1;
    // a sample sitting exactly on the upper bound must land in the last cell: cell_coord clamps.
    ncells *= b.nc[d];
  }
  b.n = (int)n;

  GSM_CHECK(gsm_reserve(ctx, ctx->cell_of, n * sizeof(int)));
  GSM_CHECK(gsm_reserve(ctx, ctx->cell_start, (ncells + 1) * sizeof(int)));
  GSM_CHECK(gsm_reserve(ctx, ctx->cell_fill, (ncells + 1) * sizeof(int)));
  GSM_CHECK(gsm_reserve(ctx, ctx->rec, n * sizeof(double4)));
  GSM_CHECK(gsm_reserve(ctx, ctx->sidx, n * sizeof(int)));
  int* counts = (int*)ctx->cell_fill.p;
  GSM_CUDA(ctx, cudaMemsetAsync(counts, 0, (ncells + 1) * sizeof(int), st));
  int nb = (int)((n + 255) / 256);
  count_kernel<<<nb, 256, 0, st>>>((const double*)ctx->coords.p, dim, (int)n, b, (int*)ctx->cell_of.p, counts);
  ctx->tim.launches++;
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, (int*)ctx->cell_start.p, (int)(ncells + 1), st);
  GSM_CHECK(gsm_reserve(ctx, ctx->cub_tmp, tmp_bytes));
  GSM_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, counts, (int*)ctx->cell_start.p,
                                              (int)(ncells + 1), st));
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaMemsetAsync(counts, 0, (ncells + 1) * sizeof(int), st));
  scatter_kernel<<<nb, 256, 0, st>>>((const double*)ctx->coords.p, (const double*)ctx->values.p, dim, (int)n,
                                     (const int*)ctx->cell_of.p, (const int*)ctx->cell_start.p, counts,
                                     (double4*)ctx->rec.p, (int*)ctx->sidx.p);
  ctx->tim.launches++;
  sort_cells_kernel<<<(int)((ncells + 255) / 256), 256, 0, st>>>((int)ncells, (const int*)ctx->cell_start.p,
                                                                (double4*)ctx->rec.p, (int*)ctx->sidx.p);
  ctx->tim.launches++;
  GSM_CUDA(ctx, cudaGetLastError());
  b.cell_start = (const int*)ctx->cell_start.p;
  b.rec = (const double4*)ctx->rec.p;
  b.sidx = (const int*)ctx->sidx.p;
  return GSM_OK;
}
