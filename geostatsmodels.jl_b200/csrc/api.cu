// C-ABI glue of libgsm_b200.so (declarations and reference citations: include/gsm_b200.h).
// Owns the context, moves data, chunks the target range and sequences the kernels; no arithmetic of
// the hot path lives here.  Never throws; every failure is an int code + gsm_last_error message.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

#include "gsm_internal.cuh"

static thread_local std::string g_create_error;

void gsm_set_error(gsm_ctx* ctx, const std::string& msg) {
  if (ctx) ctx->err = msg;
  else g_create_error = msg;
}

bool gsm_is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int gsm_reserve(gsm_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap && b.p) return GSM_OK;
  if (b.p) GSM_CUDA(ctx, cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = std::max<size_t>(bytes, 256);
  GSM_CUDA(ctx, cudaMalloc(&b.p, want));
  b.cap = want;
  return GSM_OK;
}

static void free_buf(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

VarioDev make_vario(const gsm_variogram& v) { return gsm_make_vario(v); }

int check_vario(gsm_ctx* ctx, const gsm_variogram* v) { return gsm_check_vario(ctx, v); }

int make_model(gsm_ctx* ctx, const gsm_model* m, int dim, ModelDev& d) {
  if (!m || m->kind < GSM_KRIG_SIMPLE || m->kind > GSM_KRIG_UNIVERSAL) {
    gsm_set_error(ctx, "invalid Kriging model descriptor");
    return GSM_ERR_INVALID;
  }
  d.kind = m->kind;
  d.mean = m->mean;
  d.ndrift = 0;
  memset(d.exps, 0, sizeof(d.exps));
  if (m->kind == GSM_KRIG_UNIVERSAL) {
    if (m->ndrift < 1 || m->ndrift > GSM_MAX_DRIFT) {
      gsm_set_error(ctx, "UniversalKriging: ndrift must be in 1..GSM_MAX_DRIFT");
      return GSM_ERR_INVALID;
    }
    d.ndrift = m->ndrift;
    for (int n = 0; n < m->ndrift; ++n)
      for (int c = 0; c < 3; ++c) {
        int e = m->exps[n][c];
        if (e < 0 || (c >= dim && e != 0)) {  // universal.jl:56-57 degree must be nonnegative
          gsm_set_error(ctx, "UniversalKriging: invalid monomial exponent");
          return GSM_ERR_INVALID;
        }
        d.exps[n][c] = e;
      }
  }
  return GSM_OK;
}

long long domain_size(const gsm_targets* t) {
  if (t->kind == GSM_TARGETS_GRID) {
    long long m = 1;
    for (int d = 0; d < t->dim; ++d) m *= t->dims[d];
    return m;
  }
  return t->npoints;
}

// Resolve the descriptor into a device-side view of the shard [first, first+count).
int make_targets(gsm_ctx* ctx, const gsm_targets* t, TargetsDev& d, long long& M) {
  if (!t || (t->kind != GSM_TARGETS_GRID && t->kind != GSM_TARGETS_POINTS) || t->dim < 1 || t->dim > 3) {
    gsm_set_error(ctx, "invalid target domain descriptor");
    return GSM_ERR_INVALID;
  }
  if (t->dim != ctx->dim) {
    gsm_set_error(ctx, "target domain and samples have different dimensions");
    return GSM_ERR_INVALID;
  }
  const long long total = domain_size(t);
  long long first = t->first, count = t->count < 0 ? total - t->first : t->count;
  if (first < 0 || count < 0 || first + count > total) {
    gsm_set_error(ctx, "target shard [first, first+count) is outside the domain");
    return GSM_ERR_INVALID;
  }
  M = count;
  d.kind = t->kind;
  d.dim = t->dim;
  d.points = nullptr;
  for (int c = 0; c < 3; ++c) {
    d.origin[c] = c < t->dim ? t->origin[c] : 0.0;
    d.spacing[c] = c < t->dim ? t->spacing[c] : 0.0;
    d.dims[c] = c < t->dim ? t->dims[c] : 1;
  }
  d.first = first;
  d.count = count;
  if (t->kind == GSM_TARGETS_GRID) {
    for (int c = 0; c < t->dim; ++c)
      if (t->dims[c] < 1) {
        gsm_set_error(ctx, "grid dims must be positive");
        return GSM_ERR_INVALID;
      }
  } else {
    if (!t->points && count > 0) {
      gsm_set_error(ctx, "PointSet targets need a coordinate pointer");
      return GSM_ERR_INVALID;
    }
    const double* src = t->points + first * t->dim;
    if (gsm_is_device_ptr(t->points)) {
      d.points = src;
    } else {
      GSM_CHECK(gsm_reserve(ctx, ctx->tgt_points, (size_t)count * t->dim * sizeof(double)));
      GSM_CUDA(ctx, cudaMemcpyAsync(ctx->tgt_points.p, src, (size_t)count * t->dim * sizeof(double),
                                    cudaMemcpyHostToDevice, ctx->stream));
      d.points = (const double*)ctx->tgt_points.p;
    }
    d.first = 0;  // points view already starts at the shard
  }
  return GSM_OK;
}

void clamp_neighbors(long long nobs, const gsm_neighbor_opts* o, int& maxn, int& minn) {
  long long mx = o->maxneighbors, mn = o->minneighbors;
  if (mx > nobs || mx < 1) mx = nobs;   // models.jl:144-147
  if (mn > mx || mn < 1) mn = 1;        // models.jl:148-150
  maxn = (int)std::min<long long>(mx, 1 << 30);
  minn = (int)mn;
}

// An output the caller may hand us as host or device memory.
template <typename T>
struct OutBuf {
  T* user = nullptr;
  T* dev = nullptr;   // where kernels write
  bool staged = false;
  int init(gsm_ctx* ctx, T* user_ptr, DevBuf& stage, long long M) {
    user = user_ptr;
    if (!user_ptr) {
      dev = nullptr;
      return GSM_OK;
    }
    if (gsm_is_device_ptr(user_ptr)) {
      dev = user_ptr;
      staged = false;
    } else {
      GSM_CHECK(gsm_reserve(ctx, stage, (size_t)std::max<long long>(M, 1) * sizeof(T)));
      dev = (T*)stage.p;
      staged = true;
    }
    return GSM_OK;
  }
  int download(gsm_ctx* ctx, long long off, long long cnt, cudaStream_t st) {
    if (!staged || cnt <= 0) return GSM_OK;
    GSM_CUDA(ctx, cudaMemcpyAsync(user + off, dev + off, (size_t)cnt * sizeof(T), cudaMemcpyDeviceToHost, st));
    return GSM_OK;
  }
};

constexpr long long CHUNK = 1LL << 20;  // targets per kernel wave (bounds the neighbour scratch)

// status / variance summary of one chunk (gsm_summary): counts land in c[0..2]
__global__ void summary_kernel(const unsigned char* __restrict__ st, const double* __restrict__ var, long long n,
                               unsigned long long* __restrict__ c) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int miss = 0, sing = 0, nonpos = 0;
  if (i < n) {
    const int s = st[i];
    miss = s == GSM_ST_MISSING;
    sing = s == GSM_ST_SINGULAR;
    nonpos = (var != nullptr && s == GSM_ST_OK && !(var[i] > 0.0)) ? 1 : 0;
  }
  const unsigned m0 = __ballot_sync(0xffffffffu, miss), m1 = __ballot_sync(0xffffffffu, sing),
                 m2 = __ballot_sync(0xffffffffu, nonpos);
  if ((threadIdx.x & 31) == 0) {
    if (m0) atomicAdd(&c[0], (unsigned long long)__popc(m0));
    if (m1) atomicAdd(&c[1], (unsigned long long)__popc(m1));
    if (m2) atomicAdd(&c[2], (unsigned long long)__popc(m2));
  }
}

// max |coords| (bit pattern of a non-negative double orders like an integer) and NaN count of the values
__global__ void absmax_nan_kernel(const double* __restrict__ c, long long nc, const double* __restrict__ v, long long nv,
                                  unsigned long long* __restrict__ out) {
  double mx = 0.0;
  int nan = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nc; i += (long long)gridDim.x * blockDim.x) {
    mx = fmax(mx, fabs(c[i]));
    if (i < nv && v[i] != v[i]) nan++;
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    nan += __shfl_xor_sync(0xffffffffu, nan, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&out[0], (unsigned long long)__double_as_longlong(mx));
    if (nan) atomicAdd(&out[1], (unsigned long long)nan);
  }
}
__global__ void scale_kernel(double* __restrict__ c, long long nc, double alpha) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < nc) c[i] = __dmul_rn(c[i], alpha);
}

struct EventTimer {
  std::vector<cudaEvent_t> evs;
  cudaEvent_t mark(cudaStream_t st) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    evs.push_back(e);
    return e;
  }
  static double ms(cudaEvent_t a, cudaEvent_t b) {
    float f = 0.f;
    cudaEventElapsedTime(&f, a, b);
    return f;
  }
  ~EventTimer() {
    for (auto e : evs) cudaEventDestroy(e);
  }
};

}  // namespace

// =============================================================================================
extern "C" {

int gsm_abi_version(void) { return GSM_ABI_VERSION; }

int gsm_create(int device, gsm_ctx** out) {
  if (!out) return GSM_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    gsm_set_error(nullptr, std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                               " (libgsm_b200 has no CPU fallback)");
    return GSM_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    gsm_set_error(nullptr, "device ordinal out of range");
    return GSM_ERR_INVALID;
  }
  gsm_ctx* ctx = new gsm_ctx();
  ctx->device = device;
  DeviceGuard g(device);
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    gsm_set_error(nullptr, std::string("stream creation failed: ") + cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return GSM_ERR_CUDA;
  }
  *out = ctx;
  return GSM_OK;
}

int gsm_destroy(gsm_ctx* ctx) {
  if (!ctx) return GSM_OK;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  DevBuf* bufs[] = {&ctx->coords, &ctx->values, &ctx->rec, &ctx->sidx, &ctx->cell_of, &ctx->cell_start,
                    &ctx->cell_fill, &ctx->cub_tmp, &ctx->red, &ctx->block_off, &ctx->nbr_pos, &ctx->nbr_cnt, &ctx->job_idx, &ctx->summ_dev, &ctx->glob_R, &ctx->glob_acc, &ctx->out_mean,
                    &ctx->out_var, &ctx->out_status, &ctx->out_nbr, &ctx->tgt_points};
  for (DevBuf* b : bufs) free_buf(*b);
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return GSM_OK;
}

const char* gsm_last_error(const gsm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gsm_get_summary(const gsm_ctx* ctx, gsm_summary* out) {
  if (!ctx || !out) return GSM_ERR_INVALID;
  *out = ctx->summ;
  return GSM_OK;
}

int gsm_get_timings(const gsm_ctx* ctx, gsm_timings* out) {
  if (!ctx || !out) return GSM_ERR_INVALID;
  *out = ctx->tim;
  return GSM_OK;
}

int gsm_set_option(gsm_ctx* ctx, const char* key, const char* value) {
  if (!ctx || !key || !value) return GSM_ERR_INVALID;
  const std::string k = key, v = value;
  if (k == "local_impl") {
    if (v == "auto") ctx->opt_local_impl = GSM_IMPL_AUTO;
    else if (v == "shr") ctx->opt_local_impl = GSM_IMPL_SHR;
    else if (v == "pipe") ctx->opt_local_impl = GSM_IMPL_PIPE;
    else if (v == "simt") ctx->opt_local_impl = GSM_IMPL_SIMT;
    else if (v == "patch") ctx->opt_local_impl = GSM_IMPL_PATCH;
    else if (v == "patch2") ctx->opt_local_impl = GSM_IMPL_PATCH2;
    else {
      gsm_set_error(ctx, "gsm_set_option: local_impl must be auto | patch | shr | pipe | simt");
      return GSM_ERR_INVALID;
    }
    return GSM_OK;
  }
  if (k == "shr_smax") {
    ctx->opt_shr_smax = std::max(0, std::min(3, atoi(value)));
    return GSM_OK;
  }
  if (k == "knn_pad_smem") {
    ctx->opt_knn_pad_smem = std::max(0, atoi(value));
    return GSM_OK;
  }
  if (k == "patch_nbmax") {
    ctx->opt_patch_nbmax = atoi(value);
    return GSM_OK;
  }
  if (k == "bin_occupancy") {
    ctx->opt_bin_occupancy = atof(value);   // takes effect at the next gsm_set_samples
    return GSM_OK;
  }
  gsm_set_error(ctx, "gsm_set_option: unknown key " + k);
  return GSM_ERR_INVALID;
}

int gsm_set_block_support(gsm_ctx* ctx, const double* offsets, int dim, int nq) {
  if (!ctx) return GSM_ERR_INVALID;
  if (nq == 0) {
    ctx->block_nq = 0;
    return GSM_OK;
  }
  if (!offsets || dim < 1 || dim > 3 || nq < 0 || nq > GSM_MAX_BLOCK_POINTS) {
    gsm_set_error(ctx, "gsm_set_block_support: need offsets, 1 <= dim <= 3 and 0 <= nq <= GSM_MAX_BLOCK_POINTS");
    return GSM_ERR_INVALID;
  }
  std::vector<double> off((size_t)nq * 3, 0.0);
  for (int q = 0; q < nq; ++q)
    for (int d = 0; d < dim; ++d) {
      const double o = offsets[(size_t)q * dim + d];
      if (!std::isfinite(o)) {
        gsm_set_error(ctx, "gsm_set_block_support: offsets must be finite");
        return GSM_ERR_INVALID;
      }
      off[(size_t)q * 3 + d] = o;
    }
  DeviceGuard g(ctx->device);
  GSM_CHECK(gsm_reserve(ctx, ctx->block_off, off.size() * sizeof(double)));
  // (pageable source: cudaMemcpyAsync returns after staging it, so `off` may go out of scope)
  GSM_CUDA(ctx, cudaMemcpyAsync(ctx->block_off.p, off.data(), off.size() * sizeof(double), cudaMemcpyHostToDevice,
                                ctx->stream));
  GSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->block_nq = nq;
  return GSM_OK;
}

int gsm_synchronize(gsm_ctx* ctx) {
  if (!ctx) return GSM_ERR_INVALID;
  DeviceGuard g(ctx->device);
  GSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  GSM_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  return GSM_OK;
}

int gsm_host_alloc(void** out, int64_t bytes) {
  if (!out || bytes < 0) return GSM_ERR_INVALID;
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    gsm_set_error(nullptr, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    cudaGetLastError();
    return GSM_ERR_CUDA;
  }
  return GSM_OK;
}

int gsm_host_free(void* p) {
  if (!p) return GSM_OK;
  return cudaFreeHost(p) == cudaSuccess ? GSM_OK : GSM_ERR_CUDA;
}

}  // extern "C"

int gsm_check_vario(gsm_ctx* ctx, const gsm_variogram* v) {
  bool ok = v != nullptr && v->nstruct >= 0 && v->nstruct <= GSM_MAX_STRUCTURES;
  if (ok && v->nstruct == 0)
    ok = v->kind >= GSM_VARIO_GAUSSIAN && v->kind <= GSM_VARIO_NUGGET && v->range > 0.0 && std::isfinite(v->range) &&
         std::isfinite(v->sill) && std::isfinite(v->nugget);
  for (int i = 0; ok && i < v->nstruct; ++i) {
    const gsm_structure& s = v->structs[i];
    ok = s.kind >= GSM_VARIO_GAUSSIAN && s.kind <= GSM_VARIO_NUGGET && s.range > 0.0 && std::isfinite(s.range) &&
         std::isfinite(s.sill) && std::isfinite(s.nugget);
    for (int e = 0; ok && e < 9; ++e) ok = std::isfinite(s.metric[e]);
  }
  if (!ok) {
    gsm_set_error(ctx, "invalid variogram descriptor");
    return GSM_ERR_INVALID;
  }
  return GSM_OK;
}

namespace {
// covzero of a block target: mean covariance over all pairs of its points (krig.jl:295-301 with
// GeoStatsFunctions' f(g0, g0)); one CTA, fixed summation order
__global__ void __launch_bounds__(256) block_c0_kernel(VarioDev v, const double* __restrict__ off, int nq,
                                                        double* __restrict__ out) {
  __shared__ double part[256];
  double s = 0.0;
  for (int e = threadIdx.x; e < nq * nq; e += 256) {
    const int p = e / nq, q = e - p * nq;
    s += gsm_cov_vec(v, off[3 * p] - off[3 * q], off[3 * p + 1] - off[3 * q + 1], off[3 * p + 2] - off[3 * q + 2]);
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = part[0] / ((double)nq * (double)nq);
}
}  // namespace

int gsm_block_prepare(gsm_ctx* ctx, const VarioDev& v) {
  ctx->block_cur = BlockDev{0, nullptr, v.sill};
  if (ctx->block_nq <= 0) return GSM_OK;
  GSM_CHECK(gsm_reserve(ctx, ctx->red, 64 * sizeof(double)));
  double* d_c0 = (double*)ctx->red.p;
  block_c0_kernel<<<1, 256, 0, ctx->stream>>>(v, (const double*)ctx->block_off.p, ctx->block_nq, d_c0);
  GSM_CUDA(ctx, cudaGetLastError());
  double c0 = 0.0;
  GSM_CUDA(ctx, cudaMemcpyAsync(&c0, d_c0, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->block_cur = BlockDev{ctx->block_nq, (const double*)ctx->block_off.p, c0};
  return GSM_OK;
}

// Phase 1 of gsm_set_samples[_scaled]: upload (and, when `scaled`, the reference's `_scale` on the device).
// coords_dev / values_dev: when non-null the data already sit in ctx->coords / ctx->values (multi-GPU: they arrived
// by NCCL broadcast) and nothing is uploaded.
int gsm_upload_samples(gsm_ctx* ctx, const double* coords, int32_t dim, int64_t n, const double* values, bool scaled,
                       double scale_floor, double* out_alpha, int64_t* out_nan) {
  if (!ctx) return GSM_ERR_INVALID;
  if (!coords || !values || dim < 1 || dim > 3 || n < 1) {
    gsm_set_error(ctx, "gsm_set_samples: need coords, values, 1 <= dim <= 3 and n >= 1");
    return GSM_ERR_INVALID;
  }
  GSM_CHECK(gsm_reserve(ctx, ctx->coords, (size_t)n * dim * sizeof(double)));
  GSM_CHECK(gsm_reserve(ctx, ctx->values, (size_t)n * sizeof(double)));
  GSM_CUDA(ctx, cudaMemcpyAsync(ctx->coords.p, coords, (size_t)n * dim * sizeof(double), cudaMemcpyDefault, ctx->stream));
  GSM_CUDA(ctx, cudaMemcpyAsync(ctx->values.p, values, (size_t)n * sizeof(double), cudaMemcpyDefault, ctx->stream));
  {
    // absmax of the coordinates (for `_scale`) and the number of missing values (NaN), reduced on the device
    GSM_CHECK(gsm_reserve(ctx, ctx->summ_dev, 4 * sizeof(unsigned long long)));
    unsigned long long* d = (unsigned long long*)ctx->summ_dev.p;
    GSM_CUDA(ctx, cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), ctx->stream));
    const long long nc = (long long)n * dim;
    const int blocks = (int)std::min<long long>((nc + 255) / 256, 148 * 8);
    absmax_nan_kernel<<<blocks, 256, 0, ctx->stream>>>((const double*)ctx->coords.p, nc, (const double*)ctx->values.p, n, d);
    unsigned long long h[2] = {0, 0};
    GSM_CUDA(ctx, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    GSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_missing = (long long)h[1];
    if (out_nan) *out_nan = (int64_t)h[1];
    ctx->tim.launches += 1;
    if (scaled) {
      double absmax;
      static_assert(sizeof(double) == sizeof(unsigned long long), "bit cast");
      memcpy(&absmax, &h[0], sizeof(double));
      const double alpha = 1.0 / std::max(scale_floor, absmax);
      if (out_alpha) *out_alpha = alpha;
      scale_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>((double*)ctx->coords.p, nc, alpha);
      ctx->tim.launches += 1;
    }
  }
  return GSM_OK;
}

// Phase 2: the grid bins over ctx->coords / ctx->values (KNearestSearch construction, models.jl:155).
int gsm_finish_samples(gsm_ctx* ctx, int32_t dim, int64_t n) {
  ctx->dim = dim;
  ctx->n = n;
  int rc = gsm_build_bins(ctx);
  if (rc != GSM_OK) ctx->n = 0;
  return rc;
}

extern "C" {

static int set_samples_impl(gsm_ctx* ctx, const double* coords, int32_t dim, int64_t n, const double* values,
                            bool scaled, double scale_floor, double* out_alpha, int64_t* out_nan) {
  if (!ctx) return GSM_ERR_INVALID;
  DeviceGuard g(ctx->device);
  GsmRange nv("gsm_set_samples");
  ctx->tim = gsm_timings{};
  EventTimer et;
  cudaEvent_t e0 = et.mark(ctx->stream);
  {
    GsmRange nv1("upload + _scale");
    GSM_CHECK(gsm_upload_samples(ctx, coords, dim, n, values, scaled, scale_floor, out_alpha, out_nan));
  }
  cudaEvent_t e1 = et.mark(ctx->stream);
  GsmRange nv2("bins");
  GSM_CHECK(gsm_finish_samples(ctx, dim, n));
  cudaEvent_t e2 = et.mark(ctx->stream);
  GSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->tim.h2d_ms = EventTimer::ms(e0, e1);
  ctx->tim.bin_ms = EventTimer::ms(e1, e2);
  ctx->tim.total_ms = EventTimer::ms(e0, e2);
  return GSM_OK;
}

int gsm_set_samples(gsm_ctx* ctx, const double* coords, int32_t dim, int64_t n, const double* values) {
  return set_samples_impl(ctx, coords, dim, n, values, false, 1.0, nullptr, nullptr);
}

int gsm_set_samples_scaled(gsm_ctx* ctx, const double* coords, int32_t dim, int64_t n, const double* values,
                           double scale_floor, double* out_alpha, int64_t* out_nan) {
  if (!(scale_floor > 0.0)) {
    gsm_set_error(ctx, "gsm_set_samples_scaled: scale_floor must be positive");
    return GSM_ERR_INVALID;
  }
  return set_samples_impl(ctx, coords, dim, n, values, true, scale_floor, out_alpha, out_nan);
}

// Shared driver of the neighbourhood-based entry points.
//   mode 0: kNN only, 1: local Kriging, 2: local IDW
static int run_local_body(gsm_ctx* ctx, int mode, const gsm_variogram* vario, const gsm_model* model, double exponent,
                          const gsm_targets* targets, const gsm_neighbor_opts* opts, double* out_mean, double* out_var,
                          uint8_t* out_status, int32_t* out_nbr, int32_t* out_n);
static int run_local(gsm_ctx* ctx, int mode, const gsm_variogram* vario, const gsm_model* model, double exponent,
                     const gsm_targets* targets, const gsm_neighbor_opts* opts, double* out_mean, double* out_var,
                     uint8_t* out_status, int32_t* out_nbr, int32_t* out_n) {
  const int rc = run_local_body(ctx, mode, vario, model, exponent, targets, opts, out_mean, out_var, out_status, out_nbr,
                                out_n);
  if (rc != GSM_OK && ctx) {
    // an early return may leave asynchronous downloads into caller memory in flight: drain both streams so no
    // buffer (the caller's, or a pooled pinned one) is recycled under a running copy
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaGetLastError();
  }
  return rc;
}
static int run_local_body(gsm_ctx* ctx, int mode, const gsm_variogram* vario, const gsm_model* model, double exponent,
                          const gsm_targets* targets, const gsm_neighbor_opts* opts, double* out_mean, double* out_var,
                          uint8_t* out_status, int32_t* out_nbr, int32_t* out_n) {
  if (!ctx) return GSM_ERR_INVALID;
  if (ctx->n <= 0) {
    gsm_set_error(ctx, "no samples: call gsm_set_samples first");
    return GSM_ERR_NO_SAMPLES;
  }
  if (!opts) {
    gsm_set_error(ctx, "neighbour options are required");
    return GSM_ERR_INVALID;
  }
  DeviceGuard g(ctx->device);
  GsmRange nv(mode == 0 ? "gsm_knn" : (mode == 1 ? "gsm_local_krige" : (mode == 2 ? "gsm_idw (local)" : "gsm_nn_local")));
  ctx->tim = gsm_timings{};
  VarioDev vd{};
  ModelDev md{};
  if (mode == 1) {
    GSM_CHECK(check_vario(ctx, vario));
    vd = make_vario(*vario);
    GSM_CHECK(make_model(ctx, model, ctx->dim, md));
    GSM_CHECK(gsm_block_prepare(ctx, vd));
    if (!out_mean) {
      gsm_set_error(ctx, "out_mean is required");
      return GSM_ERR_INVALID;
    }
  }
  // (IDW with missing values needs no branch: a NaN value makes the weighted sum NaN, which is the reference's
  //  `missing` from sum(i -> w_i z_i) over a column with missings, and an exact hit returns that sample's value
  //  whatever the others hold, idw.jl:59-73)
  int maxn, minn;
  clamp_neighbors(ctx->n, opts, maxn, minn);
  if (maxn > GSM_MAX_NEIGHBORS) {
    gsm_set_error(ctx, "maxneighbors (after clamping to the number of samples) exceeds GSM_MAX_NEIGHBORS");
    return GSM_ERR_UNSUPPORTED;
  }
  EventTimer et;
  cudaStream_t st = ctx->stream;
  cudaEvent_t e_begin = et.mark(st);
  TargetsDev td{};
  long long M = 0;
  GSM_CHECK(make_targets(ctx, targets, td, M));
  cudaEvent_t e_h2d = et.mark(st);

  OutBuf<double> omean, ovar;
  OutBuf<uint8_t> ostat;
  OutBuf<int32_t> onbr, on;
  GSM_CHECK(omean.init(ctx, out_mean, ctx->out_mean, M));
  GSM_CHECK(ovar.init(ctx, out_var, ctx->out_var, M));
  GSM_CHECK(ostat.init(ctx, out_status, ctx->out_status, M));
  GSM_CHECK(onbr.init(ctx, out_nbr, ctx->out_nbr, M * maxn));
  // Targets per kernel wave: CHUNK (bounds the neighbour scratch; smaller waves were measured: 2^18 hides more of
  // the last download of a short multi-GPU slab but costs the persistent kernels more than it saves, 9.7 -> 10.4 ms
  // per rank at 8 GPUs).  On grids a multiple of a pair of rows / planes, so that no wave boundary cuts a patch of
  // the solve kernels and the result does not depend on how the domain is chunked or sharded.
  long long chunk = CHUNK;
  if (td.kind == GSM_TARGETS_GRID && td.dim >= 2) {
    const long long unit = 2 * td.dims[0] * (td.dim >= 3 ? td.dims[1] : 1);
    if (unit <= chunk && td.first % unit == 0) chunk = chunk / unit * unit;
  }
  chunk = std::min<long long>(chunk, std::max<long long>(M, 1));
  // The download of the LAST wave is the only one no kernel hides: split a short tail wave off the end (2^17 targets,
  // patch-aligned like the waves) so that what is exposed is 2 MB instead of 18 MB (0.5 ms of a 9 ms slab at 8 GPUs).
  long long tail = 1LL << 17;
  if (td.kind == GSM_TARGETS_GRID && td.dim >= 2) {
    const long long unit = 2 * td.dims[0] * (td.dim >= 3 ? td.dims[1] : 1);
    if (unit <= tail && td.first % unit == 0) tail = tail / unit * unit;
  }
  std::vector<std::pair<long long, long long>> waves;   // (offset, count)
  for (long long off = 0; off < M; off += chunk) {
    const long long cnt_w = std::min(chunk, M - off);
    if (off + cnt_w >= M && cnt_w >= 4 * tail) {
      waves.emplace_back(off, cnt_w - tail);
      waves.emplace_back(off + cnt_w - tail, tail);
    } else {
      waves.emplace_back(off, cnt_w);
    }
  }
  GSM_CHECK(gsm_reserve(ctx, ctx->nbr_pos, (size_t)chunk * maxn * sizeof(int)));
  DevBuf& cntbuf = ctx->nbr_cnt;
  GSM_CHECK(gsm_reserve(ctx, cntbuf, (size_t)std::max<long long>(M, 1) * sizeof(int)));
  int* d_cnt_all = (int*)cntbuf.p;
  if (out_n && gsm_is_device_ptr(out_n)) d_cnt_all = out_n;

  GSM_CHECK(gsm_reserve(ctx, ctx->summ_dev, 4 * sizeof(unsigned long long)));
  unsigned long long* d_summ = (unsigned long long*)ctx->summ_dev.p;
  GSM_CUDA(ctx, cudaMemsetAsync(d_summ, 0, 4 * sizeof(unsigned long long), st));
  ctx->summ = gsm_summary{0, 0, 0, -1};
  double knn_ms = 0, solve_ms = 0, index_ms = 0;
  for (auto e : ctx->idx_marks) cudaEventDestroy(e);
  ctx->idx_marks.clear();
  std::vector<cudaEvent_t> marks;
  std::vector<cudaEvent_t> chunk_done;
  for (const auto& wv : waves) {
    const long long off = wv.first;
    TargetsDev tc = td;
    tc.first = td.first + off;
    tc.count = wv.second;
    if (td.kind == GSM_TARGETS_POINTS) {
      tc.first = 0;
      tc.points = td.points + off * td.dim;
    }
    int* d_pos = (int*)ctx->nbr_pos.p;
    int* d_cnt = d_cnt_all + off;
    GsmRange nvw("wave: knn + solve + download");
    cudaEvent_t a = et.mark(st);
    const bool fused_idw = mode == 2 && !onbr.dev;   // local IDW without neighbour lists: predict inside the search
    if (fused_idw)
      GSM_CHECK(gsm_launch_knn_idw(ctx, tc, maxn, opts->radius, exponent, minn, d_cnt, omean.dev + off,
                                   ostat.dev ? ostat.dev + off : nullptr));
    else
      GSM_CHECK(gsm_launch_knn(ctx, tc, maxn, opts->radius, d_pos, d_cnt));
    cudaEvent_t bq = et.mark(st);
    if (onbr.dev) {
      GSM_CHECK(gsm_launch_sort_lists(ctx, tc, maxn, d_pos, d_cnt));   // lists are promised ascending by (d2, row)
      GSM_CHECK(gsm_launch_nbr_to_rows(ctx, d_pos, tc.count * maxn, onbr.dev + off * maxn));
    }
    if (mode == 1) {
      GSM_CHECK(gsm_launch_local_krige(ctx, vd, md, tc, maxn, minn, d_pos, d_cnt, omean.dev + off,
                                       ovar.dev ? ovar.dev + off : nullptr, ostat.dev ? ostat.dev + off : nullptr));
    } else if (mode == 2 && !fused_idw) {
      GSM_CHECK(gsm_launch_idw_local(ctx, exponent, tc, maxn, minn, d_pos, d_cnt, omean.dev + off,
                                     ostat.dev ? ostat.dev + off : nullptr));
    }
    if (mode == 3)
      GSM_CHECK(gsm_launch_nn_local(ctx, tc, maxn, minn, d_pos, d_cnt, omean.dev + off, ostat.dev ? ostat.dev + off : nullptr));
    cudaEvent_t c = et.mark(st);
    if (mode != 0 && ostat.dev) {   // (after the solve mark: not part of solve_ms)
      summary_kernel<<<(unsigned)((tc.count + 255) / 256), 256, 0, st>>>(ostat.dev + off, ovar.dev ? ovar.dev + off : nullptr,
                                                                       tc.count, d_summ);
      ctx->tim.launches++;
    }
    marks.push_back(a);
    marks.push_back(bq);
    marks.push_back(c);
    // overlap this chunk's download with the next chunk's kernels
    GSM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, c, 0));
    GSM_CHECK(omean.download(ctx, off, tc.count, ctx->copy_stream));
    GSM_CHECK(ovar.download(ctx, off, tc.count, ctx->copy_stream));
    GSM_CHECK(ostat.download(ctx, off, tc.count, ctx->copy_stream));
    GSM_CHECK(onbr.download(ctx, off * maxn, tc.count * maxn, ctx->copy_stream));
    if (out_n && d_cnt_all != out_n)
      GSM_CUDA(ctx, cudaMemcpyAsync(out_n + off, d_cnt, (size_t)tc.count * sizeof(int), cudaMemcpyDeviceToHost,
                                    ctx->copy_stream));
  }
  cudaEvent_t e_k = et.mark(st);
  cudaEvent_t e_copy = et.mark(ctx->copy_stream);
  GSM_CUDA(ctx, cudaStreamSynchronize(st));
  GSM_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  for (size_t i = 0; i + 3 <= marks.size(); i += 3) {
    knn_ms += EventTimer::ms(marks[i], marks[i + 1]);
    solve_ms += EventTimer::ms(marks[i + 1], marks[i + 2]);
  }
  for (size_t i = 0; i + 2 <= ctx->idx_marks.size(); i += 2) index_ms += EventTimer::ms(ctx->idx_marks[i], ctx->idx_marks[i + 1]);
  for (auto e : ctx->idx_marks) cudaEventDestroy(e);
  ctx->idx_marks.clear();
  if (mode != 0 && ostat.dev) {
    unsigned long long hs[4] = {0, 0, 0, 0};
    GSM_CUDA(ctx, cudaMemcpy(hs, d_summ, sizeof(hs), cudaMemcpyDeviceToHost));
    ctx->summ = gsm_summary{(int64_t)hs[0], (int64_t)hs[1], (int64_t)hs[2], (int64_t)M};
  }
  ctx->tim.h2d_ms = EventTimer::ms(e_begin, e_h2d);
  ctx->tim.knn_ms = knn_ms;
  ctx->tim.solve_ms = solve_ms;
  ctx->tim.index_ms = index_ms;
  ctx->tim.bin_ms = 0;
  double t_k = EventTimer::ms(e_begin, e_k);
  double t_c = EventTimer::ms(e_begin, e_copy);
  ctx->tim.total_ms = std::max(t_k, t_c);
  ctx->tim.d2h_ms = std::max(0.0, t_c - t_k);  // download time NOT hidden behind kernels
  return GSM_OK;
}

int gsm_knn(gsm_ctx* ctx, const gsm_targets* targets, const gsm_neighbor_opts* opts, int32_t* out_idx,
            int32_t* out_n) {
  return run_local(ctx, 0, nullptr, nullptr, 0.0, targets, opts, nullptr, nullptr, nullptr, out_idx, out_n);
}

int gsm_local_krige(gsm_ctx* ctx, const gsm_variogram* vario, const gsm_model* model, const gsm_targets* targets,
                    const gsm_neighbor_opts* opts, double* out_mean, double* out_var, uint8_t* out_status,
                    int32_t* out_nbr) {
  return run_local(ctx, 1, vario, model, 0.0, targets, opts, out_mean, out_var, out_status, out_nbr, nullptr);
}

int gsm_nn(gsm_ctx* ctx, const gsm_targets* targets, double* out_value) {
  if (!ctx) return GSM_ERR_INVALID;
  if (!out_value) {
    gsm_set_error(ctx, "out_value is required");
    return GSM_ERR_INVALID;
  }
  if (ctx->n <= 0) {
    gsm_set_error(ctx, "no samples: call gsm_set_samples first");
    return GSM_ERR_NO_SAMPLES;
  }
  if (ctx->n_missing > 0) {
    gsm_set_error(ctx, "NN with missing values is not claimed by the GPU path");
    return GSM_ERR_UNSUPPORTED;
  }
  DeviceGuard g(ctx->device);
  ctx->tim = gsm_timings{};
  EventTimer et;
  cudaStream_t st = ctx->stream;
  cudaEvent_t e0 = et.mark(st);
  TargetsDev td{};
  long long M = 0;
  GSM_CHECK(make_targets(ctx, targets, td, M));
  cudaEvent_t e1 = et.mark(st);
  OutBuf<double> oval;
  GSM_CHECK(oval.init(ctx, out_value, ctx->out_mean, M));
  GSM_CHECK(gsm_reserve(ctx, ctx->nbr_cnt, (size_t)std::max<long long>(M, 1) * sizeof(int)));
  GSM_CHECK(gsm_launch_knn_nearest(ctx, td, (int*)ctx->nbr_cnt.p, oval.dev));
  cudaEvent_t e2 = et.mark(st);
  GSM_CHECK(oval.download(ctx, 0, M, st));
  cudaEvent_t e3 = et.mark(st);
  GSM_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->tim.h2d_ms = EventTimer::ms(e0, e1);
  ctx->tim.knn_ms = EventTimer::ms(e1, e2);
  ctx->tim.d2h_ms = EventTimer::ms(e2, e3);
  ctx->tim.total_ms = EventTimer::ms(e0, e3);
  return GSM_OK;
}

int gsm_nn_local(gsm_ctx* ctx, const gsm_targets* targets, const gsm_neighbor_opts* opts, double* out_value,
                 uint8_t* out_status) {
  if (ctx && !out_value) {
    gsm_set_error(ctx, "out_value is required");
    return GSM_ERR_INVALID;
  }
  return run_local(ctx, 3, nullptr, nullptr, 0.0, targets, opts, out_value, nullptr, out_status, nullptr, nullptr);
}

int gsm_idw(gsm_ctx* ctx, double exponent, const gsm_targets* targets, const gsm_neighbor_opts* opts,
            double* out_mean, uint8_t* out_status) {
  if (!ctx) return GSM_ERR_INVALID;
  if (!out_mean) {
    gsm_set_error(ctx, "out_mean is required");
    return GSM_ERR_INVALID;
  }
  if (!(exponent > 0.0) || !std::isfinite(exponent)) {
    gsm_set_error(ctx, "IDW exponent must be positive and finite");
    return GSM_ERR_INVALID;
  }
  if (opts && opts->maxneighbors != 0)
    return run_local(ctx, 2, nullptr, nullptr, exponent, targets, opts, out_mean, nullptr, out_status, nullptr, nullptr);
  if (ctx->n <= 0) {
    gsm_set_error(ctx, "no samples: call gsm_set_samples first");
    return GSM_ERR_NO_SAMPLES;
  }
  DeviceGuard g(ctx->device);
  ctx->tim = gsm_timings{};
  EventTimer et;
  cudaStream_t st = ctx->stream;
  cudaEvent_t e0 = et.mark(st);
  TargetsDev td{};
  long long M = 0;
  GSM_CHECK(make_targets(ctx, targets, td, M));
  cudaEvent_t e1 = et.mark(st);
  OutBuf<double> omean;
  OutBuf<uint8_t> ostat;
  GSM_CHECK(omean.init(ctx, out_mean, ctx->out_mean, M));
  GSM_CHECK(ostat.init(ctx, out_status, ctx->out_status, M));
  GSM_CHECK(gsm_launch_idw_global(ctx, exponent, td, omean.dev));
  if (ostat.dev) GSM_CUDA(ctx, cudaMemsetAsync(ostat.dev, 0, (size_t)M, st));
  cudaEvent_t e2 = et.mark(st);
  GSM_CHECK(omean.download(ctx, 0, M, st));
  GSM_CHECK(ostat.download(ctx, 0, M, st));
  cudaEvent_t e3 = et.mark(st);
  GSM_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->tim.h2d_ms = EventTimer::ms(e0, e1);
  ctx->tim.solve_ms = EventTimer::ms(e1, e2);
  ctx->tim.d2h_ms = EventTimer::ms(e2, e3);
  ctx->tim.total_ms = EventTimer::ms(e0, e3);
  return GSM_OK;
}

int gsm_measure_fp64_peak(gsm_ctx* ctx, double* out_dfma_tflops, double* out_dmma_tflops) {
  if (!ctx) return GSM_ERR_INVALID;
  DeviceGuard g(ctx->device);
  return gsm_fp64_peak(ctx, out_dfma_tflops, out_dmma_tflops);
}

}  // extern "C"
