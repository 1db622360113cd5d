// Multi-GPU entry points of the C ABI: ONE caller (one Julia process, one `ccall`) drives all GPUs of a box.
//
// Reference analogue: `_predictionthread!` (src/models.jl:236-247) splits the target index into contiguous chunks,
// one task and one private copy of the fitted state per chunk, no communication between chunks.  Here a chunk is a
// GPU: the samples are uploaded ONCE to the first device and replicated with one NCCL broadcast over NVLink
// (ncclCommInitAll at gsm_multi_create, cached in the handle); every device builds its own search bins; a host
// thread per device predicts its contiguous slab of the domain and DMA-writes it straight into the caller's
// (pinned) output slice.  No gather, no collective on the data path.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a process that already carries an NCCL (torch) keeps exactly
// one copy, and the single-GPU entry points do not depend on NCCL at all.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "gsm_internal.cuh"

namespace {

// the handful of NCCL declarations used (nccl.h: ncclCommInitAll :169, ncclBroadcast :379); kept local so that the
// library builds on a box without NCCL headers
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclUint8 = 1, kNcclFloat64 = 8;
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err) {
    if (handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (handle) break;
    }
    if (!handle) {
      err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
      return false;
    }
    auto sym = [&](const char* s) { return dlsym(handle, s); };
    CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    Broadcast = (decltype(Broadcast))sym("ncclBroadcast");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !Broadcast || !GroupStart || !GroupEnd || !GetErrorString) {
      err = "libnccl.so.2 lacks a required symbol";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;

struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) {
    cudaGetDevice(&prev);
    cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct gsm_multi {
  std::vector<gsm_ctx*> ctx;
  std::vector<ncclComm_t> comm;     // empty when ndev == 1
  std::vector<int> devices;
  std::string err;
  gsm_multi_timings tim{};
  gsm_summary summ{0, 0, 0, -1};
  int32_t dim = 0;
  int64_t n = 0;
};

static std::string g_multi_create_error;

namespace {

int fail(gsm_multi* mg, int code, const std::string& msg) {
  if (mg) mg->err = msg;
  else g_multi_create_error = msg;
  return code;
}

// contiguous slab `i` of `ndev` over [first, first + count): sizes differ by at most one unit, where a unit is a
// pair of grid rows (2-D) / planes (3-D) so that no slab boundary cuts a 4 x 2 / 2 x 2 x 2 patch of the solve kernels
void slab_of(const gsm_targets* t, long long first, long long count, int i, int ndev, long long& f, long long& c) {
  long long unit = 1;
  if (t->kind == GSM_TARGETS_GRID && t->dim >= 2) {
    unit = 2 * t->dims[0];
    if (t->dim >= 3) unit *= t->dims[1];
    if (first % unit != 0 || count / unit < 4LL * ndev) unit = 1;   // small or unaligned ranges: plain index chunks
  }
  const long long units = count / unit, base = units / ndev, rem = units % ndev;
  const long long u0 = i * base + std::min<long long>(i, rem), nu = base + (i < rem ? 1 : 0);
  f = first + u0 * unit;
  c = nu * unit;
  if (i == ndev - 1) c = first + count - f;   // the tail (count not a multiple of the unit) goes to the last slab
}

long long domain_count(const gsm_targets* t, long long& first) {
  long long total = 1;
  if (t->kind == GSM_TARGETS_GRID)
    for (int d = 0; d < t->dim; ++d) total *= t->dims[d];
  else
    total = t->npoints;
  first = t->first;
  return t->count < 0 ? total - first : t->count;
}

// run fn(i) on one host thread per device; returns the first non-OK code
template <class F>
int per_device(gsm_multi* mg, F fn) {
  const int nd = (int)mg->ctx.size();
  std::vector<int> rc(nd, GSM_OK);
  if (nd == 1) {
    rc[0] = fn(0);
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < nd; ++i) th.emplace_back([&, i]() { rc[i] = fn(i); });
    for (auto& t : th) t.join();
  }
  for (int i = 0; i < nd; ++i)
    if (rc[i] != GSM_OK) {
      mg->err = "device " + std::to_string(mg->devices[i]) + ": " + gsm_last_error(mg->ctx[i]);
      return rc[i];
    }
  return GSM_OK;
}

void merge_summaries(gsm_multi* mg) {
  gsm_summary s{0, 0, 0, 0};
  bool known = true;
  for (gsm_ctx* c : mg->ctx) {
    gsm_summary q{};
    gsm_get_summary(c, &q);
    if (q.total < 0) known = false;
    s.missing += q.missing;
    s.singular += q.singular;
    s.nonpositive_var += q.nonpositive_var;
    s.total += std::max<int64_t>(q.total, 0);
  }
  if (!known) s.total = -1;
  mg->summ = s;
}

int set_samples_multi(gsm_multi* mg, const double* coords, int32_t dim, int64_t n, const double* values, bool scaled,
                      double scale_floor, double* out_alpha, int64_t* out_nan) {
  if (!mg) return GSM_ERR_INVALID;
  GsmRange nv("gsm_multi_set_samples (upload + ncclBroadcast + bins)");
  const int nd = (int)mg->ctx.size();
  mg->tim = gsm_multi_timings{};
  const double t0 = now_ms();
  gsm_ctx* c0 = mg->ctx[0];
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  {
    DevGuard g(c0->device);
    for (auto& e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], c0->stream);
    int rc = gsm_upload_samples(c0, coords, dim, n, values, scaled, scale_floor, out_alpha, out_nan);
    if (rc != GSM_OK) return fail(mg, rc, std::string("device ") + std::to_string(c0->device) + ": " + gsm_last_error(c0));
    cudaEventRecord(ev[1], c0->stream);
  }
  if (nd > 1) {
    for (int i = 1; i < nd; ++i) {
      DevGuard g(mg->ctx[i]->device);
      int rc = gsm_reserve(mg->ctx[i], mg->ctx[i]->coords, (size_t)n * dim * sizeof(double));
      if (rc == GSM_OK) rc = gsm_reserve(mg->ctx[i], mg->ctx[i]->values, (size_t)n * sizeof(double));
      if (rc != GSM_OK) return fail(mg, rc, gsm_last_error(mg->ctx[i]));
    }
    // ONE collective: the scaled coordinates and the values, root = the first device, every rank on its own stream
    ncclResult_t r = g_nccl.GroupStart();
    for (int i = 0; i < nd && r == 0; ++i) {
      gsm_ctx* c = mg->ctx[i];
      r = g_nccl.Broadcast(c0->coords.p, c->coords.p, (size_t)n * dim, kNcclFloat64, 0, mg->comm[i], c->stream);
      if (r == 0) r = g_nccl.Broadcast(c0->values.p, c->values.p, (size_t)n, kNcclFloat64, 0, mg->comm[i], c->stream);
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r != 0 || r2 != 0) return fail(mg, GSM_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(r != 0 ? r : r2));
  }
  {
    DevGuard g(c0->device);
    cudaEventRecord(ev[2], c0->stream);
  }
  for (gsm_ctx* c : mg->ctx) c->n_missing = c0->n_missing;
  int rc = per_device(mg, [&](int i) -> int {
    gsm_ctx* c = mg->ctx[i];
    DevGuard g(c->device);
    c->tim = gsm_timings{};
    int r = gsm_finish_samples(c, dim, n);
    if (r != GSM_OK) return r;
    return (int)(cudaStreamSynchronize(c->stream) == cudaSuccess ? GSM_OK : GSM_ERR_CUDA);
  });
  {
    DevGuard g(c0->device);
    float a = 0.f, b = 0.f;
    cudaEventSynchronize(ev[2]);
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[1], ev[2]);
    mg->tim.upload_ms = a;
    mg->tim.bcast_ms = nd > 1 ? b : 0.0;
    for (auto& e : ev) cudaEventDestroy(e);
  }
  mg->tim.wall_ms = now_ms() - t0;
  mg->tim.bins_ms = mg->tim.wall_ms - mg->tim.upload_ms - mg->tim.bcast_ms;
  if (rc != GSM_OK) return rc;
  mg->dim = dim;
  mg->n = n;
  return GSM_OK;
}

}  // namespace

extern "C" {

int gsm_multi_create(const int32_t* devices, int32_t ndev, gsm_multi** out) {
  if (!out) return GSM_ERR_INVALID;
  *out = nullptr;
  if (!devices || ndev < 1 || ndev > 64) return fail(nullptr, GSM_ERR_INVALID, "gsm_multi_create: need 1..64 device ordinals");
  for (int i = 0; i < ndev; ++i)
    for (int j = 0; j < i; ++j)
      if (devices[i] == devices[j]) return fail(nullptr, GSM_ERR_INVALID, "gsm_multi_create: duplicate device ordinal");
  gsm_multi* mg = new gsm_multi();
  for (int i = 0; i < ndev; ++i) {
    gsm_ctx* c = nullptr;
    int rc = gsm_create(devices[i], &c);
    if (rc != GSM_OK) {
      g_multi_create_error = gsm_last_error(nullptr);
      gsm_multi_destroy(mg);
      return rc;
    }
    mg->ctx.push_back(c);
    mg->devices.push_back(devices[i]);
  }
  if (ndev > 1) {
    std::string err;
    if (!g_nccl.load(err)) {
      gsm_multi_destroy(mg);
      return fail(nullptr, GSM_ERR_UNSUPPORTED, err);
    }
    mg->comm.assign(ndev, nullptr);
    ncclResult_t r = g_nccl.CommInitAll(mg->comm.data(), ndev, mg->devices.data());
    if (r != 0) {
      mg->comm.clear();
      gsm_multi_destroy(mg);
      return fail(nullptr, GSM_ERR_CUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
    }
    // warm the communicator: the first collective on a fresh communicator sets up its channels (tens of ms)
    std::vector<void*> tmp(ndev, nullptr);
    for (int i = 0; i < ndev; ++i) {
      DevGuard g(mg->devices[i]);
      cudaMalloc(&tmp[i], 1 << 20);
    }
    g_nccl.GroupStart();
    for (int i = 0; i < ndev; ++i) g_nccl.Broadcast(tmp[0], tmp[i], 1 << 20, kNcclUint8, 0, mg->comm[i], mg->ctx[i]->stream);
    g_nccl.GroupEnd();
    for (int i = 0; i < ndev; ++i) {
      DevGuard g(mg->devices[i]);
      cudaStreamSynchronize(mg->ctx[i]->stream);
      cudaFree(tmp[i]);
    }
  }
  *out = mg;
  return GSM_OK;
}

int gsm_multi_destroy(gsm_multi* mg) {
  if (!mg) return GSM_OK;
  for (size_t i = 0; i < mg->comm.size(); ++i)
    if (mg->comm[i]) g_nccl.CommDestroy(mg->comm[i]);
  for (gsm_ctx* c : mg->ctx) gsm_destroy(c);
  delete mg;
  return GSM_OK;
}

const char* gsm_multi_last_error(const gsm_multi* mg) { return mg ? mg->err.c_str() : g_multi_create_error.c_str(); }

int gsm_multi_device_count(const gsm_multi* mg) { return mg ? (int)mg->ctx.size() : 0; }

gsm_ctx* gsm_multi_context(gsm_multi* mg, int32_t i) {
  return (mg && i >= 0 && i < (int)mg->ctx.size()) ? mg->ctx[i] : nullptr;
}

int gsm_multi_get_timings(const gsm_multi* mg, gsm_multi_timings* out) {
  if (!mg || !out) return GSM_ERR_INVALID;
  *out = mg->tim;
  return GSM_OK;
}

int gsm_multi_get_summary(const gsm_multi* mg, gsm_summary* out) {
  if (!mg || !out) return GSM_ERR_INVALID;
  *out = mg->summ;
  return GSM_OK;
}

int gsm_multi_set_samples(gsm_multi* mg, const double* coords, int32_t dim, int64_t n, const double* values) {
  return set_samples_multi(mg, coords, dim, n, values, false, 1.0, nullptr, nullptr);
}

int gsm_multi_set_samples_scaled(gsm_multi* mg, const double* coords, int32_t dim, int64_t n, const double* values,
                                 double scale_floor, double* out_alpha, int64_t* out_nan) {
  if (!mg) return GSM_ERR_INVALID;
  if (!(scale_floor > 0.0)) return fail(mg, GSM_ERR_INVALID, "gsm_multi_set_samples_scaled: scale_floor must be positive");
  return set_samples_multi(mg, coords, dim, n, values, true, scale_floor, out_alpha, out_nan);
}

int gsm_multi_slab(const gsm_multi* mg, const gsm_targets* targets, int32_t i, int64_t* out_first, int64_t* out_count) {
  if (!mg || !targets || i < 0 || i >= (int)mg->ctx.size() || !out_first || !out_count) return GSM_ERR_INVALID;
  long long first = 0;
  const long long count = domain_count(targets, first);
  long long f = 0, c = 0;
  slab_of(targets, first, count, i, (int)mg->ctx.size(), f, c);
  *out_first = f;
  *out_count = c;
  return GSM_OK;
}

static int multi_predict(gsm_multi* mg, const gsm_targets* targets, bool outputs_ok,
                         const std::function<int(int, gsm_ctx*, const gsm_targets*, long long)>& call) {
  if (!mg || !targets) return GSM_ERR_INVALID;
  GsmRange nv("gsm_multi predict (slab per device)");
  if (mg->n <= 0) return fail(mg, GSM_ERR_NO_SAMPLES, "no samples: call gsm_multi_set_samples first");
  if (!outputs_ok) return fail(mg, GSM_ERR_INVALID, "multi-GPU outputs must be HOST buffers (pinned for direct DMA)");
  const int nd = (int)mg->ctx.size();
  long long first = 0;
  const long long count = domain_count(targets, first);
  const double t0 = now_ms();
  int rc = per_device(mg, [&](int i) -> int {
    long long f = 0, c = 0;
    slab_of(targets, first, count, i, nd, f, c);
    if (c <= 0) {
      mg->ctx[i]->tim = gsm_timings{};
      mg->ctx[i]->summ = gsm_summary{0, 0, 0, 0};
      return (int)GSM_OK;
    }
    gsm_targets t = *targets;
    t.first = f;
    t.count = c;
    return call(i, mg->ctx[i], &t, f - first);
  });
  mg->tim.wall_ms = now_ms() - t0;
  double mx = 0.0;
  for (gsm_ctx* c : mg->ctx) mx = std::max(mx, c->tim.total_ms);
  mg->tim.predict_ms = mx;
  merge_summaries(mg);
  return rc;
}

int gsm_multi_local_krige(gsm_multi* mg, const gsm_variogram* vario, const gsm_model* model, const gsm_targets* targets,
                          const gsm_neighbor_opts* opts, double* out_mean, double* out_var, uint8_t* out_status) {
  const bool host = mg && (mg->ctx.size() == 1 || (!gsm_is_device_ptr(out_mean) && !gsm_is_device_ptr(out_var) &&
                                                  !gsm_is_device_ptr(out_status)));
  return multi_predict(mg, targets, host, [&](int, gsm_ctx* c, const gsm_targets* t, long long off) {
    return gsm_local_krige(c, vario, model, t, opts, out_mean ? out_mean + off : nullptr, out_var ? out_var + off : nullptr,
                           out_status ? out_status + off : nullptr, nullptr);
  });
}

int gsm_multi_idw(gsm_multi* mg, double exponent, const gsm_targets* targets, const gsm_neighbor_opts* opts,
                  double* out_mean, uint8_t* out_status) {
  const bool host = mg && (mg->ctx.size() == 1 || (!gsm_is_device_ptr(out_mean) && !gsm_is_device_ptr(out_status)));
  return multi_predict(mg, targets, host, [&](int, gsm_ctx* c, const gsm_targets* t, long long off) {
    return gsm_idw(c, exponent, t, opts, out_mean ? out_mean + off : nullptr, out_status ? out_status + off : nullptr);
  });
}

int gsm_multi_global_krige(gsm_multi* mg, const gsm_variogram* vario, const gsm_model* model, const gsm_targets* targets,
                           double* out_mean, double* out_var, int32_t* out_fit_status) {
  const bool host = mg && (mg->ctx.size() == 1 || (!gsm_is_device_ptr(out_mean) && !gsm_is_device_ptr(out_var)));
  std::vector<int> st(mg ? mg->ctx.size() : 0, 1);
  // the (N + ncon)^2 factor is recomputed on every device (cheaper than moving it)
  int rc = multi_predict(mg, targets, host, [&](int i, gsm_ctx* c, const gsm_targets* t, long long off) {
    gsm_fit* f = nullptr;
    int r = gsm_global_krige_fit(c, vario, model, &f);
    if (r != GSM_OK) return r;
    st[i] = gsm_global_krige_status(f);
    r = gsm_global_krige_predict(f, t, out_mean ? out_mean + off : nullptr, out_var ? out_var + off : nullptr);
    gsm_global_krige_free(f);
    return r;
  });
  if (out_fit_status) {
    *out_fit_status = 1;
    for (int s : st) *out_fit_status &= s;
  }
  return rc;
}

}  // extern "C"
