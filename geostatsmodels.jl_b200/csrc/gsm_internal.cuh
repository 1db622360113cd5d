// Internal declarations shared by the kernels and the C-ABI glue of libgsm_b200.so.
// Nothing here is part of the public ABI (see include/gsm_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "gsm_b200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing: every CUDA call funnels through GSM_CUDA so failures surface as GSM_ERR_CUDA
// with a message retrievable through gsm_last_error (the C layer never throws).
// ---------------------------------------------------------------------------------------------
#define GSM_CUDA(ctx, call)                                                              \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      gsm_set_error((ctx), std::string(#call) + ": " + cudaGetErrorString(e_) + " at " + \
                               __FILE__ + ":" + std::to_string(__LINE__));               \
      return GSM_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define GSM_CHECK(expr)        \
  do {                         \
    int rc_ = (expr);          \
    if (rc_ != GSM_OK) return rc_; \
  } while (0)

void gsm_set_error(gsm_ctx* ctx, const std::string& msg);

// NVTX ranges around the phases of a call (SURVEY section 5 "tracing"): header-only NVTX3, a no-op unless a profiler
// (nsys / ncu --nvtx) has injected itself into the process.
#include <nvtx3/nvToolsExt.h>
struct GsmRange {
  explicit GsmRange(const char* name) { nvtxRangePushA(name); }
  ~GsmRange() { nvtxRangePop(); }
  GsmRange(const GsmRange&) = delete;
  GsmRange& operator=(const GsmRange&) = delete;
};

// A pivot of the covariance block at or below GSM_PIVOT_RTOL * sill, or a pivot of the drift Schur complement at
// or below GSM_SCHUR_RTOL times its own diagonal entry, marks the system singular (coincident samples, fewer points
// than drift terms).  The reference's status(fit) = issuccess(bunchkaufman!(...)) (krig.jl:152) catches these
// cases through exact zero pivots; here rounding can leave a pivot of a few ulps, which must not turn into an
// unflagged garbage prediction.
#define GSM_PIVOT_RTOL 1e-12
#define GSM_SCHUR_RTOL 1e-13

// ---------------------------------------------------------------------------------------------
// device-side POD parameter blocks (passed by value as kernel arguments)
// ---------------------------------------------------------------------------------------------
struct VarioStruct {   // one structure of a general (composite / anisotropic / new-family) variogram
  int kind;
  double cs, nadd;       // partial sill and the nugget added for h > 0
  double m[9];           // metric / range: t = |m (x - y)|
};
struct VarioDev {
  int kind;
  double sill, nugget, range, inv_range;
  double cs;    // partial sill: sill - nadd
  double nadd;  // nugget added for h > 0 (Gaussian: nugget + 1e-6)
  int general;  // 1: evaluate through st[] (gsm_cov_vec); the fast kernels only take general == 0
  int nstruct;
  VarioStruct st[GSM_MAX_STRUCTURES];
};

// The ONE place where a variogram descriptor becomes device constants (include/gsm_b200.h lists the formulas).
// GaussianVariogram replaces the nugget by n' = nugget + 1e-6 BEFORE forming the partial sill:
// gamma(h) = (sill - n')(1 - exp(-3 t^2)) + (h > 0) n'  (oracle GAUSSIAN_NUGGET_EPS).
inline VarioDev gsm_make_vario(const gsm_variogram& v) {
  VarioDev d{};
  d.kind = v.kind;
  d.sill = v.sill;
  d.nugget = v.nugget;
  d.range = v.range;
  d.inv_range = 1.0 / v.range;
  d.nadd = v.kind == GSM_VARIO_GAUSSIAN ? v.nugget + 1e-6 : v.nugget;
  d.cs = v.sill - d.nadd;
  d.general = (v.nstruct > 0 || v.kind > GSM_VARIO_EXPONENTIAL) ? 1 : 0;
  d.nstruct = 0;
  if (d.general) {
    const int ns = v.nstruct > 0 ? v.nstruct : 1;
    d.nstruct = ns;
    double total = 0.0;
    for (int i = 0; i < ns; ++i) {
      gsm_structure s{};
      if (v.nstruct > 0) {
        s = v.structs[i];
      } else {   // a single isotropic structure of one of the newer families
        s.kind = v.kind;
        s.sill = v.sill;
        s.nugget = v.nugget;
        s.range = v.range;
        s.metric[0] = s.metric[4] = s.metric[8] = 1.0;
      }
      VarioStruct& q = d.st[i];
      q.kind = s.kind;
      q.nadd = s.kind == GSM_VARIO_GAUSSIAN ? s.nugget + 1e-6 : s.nugget;
      q.cs = s.kind == GSM_VARIO_NUGGET ? 0.0 : s.sill - q.nadd;
      for (int e = 0; e < 9; ++e) q.m[e] = s.metric[e] / s.range;
      total += s.kind == GSM_VARIO_NUGGET ? s.nugget : s.sill;
    }
    d.sill = total;   // covariance form: total sill - gamma
  }
  return d;
}

struct TargetsDev {
  int kind;               // GSM_TARGETS_GRID / POINTS
  int dim;                // user dimension (1..3)
  double origin[3], spacing[3];
  long long dims[3];
  const double* points;   // device pointer, npoints x dim (AoS)
  long long first;        // first linear target index of this launch
  long long count;        // targets in this launch
};

struct ModelDev {
  int kind;
  int ndrift;
  double mean;
  int exps[GSM_MAX_DRIFT][3];
};

// Uniform grid bins over the sample bounding box (the KNearestSearch index).
struct BinsDev {
  double bbmin[3];
  double h;        // cell edge
  double inv_h;
  int nc[3];       // cells per dimension (1 in unused dimensions)
  const int* cell_start;      // ncells + 1, exclusive prefix of cell counts
  const double4* rec;         // samples sorted by cell: {x, y, z (0 if dim<3), value}
  const int* sidx;            // original (0-based) sample row of each sorted record
  int n;                      // number of samples
};

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

// gsm_set_option(ctx, "local_impl", ...): which local-Kriging kernel serves k <= 32 (default: dispatch by shape)
enum { GSM_IMPL_AUTO = 0, GSM_IMPL_SHR = 1, GSM_IMPL_PIPE = 2, GSM_IMPL_SIMT = 3, GSM_IMPL_PATCH = 4, GSM_IMPL_PATCH2 = 5 };

// Change of support (gsm_set_block_support): the points that stand for a target, relative to its centroid.
struct BlockDev {
  int nq;              // 0: point support
  const double* off;   // nq x 3 (z zero-padded in 2-D), device
  double c0;           // covzero: mean covariance over the nq^2 pairs (the total sill for a point, krig.jl:295-301)
};

struct gsm_ctx {
  int device = 0;
  // developer options (gsm_set_option; never read from the environment)
  int opt_local_impl = GSM_IMPL_AUTO;
  int opt_shr_smax = 3;            // cap on the shared block columns of the shared-factor kernel
  double opt_bin_occupancy = 0.0;  // samples per bin cell of the neighbour search (0: default)
  int opt_patch_nbmax = 0;         // block rows of the patch kernel's block store (0: default)
  int opt_knn_pad_smem = 0;        // extra dynamic shared memory per k-NN CTA (occupancy experiments)
  cudaStream_t stream = nullptr;       // compute
  cudaStream_t copy_stream = nullptr;  // D2H overlap
  std::string err;
  gsm_timings tim{};
  gsm_summary summ{};

  // samples
  int dim = 0;
  long long n = 0;
  DevBuf coords;      // n x dim doubles, caller order (AoS)
  DevBuf values;      // n doubles
  DevBuf rec;         // n double4, cell-sorted
  DevBuf sidx;        // n int, cell-sorted -> original row
  DevBuf cell_of;     // n int scratch
  DevBuf cell_start;  // ncells + 1 int
  DevBuf cell_fill;   // ncells int scratch
  DevBuf cub_tmp;
  DevBuf red;         // small reduction scratch
  BinsDev bins{};
  double sample_mean = 0.0;
  long long n_missing = 0;   // samples whose value is NaN (`missing`)

  // change of support
  int block_nq = 0;
  DevBuf block_off;        // block_nq x 3 doubles
  BlockDev block_cur{};    // descriptor of the running call (gsm_block_prepare)

  // per-call scratch
  DevBuf nbr_pos;     // chunk x k int : positions into rec
  DevBuf nbr_cnt;     // chunk int
  DevBuf job_idx;     // job descriptors of the shared-factor local-Kriging kernel
  DevBuf summ_dev;    // 4 x u64 status counters of the current call
  DevBuf glob_R, glob_acc;   // global Kriging prediction: covariance tile of a target batch, per-target sums
  DevBuf out_mean, out_var, out_status, out_nbr, tgt_points;

  // events for gsm_timings
  cudaEvent_t ev[8] = {};
  std::vector<cudaEvent_t> idx_marks;   // (begin, end) pairs around the job-index kernel of the current call
};

struct gsm_fit {
  gsm_ctx* ctx = nullptr;
  gsm_variogram vario{};
  gsm_model model{};
  int status = 0;
  long long n = 0;   // samples
  int ncon = 0;
  DevBuf Linv;       // n x n lower-triangular inverse Cholesky factor (column-major)
  DevBuf W;          // n x (1 + ncon) = Linv * [z, F]
  DevBuf schur;      // small host-side quantities mirrored to device
  std::vector<double> S;   // ncon x ncon Cholesky factor of the Schur complement (host)
  std::vector<double> h;   // ncon: V^T w  (host)
};

int gsm_reserve(gsm_ctx* ctx, DevBuf& b, size_t bytes);
bool gsm_is_device_ptr(const void* p);

// ---------------------------------------------------------------------------------------------
// kernel launchers (each returns GSM_OK / GSM_ERR_*; all enqueue on ctx->stream)
// ---------------------------------------------------------------------------------------------
int gsm_build_bins(gsm_ctx* ctx);
int gsm_check_vario(gsm_ctx* ctx, const gsm_variogram* v);
// fills ctx->block_cur for a call with variogram v (covzero of the block evaluated on the device); point support: nq = 0
int gsm_block_prepare(gsm_ctx* ctx, const VarioDev& v);
int gsm_upload_samples(gsm_ctx* ctx, const double* coords, int32_t dim, int64_t n, const double* values, bool scaled,
                       double scale_floor, double* out_alpha, int64_t* out_nan);
int gsm_finish_samples(gsm_ctx* ctx, int32_t dim, int64_t n);
int gsm_launch_knn(gsm_ctx* ctx, const TargetsDev& tg, int k, double radius, int* d_pos, int* d_cnt);
int gsm_launch_knn_idw(gsm_ctx* ctx, const TargetsDev& tg, int k, double radius, double exponent, int minneighbors,
                       int* d_cnt, double* d_mean, unsigned char* d_status);
int gsm_launch_knn_nearest(gsm_ctx* ctx, const TargetsDev& tg, int* d_cnt, double* d_value);
int gsm_launch_nn_local(gsm_ctx* ctx, const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                        double* d_val, unsigned char* d_status);
int gsm_launch_sort_lists(gsm_ctx* ctx, const TargetsDev& tg, int k, int* d_pos, const int* d_cnt);
int gsm_launch_nbr_to_rows(gsm_ctx* ctx, const int* d_pos, long long total, int* d_rows);
int gsm_launch_local_krige(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, const TargetsDev& tg, int k,
                           int minneighbors, const int* d_pos, const int* d_cnt, double* d_mean,
                           double* d_var, unsigned char* d_status);
int gsm_launch_local_krige_pipe(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered,
                                const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                                double* d_mean, double* d_var, unsigned char* d_status);
int gsm_launch_local_krige_shr(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered,
                               const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                               double* d_mean, double* d_var, unsigned char* d_status);
int gsm_launch_local_krige_patch(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered,
                                 const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                                 double* d_mean, double* d_var, unsigned char* d_status);
int gsm_launch_local_krige_mid(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered,
                               const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                               double* d_mean, double* d_var, unsigned char* d_status);
int gsm_launch_idw_global(gsm_ctx* ctx, double exponent, const TargetsDev& tg, double* d_mean);
int gsm_launch_idw_local(gsm_ctx* ctx, double exponent, const TargetsDev& tg, int k, int minneighbors,
                         const int* d_pos, const int* d_cnt, double* d_mean, unsigned char* d_status);
int gsm_fp64_peak(gsm_ctx* ctx, double* dfma, double* dmma);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Target coordinates of linear index `lin` (0-based over the WHOLE domain).  Grid: centroid
// (origin + i*spacing) + spacing/2 with every operation rounded once (no FMA) -- the declared
// arithmetic of include/gsm_b200.h and oracle Grid.centroids.
__device__ __forceinline__ void gsm_target_coords(const TargetsDev& tg, long long lin, double& x, double& y,
                                                  double& z) {
  if (tg.kind == GSM_TARGETS_GRID) {
    long long i0, i1, i2;
    if ((unsigned long long)lin < 0x80000000ULL && tg.dims[0] < 0x80000000LL && tg.dims[1] < 0x80000000LL) {
      // 32-bit index arithmetic (the common case; 64-bit division costs hundreds of instructions)
      const unsigned l = (unsigned)lin, d0 = (unsigned)tg.dims[0], d1 = (unsigned)tg.dims[1];
      const unsigned r = l / d0;
      i0 = l - r * d0;
      if (tg.dim > 2) {
        const unsigned q = r / d1;
        i1 = r - q * d1;
        i2 = q;
      } else {
        i1 = r;
        i2 = 0;
      }
    } else {
      i0 = lin % tg.dims[0];
      const long long r = lin / tg.dims[0];
      i1 = r % tg.dims[1];
      i2 = r / tg.dims[1];
    }
    x = __dadd_rn(__dadd_rn(tg.origin[0], __dmul_rn((double)i0, tg.spacing[0])), __dmul_rn(tg.spacing[0], 0.5));
    y = __dadd_rn(__dadd_rn(tg.origin[1], __dmul_rn((double)i1, tg.spacing[1])), __dmul_rn(tg.spacing[1], 0.5));
    z = __dadd_rn(__dadd_rn(tg.origin[2], __dmul_rn((double)i2, tg.spacing[2])), __dmul_rn(tg.spacing[2], 0.5));
  } else {
    const double* p = tg.points + lin * tg.dim;
    x = p[0];
    y = tg.dim > 1 ? p[1] : 0.0;
    z = tg.dim > 2 ? p[2] : 0.0;
  }
}

// Squared distance key of the neighbour search: ((dx*dx + dy*dy) + dz*dz), each op rounded once.
// Bit-identical to oracle sqdist_keys (a zero dz term adds exactly 0).
__device__ __forceinline__ double gsm_d2_key(double ax, double ay, double az, double bx, double by, double bz) {
  double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Covariance form sill - gamma(h) from a squared distance (krig.jl:165-179 on top of the variogram).
__device__ __forceinline__ double gsm_cov_d2(const VarioDev& v, double d2) {
  if (!(d2 > 0.0)) return v.sill;  // gamma(0) = 0
  double h = sqrt(d2);
  double t = h * v.inv_range;
  double g;
  if (v.kind == GSM_VARIO_SPHERICAL) {
    g = (h < v.range) ? v.cs * (1.5 * t - 0.5 * (t * t * t)) : v.cs;
  } else if (v.kind == GSM_VARIO_GAUSSIAN) {
    g = v.cs * (1.0 - exp(-3.0 * (t * t)));
  } else {
    g = v.cs * (1.0 - exp(-3.0 * t));
  }
  return v.sill - (g + v.nadd);
}

// General covariance sill - sum_i gamma_i(|m_i (dx, dy, dz)|): composite / anisotropic / newer families.
static __device__ __noinline__ double gsm_cov_general(const VarioDev& v, double dx, double dy, double dz) {
  double g = 0.0;
  bool zero = true;
  for (int i = 0; i < v.nstruct; ++i) {
    const VarioStruct& s = v.st[i];
    const double ux = s.m[0] * dx + s.m[1] * dy + s.m[2] * dz;
    const double uy = s.m[3] * dx + s.m[4] * dy + s.m[5] * dz;
    const double uz = s.m[6] * dx + s.m[7] * dy + s.m[8] * dz;
    const double t2 = ux * ux + uy * uy + uz * uz;
    if (!(t2 > 0.0)) continue;   // gamma_i(0) = 0
    zero = false;
    const double t = sqrt(t2);
    double f;
    switch (s.kind) {
      case GSM_VARIO_GAUSSIAN: f = 1.0 - exp(-3.0 * t2); break;
      case GSM_VARIO_SPHERICAL: f = t < 1.0 ? 1.5 * t - 0.5 * (t * t2) : 1.0; break;
      case GSM_VARIO_EXPONENTIAL: f = 1.0 - exp(-3.0 * t); break;
      case GSM_VARIO_CUBIC: f = t < 1.0 ? t2 * (7.0 + t * (-8.75 + t2 * (3.5 - 0.75 * t2))) : 1.0; break;
      case GSM_VARIO_PENTASPHERICAL: f = t < 1.0 ? t * (1.875 + t2 * (-1.25 + 0.375 * t2)) : 1.0; break;
      case GSM_VARIO_SINEHOLE: f = 1.0 - sin(M_PI * t) / (M_PI * t); break;
      case GSM_VARIO_CIRCULAR: f = t < 1.0 ? 1.0 - (2.0 / M_PI) * acos(t) + (2.0 * t / M_PI) * sqrt(1.0 - t2) : 1.0; break;
      default: f = 0.0; break;   // GSM_VARIO_NUGGET
    }
    g += s.cs * f + s.nadd;
  }
  (void)zero;
  return v.sill - g;
}
// Covariance between two points from their coordinate difference: the three isotropic single-structure families
// from the squared distance, everything else through gsm_cov_general.
__device__ __forceinline__ double gsm_cov_vec(const VarioDev& v, double dx, double dy, double dz) {
  if (v.general) return gsm_cov_general(v, dx, dy, dz);
  return gsm_cov_d2(v, dx * dx + dy * dy + dz * dz);
}

// Mean covariance between a sample and the points of a block target; (dx, dy, dz) = sample - centroid
// (GeoStatsFunctions' f(point, geometry) behind pairwise!, krig.jl:350; summed in offset order).
static __device__ __noinline__ double gsm_cov_block(const VarioDev& v, const BlockDev& bk, double dx, double dy, double dz) {
  double s = 0.0;
  for (int q = 0; q < bk.nq; ++q)
    s += gsm_cov_vec(v, dx - bk.off[3 * q], dy - bk.off[3 * q + 1], dz - bk.off[3 * q + 2]);
  return s / (double)bk.nq;
}

// x^p by repeated multiplication (universal.jl:60-62: `x .^ n` with integer n), same rounding as the plain loop
// r = 1; r *= x (p times); exponents 0..2 (every drift up to quadratic) take no loop
__device__ __forceinline__ double gsm_ipow(double x, int p) {
  if (p <= 0) return 1.0;
  if (p == 1) return x;
  double r = x * x;
  for (int i = 2; i < p; ++i) r *= x;
  return r;
}

__device__ __forceinline__ double gsm_drift(const ModelDev& m, int n, double x, double y, double z) {
  return gsm_ipow(x, m.exps[n][0]) * gsm_ipow(y, m.exps[n][1]) * gsm_ipow(z, m.exps[n][2]);
}

__device__ __forceinline__ double gsm_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double gsm_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
// Ascending sort of one int per lane (bitonic network, 15 compare-exchange steps).  The solve kernels
// use it to put every neighbour list into a canonical order (ascending record position, unused slots
// last), so results do not depend on the order in which the search kernel emitted the neighbours.
__device__ __forceinline__ int gsm_warp_sort_int(int v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = ((lane & k) == 0);          // ascending block?
      const bool lower = ((lane & j) == 0);        // this lane keeps the smaller element of the pair
      v = (lower == up) ? min(v, o) : max(v, o);
    }
  }
  return v;
}
__device__ __forceinline__ double gsm_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += gsm_shfl_xor(v, o);
  return v;
}
#endif
