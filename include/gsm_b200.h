/*
 * gsm_b200.h -- C ABI of the B200-native interpolation hot path for the GeoStatsModels.jl API.
 *
 * Plain C: POD structs, raw pointers + sizes, `int` status returns (0 = ok), no exceptions, no
 * torch / CUDA types in any signature.  This is exactly what a Julia `ccall` (or ctypes / cgo) binds.
 * Every entry point names the reference interface it replaces (paths under JuliaEarth/GeoStatsModels.jl
 * v0.13.10, `src/`).  The library is single-device per context: one process (or one context) per GPU.
 *
 * Coordinates handed to the library are ALREADY scaled by the caller with the alpha of
 * models.jl:272-284 (`_scale`), exactly as the reference scales data, domain and variogram range before
 * calling fitpredictneigh / fitpredictfull (models.jl:118).  The shim in INTEGRATION.md shows how.
 *
 * Pointer arguments documented as "host or device" are classified with cudaPointerGetAttributes;
 * pinned host memory (gsm_host_alloc) makes the copies asynchronous DMA.
 */
#ifndef GSM_B200_H
#define GSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GSM_API __attribute__((visibility("default")))
#else
#define GSM_API
#endif

/* 2: nested / anisotropic variogram descriptors, gsm_set_option, the gsm_multi_* family.  Entry points added since
 * (gsm_set_block_support, gsm_nn_local) do not change any existing layout or signature, so the number stays. */
#define GSM_ABI_VERSION 2
#define GSM_MAX_DIM 3
#define GSM_MAX_DRIFT 10   /* UniversalKriging(fun, deg=2, dim=3) -> 10 monomials (universal.jl:54-77) */
#define GSM_MAX_NEIGHBORS 64
#define GSM_MAX_STRUCTURES 4   /* nested structures of a composite variogram */

/* ---- status codes ------------------------------------------------------------------------- */
enum {
  GSM_OK = 0,
  GSM_ERR_INVALID = 1,      /* bad argument (mirrors ArgumentError / @assert in the reference)   */
  GSM_ERR_CUDA = 2,         /* CUDA runtime error, see gsm_last_error                              */
  GSM_ERR_NO_SAMPLES = 3,   /* gsm_set_samples not called                                          */
  GSM_ERR_UNSUPPORTED = 4,  /* signature the GPU path does not claim (falls to the generic Julia method) */
  GSM_ERR_SINGULAR = 5      /* global factorisation failed: status(fitted) == false (krig.jl:49-52) */
};

/* per-target status byte written to out_status (models.jl:178-181; krig.jl:152 check=false) */
enum {
  GSM_ST_OK = 0,
  GSM_ST_MISSING = 1,   /* fewer than minneighbors found -> `missing`               */
  GSM_ST_SINGULAR = 2   /* non-positive pivot in the local system -> NaN outputs    */
};

/* ---- descriptors (POD) -------------------------------------------------------------------- */

/* GeoStatsFunctions variogram, isotropic, Euclidean metric ball (call sites krig.jl:125,347,297).
 * gamma(h), t = h/range:
 *   GAUSSIAN     (sill-n')(1-exp(-3t^2)) + (h>0) n',  n' = nugget + 1e-6 (the floor replaces the nugget everywhere)
 *   SPHERICAL    (sill-nugget)(1.5t-0.5t^3) for h<range, (sill-nugget) otherwise, + (h>0) nugget
 *   EXPONENTIAL  (sill-nugget)(1-exp(-3t)) + (h>0) nugget
 *   CUBIC        (sill-nugget)(7t^2 - 8.75t^3 + 3.5t^5 - 0.75t^7) for t<1, (sill-nugget) otherwise, + (h>0) nugget
 *   PENTASPHERICAL (sill-nugget)(1.875t - 1.25t^3 + 0.375t^5) for t<1, (sill-nugget) otherwise, + (h>0) nugget
 *   SINEHOLE     (sill-nugget)(1 - sin(pi t)/(pi t)) + (h>0) nugget
 *   CIRCULAR     (sill-nugget)(1 - (2/pi) acos(t) + (2t/pi) sqrt(1-t^2)) for t<1, (sill-nugget) otherwise, + (h>0) nugget
 *   NUGGET       (h>0) nugget   (NuggetEffect; sill = nugget)
 * The system is assembled in covariance form sill - gamma(h) (krig.jl:165-179, 364-378).
 *
 * nstruct == 0: one isotropic structure described by kind / sill / nugget / range (the round-1 layout; the first 32
 * bytes are unchanged).  nstruct >= 1: a composite (nested) variogram gamma = sum_i gamma_i (GeoStatsFunctions
 * CompositeFunction, "sum of structures", comment at krig.jl:147); every structure has its own family, total sill,
 * nugget and metric ball: t_i = |metric_i (x - y)| / range_i with `metric` a row-major 3 x 3 matrix -- the identity for
 * an isotropic ball, diag(1/r_1, 1/r_2, 1/r_3) R' with range = 1 for an anisotropic (Mahalanobis) ball with radii r and
 * rotation R.  The top-level kind / sill / nugget / range are ignored then (total sill = sum of the structures' sills).
 * The neighbour search stays Euclidean (`distance` of fitpredict, models.jl:112).  Composite / anisotropic / new-family
 * variograms run on the general kernels (maxneighbors <= 64; the shared-factor fast path serves the three
 * isotropic single-structure families above). */
enum { GSM_VARIO_GAUSSIAN = 0, GSM_VARIO_SPHERICAL = 1, GSM_VARIO_EXPONENTIAL = 2, GSM_VARIO_CUBIC = 3,
       GSM_VARIO_PENTASPHERICAL = 4, GSM_VARIO_SINEHOLE = 5, GSM_VARIO_CIRCULAR = 6, GSM_VARIO_NUGGET = 7 };
typedef struct gsm_structure {
  int32_t kind;
  int32_t _pad;
  double sill;        /* total sill of this structure (its nugget included) */
  double nugget;
  double range;       /* ALREADY multiplied by alpha */
  double metric[9];   /* row-major 3 x 3, see above */
} gsm_structure;
typedef struct gsm_variogram {
  int32_t kind;
  int32_t nstruct;  /* 0: the single isotropic structure below; 1..GSM_MAX_STRUCTURES: structs[] */
  double sill;      /* total sill                         */
  double nugget;
  double range;     /* ALREADY multiplied by alpha        */
  gsm_structure structs[GSM_MAX_STRUCTURES];
} gsm_variogram;

/* Kriging variant: SimpleKriging (simple.jl:29-69), OrdinaryKriging (ordinary.jl:25-77),
 * UniversalKriging built by (deg, dim) (universal.jl:45-137): drift n at point x is
 * prod_d x[d]^exps[n][d] on the scaled raw coordinates. */
enum { GSM_KRIG_SIMPLE = 0, GSM_KRIG_ORDINARY = 1, GSM_KRIG_UNIVERSAL = 2 };
typedef struct gsm_model {
  int32_t kind;
  int32_t ndrift;                         /* UK only, 1..GSM_MAX_DRIFT                    */
  double mean;                            /* SK only (simple.jl:55-69)                    */
  int32_t exps[GSM_MAX_DRIFT][GSM_MAX_DIM];
} gsm_model;

/* Target domain: a CartesianGrid (centroid of element (i0,j0,k0), 0-based, i fastest:
 * (origin[d] + i_d*spacing[d]) + spacing[d]/2, every operation rounded once) or an explicit
 * PointSet (npoints x dim doubles, AoS = memory order of a Julia Vector{Point}).
 * [first, first+count) selects the shard of the linear target index this call predicts
 * (models.jl:236-247 chunks the same index range over threads); count < 0 means "to the end". */
enum { GSM_TARGETS_GRID = 0, GSM_TARGETS_POINTS = 1 };
typedef struct gsm_targets {
  int32_t kind;
  int32_t dim;
  double origin[GSM_MAX_DIM];    /* scaled */
  double spacing[GSM_MAX_DIM];   /* scaled */
  int64_t dims[GSM_MAX_DIM];
  const double* points;          /* host or device, scaled; GSM_TARGETS_POINTS only */
  int64_t npoints;
  int64_t first;
  int64_t count;
} gsm_targets;

/* Options of fitpredictneigh (models.jl:131-198). */
typedef struct gsm_neighbor_opts {
  int32_t maxneighbors;   /* clamped like models.jl:144-147 (outside [1,N] -> N); more than GSM_MAX_NEIGHBORS
                             after clamping -> GSM_ERR_UNSUPPORTED */
  int32_t minneighbors;   /* clamped like models.jl:148-150                                        */
  double radius;          /* > 0: KBallSearch with this (scaled) ball radius (models.jl:158); <= 0: KNearestSearch */
} gsm_neighbor_opts;

/* Per-phase device times of the last call on a context (CUDA events on the library's stream). */
typedef struct gsm_timings {
  double h2d_ms;       /* sample / target upload                               */
  double bin_ms;       /* grid-bin build (KNearestSearch construction analogue) */
  double knn_ms;       /* neighbour search kernel(s)                            */
  double solve_ms;     /* assemble + factor + solve kernel(s)                   */
  double d2h_ms;       /* result download                                       */
  double total_ms;     /* first event to last event                             */
  int64_t launches;    /* kernels launched by the library in the call           */
  double index_ms;     /* part of solve_ms spent in the job-index kernel (shared-factor local Kriging) */
} gsm_timings;

/* Per-target status counts of the last neighbourhood call (reduced on the device, so a host that only needs
 * "did anything go wrong" does not have to scan the status / variance columns).  The reference raises on a
 * non-positive variance when it builds Normal(mean, var) (models.jl:298-299): nonpositive_var counts the
 * targets with status OK whose variance is not > 0.  total = targets covered, -1 when no status was produced. */
typedef struct gsm_summary {
  int64_t missing;
  int64_t singular;
  int64_t nonpositive_var;
  int64_t total;
} gsm_summary;

typedef struct gsm_ctx gsm_ctx;   /* opaque; owns a device, a stream, sample + scratch buffers */
typedef struct gsm_fit gsm_fit;   /* opaque; a fitted global Kriging model (FittedKriging)     */

/* ---- lifecycle ---------------------------------------------------------------------------- */
GSM_API int gsm_abi_version(void);
GSM_API int gsm_create(int device, gsm_ctx** out);
GSM_API int gsm_destroy(gsm_ctx* ctx);
/* Message of the last error on this context (or of the last failed gsm_create when ctx == NULL). */
GSM_API const char* gsm_last_error(const gsm_ctx* ctx);
GSM_API int gsm_get_timings(const gsm_ctx* ctx, gsm_timings* out);
GSM_API int gsm_get_summary(const gsm_ctx* ctx, gsm_summary* out);
GSM_API int gsm_synchronize(gsm_ctx* ctx);
/* Developer options (A/B profiling and per-kernel parity tests; the defaults are the product, and the library
 * reads no environment variable): "local_impl" = auto | patch | shr | pipe | simt (which local-Kriging kernel serves
 * maxneighbors <= 32), "shr_smax" = 0..3, "bin_occupancy" = samples per search cell (next gsm_set_samples). */
GSM_API int gsm_set_option(gsm_ctx* ctx, const char* key, const char* value);

/* Change of support (fitpredict `point=false`, /root/reference/src/models.jl:114-115,136,205): every target of the
 * following gsm_local_krige / gsm_global_krige_predict calls stands for the nq points
 * centroid + offsets[q] (equal weights; offsets nq x dim, AoS, in the coordinates of the samples, i.e. already
 * multiplied by the scale factor).  The right-hand side becomes the mean covariance between a sample and those points
 * (pairwise! on [g0] + rhsbanded!, krig.jl:342-378), covzero the mean covariance over all nq^2 pairs of them
 * (krig.jl:295-301); the search centre and the drift functions stay at the centroid (models.jl:164,
 * universal.jl:125-126), and IDW / NN only ever see the centroid (idw.jl:82, nn.jl:57-66), so they ignore this.
 * Which points represent a cell is the caller's rule (upstream it lives in GeoStatsFunctions, not in the reference
 * tree); congruent grid cells need one offset table.  nq = 0 returns to point support. */
#define GSM_MAX_BLOCK_POINTS 512
GSM_API int gsm_set_block_support(gsm_ctx* ctx, const double* offsets, int dim, int nq);

/* Pinned host buffers so outputs DMA straight into caller-owned memory (Julia: unsafe_wrap). */
GSM_API int gsm_host_alloc(void** out, int64_t bytes);
GSM_API int gsm_host_free(void* p);

/* ---- data: replaces `domain(dat)` / `values(dat)` of the geotable given to fit / fitpredict ----
 * coords: n x dim doubles (AoS), values: n doubles; host or device; copied, never retained.
 * Builds the grid bins used by the neighbour search (KNearestSearch construction, models.jl:155). */
GSM_API int gsm_set_samples(gsm_ctx* ctx, const double* coords, int32_t dim, int64_t n, const double* values);

/* Same, with the reference's `_scale` (models.jl:272-291) done on the device: uploads the RAW coordinates,
 * reduces absmax = max |coordinate|, sets alpha = 1 / max(scale_floor, absmax) -- the caller passes
 * scale_floor = max(variogram range (1 if infinite), absmax of the target domain) -- multiplies the coordinates
 * by alpha (one rounding each, like `dat |> Scale(alpha)`) and builds the bins.  out_alpha receives alpha (the
 * caller scales targets, radius and variogram with it); out_nan the number of NaN values (missing data are not
 * claimed by the GPU path).  The host never walks the sample arrays. */
GSM_API int gsm_set_samples_scaled(gsm_ctx* ctx, const double* coords, int32_t dim, int64_t n, const double* values,
                           double scale_floor, double* out_alpha, int64_t* out_nan);

/* ---- test hook: Meshes `search!` (models.jl:164).  out_idx: M x k int32 (0-based sample rows,
 * ascending by (d2, index); unused slots = -1), out_n: M int32 neighbours found. host or device. */
GSM_API int gsm_knn(gsm_ctx* ctx, const gsm_targets* targets, const gsm_neighbor_opts* opts,
            int32_t* out_idx, int32_t* out_n);

/* ---- fitpredictneigh with a Kriging model (models.jl:131-198 + krig.jl:58-75,316-339,245-293) --
 * out_mean: M doubles.  out_var: M doubles or NULL (prob=false): sigma^2 INCLUDING the +1e-10 of
 * krig.jl:278-279, i.e. the second field of the Normal that `_marginals` builds (models.jl:298-299).
 * out_status: M bytes or NULL.  out_nbr: M x maxneighbors int32 or NULL.  All host or device. */
GSM_API int gsm_local_krige(gsm_ctx* ctx, const gsm_variogram* vario, const gsm_model* model,
                    const gsm_targets* targets, const gsm_neighbor_opts* opts,
                    double* out_mean, double* out_var, uint8_t* out_status, int32_t* out_nbr);

/* ---- fit / predict / predictprob / status with all samples (krig.jl:58-75, 215-237, 49-52);
 * fitpredictfull (models.jl:200-234) = fit once + predict over the domain. */
GSM_API int gsm_global_krige_fit(gsm_ctx* ctx, const gsm_variogram* vario, const gsm_model* model, gsm_fit** out);
GSM_API int gsm_global_krige_status(const gsm_fit* fit);   /* 1 = factorisation succeeded */
GSM_API int gsm_global_krige_predict(gsm_fit* fit, const gsm_targets* targets, double* out_mean, double* out_var);
GSM_API int gsm_global_krige_free(gsm_fit* fit);

/* ---- IDW (idw.jl:43-89): weights 1/d^exponent; exact hits return the first zero-distance sample.
 * opts == NULL or opts->maxneighbors == 0: all samples (fitpredictfull); else local (fitpredictneigh). */
GSM_API int gsm_idw(gsm_ctx* ctx, double exponent, const gsm_targets* targets, const gsm_neighbor_opts* opts,
            double* out_mean, uint8_t* out_status);

/* ---- NN (nn.jl:47-54): value of the nearest sample (ties: the lowest sample row, the order a stable sortperm of the
 * distances gives).  SURVEY section 8(f) "next" row; missing values are not claimed. */
GSM_API int gsm_nn(gsm_ctx* ctx, const gsm_targets* targets, double* out_value);
/* The NN model behind a neighbour search: per-target body of fitpredictneigh (/root/reference/src/models.jl:162-184)
 * with fit / predict of src/nn.jl:30-54 -- the nearest of the maxneighbors nearest samples (inside the ball when
 * opts->radius > 0) THAT HAS A VALUE (NaN = `missing` is skipped, nn.jl:51-53); out_status 1 = `missing`: fewer than
 * minneighbors samples found, or none of them has a value. */
GSM_API int gsm_nn_local(gsm_ctx* ctx, const gsm_targets* targets, const gsm_neighbor_opts* opts, double* out_value,
                         uint8_t* out_status);

/* ---- multi-GPU: ONE caller drives all GPUs of a box (reference analogue: the contiguous
 * index chunks of `_predictionthread!`, models.jl:236-247, one private fitted state per chunk, no exchange step).
 * gsm_multi_create sets up one context per device and, for ndev > 1, an NCCL communicator (ncclCommInitAll, warmed
 * once, cached in the handle; NCCL is bound at run time, so the single-GPU entry points do not depend on it).
 * gsm_multi_set_samples[_scaled] uploads the samples ONCE to devices[0], replicates them with one ncclBroadcast over
 * NVLink and builds the search bins on every device.  The predict calls split the selected target range
 * [first, first + count) into ndev contiguous slabs (aligned to pairs of grid rows / planes so that no slab cuts a
 * patch of the solve kernels), run one host thread per device and write every slab straight into the caller's HOST
 * output arrays (pinned memory, gsm_host_alloc, makes that a direct DMA); no gather.  The handle is not re-entrant. */
typedef struct gsm_multi gsm_multi;
typedef struct gsm_multi_timings {
  double upload_ms;    /* set_samples: host -> devices[0] (+ absmax / scale kernels)                 */
  double bcast_ms;     /* set_samples: the NCCL broadcast (device time on devices[0]'s stream)       */
  double bins_ms;      /* set_samples: bins on every device (host wall clock, devices in parallel)   */
  double predict_ms;   /* last predict call: max over devices of gsm_timings.total_ms                */
  double wall_ms;      /* last call: host wall clock                                                 */
} gsm_multi_timings;
GSM_API int gsm_multi_create(const int32_t* devices, int32_t ndev, gsm_multi** out);
GSM_API int gsm_multi_destroy(gsm_multi* mg);
GSM_API const char* gsm_multi_last_error(const gsm_multi* mg);   /* mg == NULL: last failed gsm_multi_create */
GSM_API int gsm_multi_device_count(const gsm_multi* mg);
GSM_API gsm_ctx* gsm_multi_context(gsm_multi* mg, int32_t i);    /* per-device context: timings, options */
GSM_API int gsm_multi_get_timings(const gsm_multi* mg, gsm_multi_timings* out);
GSM_API int gsm_multi_get_summary(const gsm_multi* mg, gsm_summary* out);   /* summed over the devices */
GSM_API int gsm_multi_set_samples(gsm_multi* mg, const double* coords, int32_t dim, int64_t n, const double* values);
GSM_API int gsm_multi_set_samples_scaled(gsm_multi* mg, const double* coords, int32_t dim, int64_t n, const double* values,
                                 double scale_floor, double* out_alpha, int64_t* out_nan);
/* the slab [out_first, out_first + out_count) device i predicts for this target descriptor */
GSM_API int gsm_multi_slab(const gsm_multi* mg, const gsm_targets* targets, int32_t i, int64_t* out_first, int64_t* out_count);
GSM_API int gsm_multi_local_krige(gsm_multi* mg, const gsm_variogram* vario, const gsm_model* model,
                          const gsm_targets* targets, const gsm_neighbor_opts* opts, double* out_mean, double* out_var,
                          uint8_t* out_status);
GSM_API int gsm_multi_idw(gsm_multi* mg, double exponent, const gsm_targets* targets, const gsm_neighbor_opts* opts,
                  double* out_mean, uint8_t* out_status);
/* fitpredictfull: the factorisation is recomputed on every device, targets are split like above;
 * out_fit_status: 1 = the factorisation succeeded everywhere (status(fitted), krig.jl:49-52) */
GSM_API int gsm_multi_global_krige(gsm_multi* mg, const gsm_variogram* vario, const gsm_model* model,
                           const gsm_targets* targets, double* out_mean, double* out_var, int32_t* out_fit_status);

/* ---- measurement helper: sustained FP64 FMA throughput of this device in TFLOP/s (roofline peak) */
GSM_API int gsm_measure_fp64_peak(gsm_ctx* ctx, double* out_dfma_tflops, double* out_dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* GSM_B200_H */
