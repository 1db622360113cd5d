// Covariance evaluation shared by the local-Kriging kernels: sill - gamma(h) from a squared distance with a
// Newton-refined hardware reciprocal square root instead of the IEEE sqrt + slow-path sequence (same values to
// ~1 ulp; reference: the variogram forms of GeoStatsFunctions used through krig.jl:165-179).
#pragma once
#include "gsm_internal.cuh"

namespace gsm_cov {

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
__host__ __device__ __forceinline__ CovFast make_cov(const VarioDev& v) {
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


}  // namespace gsm_cov
