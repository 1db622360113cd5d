// Inverse-distance weights (reference src/idw.jl:59-89), shared by the IDW kernels and the fused k-NN + IDW path.
#pragma once
#include "gsm_internal.cuh"

namespace gsm_idww {

// 1/d, d > 0 finite: hardware seed + two Newton steps (full double precision, no slow-path branch)
__device__ __forceinline__ double idw_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  y = fma(y, fma(-d, y, 1.0), y);
  y = fma(y, fma(-d, y, 1.0), y);
  return y;
}

// EXPMODE: 1 -> e == 1, 2 -> e == 2, 3 -> other positive integer, 0 -> real exponent
template <int EXPMODE>
__device__ __forceinline__ double idw_weight(double d2, double e, int ei) {
  if (EXPMODE == 2) return (d2 > 1e-290 && d2 < 1e290) ? idw_rcp(d2) : 1.0 / d2;
  if (EXPMODE == 1) return 1.0 / sqrt(d2);
  const double d = sqrt(d2);
  if (EXPMODE == 3) {
    double p = d;
    for (int i = 1; i < ei; ++i) p *= d;
    return 1.0 / p;
  }
  return 1.0 / pow(d, e);
}

// EXPMODE of an exponent: 1 -> e == 1, 2 -> e == 2, 3 -> other positive integer (ei), 0 -> real exponent
inline int exp_mode(double e, int& ei) {
  ei = (int)e;
  if (e == 1.0) return 1;
  if (e == 2.0) return 2;
  if (e > 0 && e == (double)ei && ei <= 64) return 3;
  return 0;
}

// run-time dispatch over EXPMODE (for callers that are not templated on it)
__device__ __forceinline__ double idw_weight_rt(int mode, double d2, double e, int ei) {
  switch (mode) {
    case 1: return idw_weight<1>(d2, e, ei);
    case 2: return idw_weight<2>(d2, e, ei);
    case 3: return idw_weight<3>(d2, e, ei);
    default: return idw_weight<0>(d2, e, ei);
  }
}

}  // namespace gsm_idww
