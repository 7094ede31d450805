// lorenz_wide.cuh -- Lorenz 96 with one WARP per particle (d = 32 * S, S in {1, 2}): lane l holds components
// l, l + 32, ...; the cyclic neighbours x_{n-1}, x_{n+1}, x_{n-2} of the RHS (reference src/lorenz96.cpp:21-30,
// generalised from the hard-coded d = 8) come from warp shuffles, so a particle's 64 doubles are read and written
// as one coalesced 512-byte row.  Fixed-step RK4 (BASELINE config 5).
#pragma once
#include "common.cuh"

namespace cpimh {

template <int S>
__device__ __forceinline__ void lorenz_wide_rhs(double F, const double* v, double* dv, int lane) {
  double l1[S], l2[S], r1[S];
#pragma unroll
  for (int s = 0; s < S; s++) {
    l1[s] = __shfl_sync(0xffffffffu, v[s], (lane + 31) & 31);
    l2[s] = __shfl_sync(0xffffffffu, v[s], (lane + 30) & 31);
    r1[s] = __shfl_sync(0xffffffffu, v[s], (lane + 1) & 31);
  }
#pragma unroll
  for (int s = 0; s < S; s++) {
    const int sp = (s + S - 1) % S, sn = (s + 1) % S;
    const double xm1 = (lane == 0) ? l1[sp] : l1[s];
    const double xm2 = (lane <= 1) ? l2[sp] : l2[s];
    const double xp1 = (lane == 31) ? r1[sn] : r1[s];
    dv[s] = xm1 * (xp1 - xm2) - v[s] + F;
  }
}
template <int S>
__device__ __forceinline__ void lorenz_wide_rk4(double F, double h, double* v, int lane) {
  double k1[S], k2[S], k3[S], k4[S], tmp[S];
  lorenz_wide_rhs<S>(F, v, k1, lane);
#pragma unroll
  for (int s = 0; s < S; s++) tmp[s] = v[s] + (0.5 * h) * k1[s];
  lorenz_wide_rhs<S>(F, tmp, k2, lane);
#pragma unroll
  for (int s = 0; s < S; s++) tmp[s] = v[s] + (0.5 * h) * k2[s];
  lorenz_wide_rhs<S>(F, tmp, k3, lane);
#pragma unroll
  for (int s = 0; s < S; s++) tmp[s] = v[s] + h * k3[s];
  lorenz_wide_rhs<S>(F, tmp, k4, lane);
#pragma unroll
  for (int s = 0; s < S; s++) v[s] = v[s] + (h / 6.0) * (((k1[s] + 2.0 * k2[s]) + 2.0 * k3[s]) + k4[s]);
}

}  // namespace cpimh
