// lorenz_wide.cuh -- Lorenz 96 with one WARP per particle (d = 32 * C, C in {1, 2}): lane l holds the C adjacent
// components l*C .. l*C + C - 1, so a particle's state is one coalesced row (512 B for d = 64, a 16-byte vector load per
// lane) and the cyclic neighbours x_{n-1}, x_{n+1}, x_{n-2} of the RHS (reference src/lorenz96.cpp:21-30, generalised
// from the hard-coded d = 8) come from at most three warp shuffles.  Fixed-step RK4 (BASELINE config 5).
#pragma once
#include "common.cuh"

namespace cpimh {

template <int C>
__device__ __forceinline__ void lorenz_wide_rhs(double F, const double* v, double* dv, int lane) {
  const int left = (lane + 31) & 31, right = (lane + 1) & 31;
  if (C == 1) {
    const double xm1 = __shfl_sync(0xffffffffu, v[0], left);
    const double xm2 = __shfl_sync(0xffffffffu, v[0], (lane + 30) & 31);
    const double xp1 = __shfl_sync(0xffffffffu, v[0], right);
    dv[0] = xm1 * (xp1 - xm2) - v[0] + F;
  } else {
    const double l0 = __shfl_sync(0xffffffffu, v[0], left);       // x_{2l-2}
    const double l1 = __shfl_sync(0xffffffffu, v[C - 1], left);   // x_{2l-1}
    const double r0 = __shfl_sync(0xffffffffu, v[0], right);      // x_{2l+2}
    dv[0] = l1 * (v[C - 1] - l0) - v[0] + F;                      // n = 2l:   x_{n-1} (x_{n+1} - x_{n-2}) - x_n + F
    dv[C - 1] = v[0] * (r0 - l1) - v[C - 1] + F;                  // n = 2l+1
  }
}
template <int C>
__device__ __forceinline__ void lorenz_wide_rk4(double F, double h, double* v, int lane) {
  double k1[C], k2[C], k3[C], k4[C], tmp[C];
  lorenz_wide_rhs<C>(F, v, k1, lane);
#pragma unroll
  for (int s = 0; s < C; s++) tmp[s] = v[s] + (0.5 * h) * k1[s];
  lorenz_wide_rhs<C>(F, tmp, k2, lane);
#pragma unroll
  for (int s = 0; s < C; s++) tmp[s] = v[s] + (0.5 * h) * k2[s];
  lorenz_wide_rhs<C>(F, tmp, k3, lane);
#pragma unroll
  for (int s = 0; s < C; s++) tmp[s] = v[s] + h * k3[s];
  lorenz_wide_rhs<C>(F, tmp, k4, lane);
#pragma unroll
  for (int s = 0; s < C; s++) v[s] = v[s] + (h / 6.0) * (((k1[s] + 2.0 * k2[s]) + 2.0 * k3[s]) + k4[s]);
}
// standard normals for the C components of this lane: Philox draw j of particle i gives components 2j, 2j+1
template <int C>
__device__ __forceinline__ void lorenz_wide_normals(const Stream& s, unsigned i, int t, int purpose, int lane, double* z) {
  double ua, ub, z0, z1;
  if (C == 2) {
    philox_u2(s, i, t, purpose, lane, ua, ub);
    normal2(ua, ub, z0, z1);
    z[0] = z0; z[C - 1] = z1;
  } else {
    philox_u2(s, i, t, purpose, lane >> 1, ua, ub);     // both lanes of a pair evaluate the same block
    normal2(ua, ub, z0, z1);
    z[0] = (lane & 1) ? z1 : z0;
  }
}

}  // namespace cpimh
