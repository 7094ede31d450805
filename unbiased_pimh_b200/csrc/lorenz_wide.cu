// lorenz_wide.cu -- stand-alone wide Lorenz 96 transition kernel (see lorenz_wide.cuh)
#include "lorenz_wide.cuh"

namespace cpimh {

template <int C>
__device__ void lorenz_wide_body(const double* xin, long long n, double t0, double t1, double F, int substeps, double* xout) {
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= n) return;
  constexpr int D = 32 * C;
  double v[C];
#pragma unroll
  for (int s = 0; s < C; s++) v[s] = xin[(size_t)p * D + lane * C + s];
  const int ns = substeps < 1 ? 1 : substeps;
  const double h = (t1 - t0) / ns;
  for (int q = 0; q < ns; q++) lorenz_wide_rk4<C>(F, h, v, lane);
#pragma unroll
  for (int s = 0; s < C; s++) xout[(size_t)p * D + lane * C + s] = v[s];
}
__global__ void lorenz_wide_rk4_kernel(const double* xin, long long n, int d, double t0, double t1, double F, int substeps, double* xout) {
  if (d == 64) lorenz_wide_body<2>(xin, n, t0, t1, F, substeps, xout);
  else lorenz_wide_body<1>(xin, n, t0, t1, F, substeps, xout);
}

}  // namespace cpimh
