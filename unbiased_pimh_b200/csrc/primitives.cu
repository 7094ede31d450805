// primitives.cu -- batched stand-alone versions of the reference's native primitives (include/cpimh.h
// "primitives"): resamplers, weight normalisation, Gaussian kernels, model pieces.  These are the entry points
// the R shim binds the `_CoupledPIMH_*` .Call names to.  Resampler outputs are BIT-EXACT with the reference
// natives for identical (weights, uniforms): the CDF is accumulated in the reference's left-to-right fp64 order
// by one lane per problem (problems are spread over CTAs), everything else is parallel.
#include <math.h>
#include <string.h>
#include <vector>
#include "host_common.h"

namespace cpimh {

struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() { for (void* p : ptrs) cudaFree(p); }
  template <class T> int alloc(T** out, size_t n) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(T) * (n ? n : 1));
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes) failed: %s", sizeof(T) * n, cudaGetErrorString(e)); return CPIMH_ERR_CUDA; }
    ptrs.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return CPIMH_OK;
  }
  template <class T> int upload(T** out, const T* host, size_t n) {
    int rc = alloc(out, n);
    if (rc) return rc;
    CPIMH_CUDA_CHECK(cudaMemcpy(*out, host, sizeof(T) * n, cudaMemcpyHostToDevice));
    return CPIMH_OK;
  }
};
static int sync_check() {
  CPIMH_CUDA_CHECK(cudaGetLastError());
  CPIMH_CUDA_CHECK(cudaDeviceSynchronize());
  return CPIMH_OK;
}

// ---------------------------------------------------------------- device helpers on global arrays
__device__ __forceinline__ int g_count_le(const double* c, int n, double u) {
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (c[mid] <= u) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int g_count_lt(const double* c, int n, double u) {
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (c[mid] < u) lo = mid + 1; else hi = mid; }
  return lo;
}
// left-to-right running sum, exactly `csw += weights(j)` (src/systematic.cpp:17) / Rcpp::cumsum (src/multinomial.cpp:17)
__device__ __forceinline__ void seq_cumsum(const double* w, double* c, int n) {
  double acc = w[0];
  c[0] = acc;
  for (int i = 1; i < n; i++) { acc = acc + w[i]; c[i] = acc; }
}
// std::random_shuffle with randWrapper (src/multinomial.cpp:6-10,30): targets in parallel, swaps in order
__device__ __forceinline__ void shuffle_targets(const double* us, int* jidx, int n) {
  for (int i = 1 + threadIdx.x; i < n; i += blockDim.x) {
    int j = (int)floor(us[i - 1] * (double)(i + 1));
    jidx[i] = j > i ? i : j;
  }
}
__device__ __forceinline__ void shuffle_apply(int* a, const int* jidx, int n) {
  for (int i = 1; i < n; i++) { int j = jidx[i]; int t = a[i]; a[i] = a[j]; a[j] = t; }
}

// ---------------------------------------------------------------- systematic_resampling_n_  (src/systematic.cpp:7-32)
__global__ void systematic_kernel(const double* w, int N, int n, const double* u, double* cdf, double* ladder, int* anc) {
  const long long b = blockIdx.x;
  const double* wb = w + (size_t)b * N;
  double* c = cdf + (size_t)b * N;
  double* l = ladder + (size_t)b * n;
  if (threadIdx.x == 0) seq_cumsum(wb, c, N);
  if (threadIdx.x == 32 || (blockDim.x <= 32 && threadIdx.x == 0)) {
    double uk = u[b] / n;                       // :11
    const double inc = 1. / n;
    for (int k = 0; k < n; k++) { l[k] = uk; uk = uk + inc; }   // :19 repeated addition
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    int a = g_count_lt(c, N, l[k]);             // smallest j with csw_j >= u_k (:15-18)
    anc[(size_t)b * n + k] = a >= N ? N - 1 : a;
  }
}

// ---------------------------------------------------------------- multinomial / coupled multinomial (src/multinomial.cpp:13-71)
// NW = 1: multinomial_resampling_n_ (u = sumw * e);  NW = 2: coupled (u = e against both CDFs, shared permutation)
template <int NW>
__global__ void multinomial_kernel(const double* w1, const double* w2, int N, int n, const double* e, const double* us,
                                   double* cdf, int* tmp, int* anc) {
  const long long b = blockIdx.x;
  double* c1 = cdf + (size_t)b * NW * N;
  double* c2 = c1 + N;
  if (threadIdx.x == 0) seq_cumsum(w1 + (size_t)b * N, c1, N);
  if (NW == 2 && threadIdx.x == 32) seq_cumsum(w2 + (size_t)b * N, c2, N);
  __syncthreads();
  const double* eb = e + (size_t)b * n;
  int* out = anc + (size_t)b * NW * n;
  if (NW == 1) {
    const double sumw = c1[N - 1];              // :19
    int* a = us ? tmp + (size_t)b * 2 * n : out;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      int j = g_count_le(c1, N, sumw * eb[k]);  // :24-28
      a[k] = j >= N ? N - 1 : j;
    }
    if (us) {
      int* jidx = a + n;
      shuffle_targets(us + (size_t)b * (n - 1), jidx, n);
      __syncthreads();
      if (threadIdx.x == 0) shuffle_apply(a, jidx, n);
      __syncthreads();
      for (int k = threadIdx.x; k < n; k += blockDim.x) out[k] = a[k];
    }
  } else {
    int* idx = tmp + (size_t)b * 2 * n;
    int* jidx = idx + n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) idx[k] = k;   // :60-64
    if (us) {
      shuffle_targets(us + (size_t)b * (n - 1), jidx, n);
      __syncthreads();
      if (threadIdx.x == 0) shuffle_apply(idx, jidx, n);            // :65
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {             // :66-69
      const double uk = eb[idx[k]];                                 // :49 no sumw scaling
      int j1 = g_count_le(c1, N, uk), j2 = g_count_le(c2, N, uk);
      out[k] = j1 >= N ? N - 1 : j1;
      out[n + k] = j2 >= N ? N - 1 : j2;
    }
  }
}

// ---------------------------------------------------------------- indexmatching_cpp (src/coupledresamplingindexmatching.cpp:8-57)
// phase 1: nu, alpha (sequential sum, :13-16), mu, R1, R2 (:18-21), coupled flags and their ranks (:26-33)
__global__ void indexmatching_prepare(const double* w1, const double* w2, int N, int n, const double* uc,
                                      double* mu, double* R1, double* R2, int* flag, int* rank, int* ncoupled) {
  const long long b = blockIdx.x;
  const double *a1 = w1 + (size_t)b * N, *a2 = w2 + (size_t)b * N;
  __shared__ double s_alpha;
  if (threadIdx.x == 0) {
    double alpha = 0;
    for (int i = 0; i < N; i++) alpha += fmin(a1[i], a2[i]);
    s_alpha = alpha;
    int nc = 0;                                  // ranks within the coupled / uncoupled subsequences
    for (int i = 0; i < n; i++) {
      int cpl = uc[(size_t)b * N + i] < alpha;
      flag[(size_t)b * n + i] = cpl;
      rank[(size_t)b * n + i] = cpl ? nc : (i - nc);
      nc += cpl;
    }
    ncoupled[b] = nc;
  }
  __syncthreads();
  const double alpha = s_alpha;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double nu = fmin(a1[i], a2[i]);
    mu[(size_t)b * N + i] = nu / alpha;
    R1[(size_t)b * N + i] = (a1[i] - nu) / (1 - alpha);
    R2[(size_t)b * N + i] = (a2[i] - nu) / (1 - alpha);
  }
}
// phase 2: nested multinomial (ncoupled draws from mu) and coupled multinomial (rest from R1/R2), then interleave (:36-55)
__global__ void indexmatching_draw(const double* mu, const double* R1, const double* R2, int N, int n, const int* ncoupled,
                                   const double* e_mult, const double* us_mult, const double* e_cm, const double* us_cm,
                                   const int* flag, const int* rank, double* cdf, int* tmp, int* anc) {
  const long long b = blockIdx.x;
  const int nc = ncoupled[b], nr = n - nc;
  double* c0 = cdf + (size_t)b * 3 * N;
  double *c1 = c0 + N, *c2 = c0 + 2 * N;
  if (threadIdx.x == 0 && nc > 0) seq_cumsum(mu + (size_t)b * N, c0, N);
  if (threadIdx.x == 32 && nr > 0) seq_cumsum(R1 + (size_t)b * N, c1, N);
  if (threadIdx.x == 64 && nr > 0) seq_cumsum(R2 + (size_t)b * N, c2, N);
  __syncthreads();
  int* a = tmp + (size_t)b * 5 * n;      // [n] coupled draws, [n] idx, [n] jidx, [2n] residual draws
  int* idx = a + n;
  int* jidx = a + 2 * n;
  int* aR = a + 3 * n;
  if (nc > 0) {
    const double sumw = c0[N - 1];
    for (int k = threadIdx.x; k < nc; k += blockDim.x) { int j = g_count_le(c0, N, sumw * e_mult[(size_t)b * n + k]); a[k] = j >= N ? N - 1 : j; }
    if (us_mult) {
      shuffle_targets(us_mult + (size_t)b * (n - 1), jidx, nc);
      __syncthreads();
      if (threadIdx.x == 0) shuffle_apply(a, jidx, nc);
    }
  }
  __syncthreads();
  if (nr > 0) {
    for (int k = threadIdx.x; k < nr; k += blockDim.x) idx[k] = k;
    if (us_cm) {
      shuffle_targets(us_cm + (size_t)b * (n - 1), jidx, nr);
      __syncthreads();
      if (threadIdx.x == 0) shuffle_apply(idx, jidx, nr);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nr; k += blockDim.x) {
      const double uk = e_cm[(size_t)b * n + idx[k]];
      int j1 = g_count_le(c1, N, uk), j2 = g_count_le(c2, N, uk);
      aR[k] = j1 >= N ? N - 1 : j1;
      aR[nr + k] = j2 >= N ? N - 1 : j2;
    }
  }
  __syncthreads();
  int* out = anc + (size_t)b * 2 * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = rank[(size_t)b * n + i];
    if (flag[(size_t)b * n + i]) { out[i] = a[r]; out[n + i] = a[r]; }
    else { out[i] = aR[r]; out[n + i] = aR[nr + r]; }
  }
}

// ---------------------------------------------------------------- sorted uniforms (src/multinomial.cpp:20-24) and weights (R/CPF.R:61-66)
__global__ void sorted_uniforms_kernel(const double* u, int n, double* e) {
  extern __shared__ __align__(16) unsigned char raw[];
  double* s = reinterpret_cast<double*>(raw);
  __shared__ double scr[32];
  const long long b = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = log(u[(size_t)b * n + i]) / (double)(i + 1);
  __syncthreads();
  block_scan_inplace<true>(s, n, scr);
  for (int i = threadIdx.x; i < n; i += blockDim.x) e[(size_t)b * n + i] = exp(s[i]);
}
__global__ void normalize_kernel(const double* logw, int N, double* w, double* logz) {
  __shared__ double scr[32];
  const long long b = blockIdx.x;
  const double* lw = logw + (size_t)b * N;
  double m = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmax(m, lw[i]);
  m = block_max(m, scr);
  double s = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { double v = exp(lw[i] - m); w[(size_t)b * N + i] = v; s += v; }
  s = block_sum(s, scr);
  for (int i = threadIdx.x; i < N; i += blockDim.x) w[(size_t)b * N + i] = w[(size_t)b * N + i] / s;
  if (logz && threadIdx.x == 0) logz[b] = log(s / (double)N) + m;
}

// ---------------------------------------------------------------- Gaussian kernels (src/mvnorm.cpp)
#define CPIMH_MVN_MAXD 64
__global__ void dmvnorm_chol_kernel(const double* x, int d, long long n, const double* mean, const double* L, double cst, double* out) {
  extern __shared__ __align__(16) unsigned char raw[];
  double* sL = reinterpret_cast<double*>(raw);
  double* sm = sL + d * d;
  for (int i = threadIdx.x; i < d * d; i += blockDim.x) sL[i] = L[i];
  for (int i = threadIdx.x; i < d; i += blockDim.x) sm[i] = mean[i];
  __syncthreads();
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double r[CPIMH_MVN_MAXD];
  for (int i = 0; i < d; i++) r[i] = x[(size_t)c * d + i] - sm[i];           // :122-126
  for (int i = 0; i < d; i++) {                                               // :127 forward substitution, column-oriented
    r[i] = r[i] / sL[i + i * d];
    for (int k = i + 1; k < d; k++) r[k] -= r[i] * sL[k + i * d];
  }
  double sq = 0;
  for (int i = 0; i < d; i++) sq += r[i] * r[i];
  out[c] = -0.5 * sq + cst;                                                   // :127-130
}
// Y = L Z + mean (src/mvnorm.cpp:52-60); z injected (d x n) or Philox: normal pair j of sample c -> components 2j, 2j+1
__global__ void rmvnorm_chol_kernel(long long n, int d, const double* mean, const double* L, unsigned long long seed, const double* z, double* out) {
  extern __shared__ __align__(16) unsigned char raw[];
  double* sL = reinterpret_cast<double*>(raw);
  double* sm = sL + d * d;
  for (int i = threadIdx.x; i < d * d; i += blockDim.x) sL[i] = L[i];
  for (int i = threadIdx.x; i < d; i += blockDim.x) sm[i] = mean[i];
  __syncthreads();
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double zz[CPIMH_MVN_MAXD];
  if (z) { for (int i = 0; i < d; i++) zz[i] = z[(size_t)c * d + i]; }
  else {
    Stream s = make_stream(seed, (unsigned long long)(c >> 32) << 32);
    for (int j = 0; 2 * j < d; j++) {
      double ua, ub, z0, z1;
      philox_u2(s, (uint32_t)c, 0, CPIMH_P_INIT, j, ua, ub);
      normal2(ua, ub, z0, z1);
      zz[2 * j] = z0;
      if (2 * j + 1 < d) zz[2 * j + 1] = z1;
    }
  }
  for (int i = 0; i < d; i++) {
    double acc = 0;
    for (int k = 0; k <= i; k++) acc += sL[i + k * d] * zz[k];
    out[(size_t)c * d + i] = acc + sm[i];
  }
}

// host: unblocked Cholesky (Eigen::LLT for small matrices), lower factor, column-major
static int cholesky_lower(const double* cov, int d, std::vector<double>& L) {
  L.assign((size_t)d * d, 0.0);
  for (int k = 0; k < d; k++) {
    double s = cov[k + (size_t)k * d];
    for (int j = 0; j < k; j++) s -= L[k + (size_t)j * d] * L[k + (size_t)j * d];
    CPIMH_REQUIRE(s > 0, "covariance matrix is not positive definite (pivot %d = %g)", k, s);
    const double lkk = sqrt(s);
    L[k + (size_t)k * d] = lkk;
    for (int i = k + 1; i < d; i++) {
      double t = cov[i + (size_t)k * d];
      for (int j = 0; j < k; j++) t -= L[i + (size_t)j * d] * L[k + (size_t)j * d];
      L[i + (size_t)k * d] = t / lkk;
    }
  }
  return CPIMH_OK;
}

// ---------------------------------------------------------------- Lorenz 96 transition (src/lorenz96.cpp:65-93)
template <int D>
__global__ void lorenz_kernel(const double* xin, long long n, double t0, double t1, double h, double F, int integrator, int substeps, double* xout) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[D];
#pragma unroll
  for (int k = 0; k < D; k++) x[k] = xin[(size_t)i * D + k];
  if (integrator == CPIMH_INTEGRATOR_DOPRI5) lorenz_dopri5<D>(F, t0, t1, h, x);
  else {
    const int ns = substeps < 1 ? 1 : substeps;
    const double hh = (t1 - t0) / ns;
    for (int s = 0; s < ns; s++) lorenz_rk4<D>(F, hh, x);
  }
#pragma unroll
  for (int k = 0; k < D; k++) xout[(size_t)i * D + k] = x[k];
}
// wide state (d up to 64 per warp pass): one warp per particle, lane l owns components l and l + 32
__global__ void lorenz_wide_rk4_kernel(const double* xin, long long n, int d, double t0, double t1, double F, int substeps, double* xout);

// ---------------------------------------------------------------- warp-instruction issue peak
// 8 independent chains per thread, fp32 FMA (fma pipe) alternating with integer xor-add (alu pipe): every scheduler can issue
// one warp instruction per clock when its four warps of the eight resident ones always have an independent instruction ready
__global__ void issue_kernel(unsigned* out, int iters) {
  float f0 = threadIdx.x * 1e-9f, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
  unsigned i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
  const float m = 1.0000001f, c = 1e-7f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      f0 = fmaf(f0, m, c); i0 = (i0 ^ 0x9E3779B9u) + i1;
      f1 = fmaf(f1, m, c); i1 = (i1 ^ 0x85EBCA6Bu) + i2;
      f2 = fmaf(f2, m, c); i2 = (i2 ^ 0xC2B2AE35u) + i3;
      f3 = fmaf(f3, m, c); i3 = (i3 ^ 0x27D4EB2Fu) + i0;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = i0 + i1 + i2 + i3 + (unsigned)(f0 + f1 + f2 + f3);
}

// ---------------------------------------------------------------- FP64 FMA peak
__global__ void dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace cpimh

using namespace cpimh;

namespace cpimh {
__global__ void rng_transform_kernel(const double* u, long long n, double* lg, double* sn, double* cs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  lg[i] = rng_log(u[i]);
  double s, c;
  rng_sincos2pi<true>(u[i], s, c);
  sn[i] = s; cs[i] = c;
}
}  // namespace cpimh
extern "C" {

int cpimh_systematic_resampling_n(const double* w, int64_t B, int N, int ndraws, const double* u, int32_t* anc) {
  CPIMH_REQUIRE(w && u && anc, "null argument");
  CPIMH_REQUIRE(B >= 1 && N >= 1 && ndraws >= 1, "B, N and ndraws must be >= 1");
  Scratch s;
  double *dw, *du, *dc, *dl; int* da;
  int rc = s.upload(&dw, w, (size_t)B * N); if (rc) return rc;
  rc = s.upload(&du, u, (size_t)B); if (rc) return rc;
  rc = s.alloc(&dc, (size_t)B * N); if (rc) return rc;
  rc = s.alloc(&dl, (size_t)B * ndraws); if (rc) return rc;
  rc = s.alloc(&da, (size_t)B * ndraws); if (rc) return rc;
  systematic_kernel<<<(unsigned)B, 128>>>(dw, N, ndraws, du, dc, dl, da);
  rc = sync_check(); if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(anc, da, sizeof(int) * (size_t)B * ndraws, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

int cpimh_multinomial_resampling_n(const double* w, int64_t B, int N, int ndraws, const double* e_sorted, const double* u_shuffle, int32_t* anc) {
  CPIMH_REQUIRE(w && e_sorted && anc, "null argument");
  CPIMH_REQUIRE(B >= 1 && N >= 1 && ndraws >= 1, "B, N and ndraws must be >= 1");
  Scratch s;
  double *dw, *de, *dus = nullptr, *dc; int *da, *dt;
  int rc = s.upload(&dw, w, (size_t)B * N); if (rc) return rc;
  rc = s.upload(&de, e_sorted, (size_t)B * ndraws); if (rc) return rc;
  if (u_shuffle && ndraws > 1) { rc = s.upload(&dus, u_shuffle, (size_t)B * (ndraws - 1)); if (rc) return rc; }
  rc = s.alloc(&dc, (size_t)B * N); if (rc) return rc;
  rc = s.alloc(&da, (size_t)B * ndraws); if (rc) return rc;
  rc = s.alloc(&dt, (size_t)B * 2 * ndraws); if (rc) return rc;
  multinomial_kernel<1><<<(unsigned)B, 128>>>(dw, nullptr, N, ndraws, de, dus, dc, dt, da);
  rc = sync_check(); if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(anc, da, sizeof(int) * (size_t)B * ndraws, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

int cpimh_coupled_multinomial_resampling_n(const double* w1, const double* w2, int64_t B, int N, int ndraws,
                                           const double* e_sorted, const double* u_shuffle, int32_t* anc) {
  CPIMH_REQUIRE(w1 && w2 && e_sorted && anc, "null argument");
  CPIMH_REQUIRE(B >= 1 && N >= 1 && ndraws >= 1, "B, N and ndraws must be >= 1");
  Scratch s;
  double *d1, *d2, *de, *dus = nullptr, *dc; int *da, *dt;
  int rc = s.upload(&d1, w1, (size_t)B * N); if (rc) return rc;
  rc = s.upload(&d2, w2, (size_t)B * N); if (rc) return rc;
  rc = s.upload(&de, e_sorted, (size_t)B * ndraws); if (rc) return rc;
  if (u_shuffle && ndraws > 1) { rc = s.upload(&dus, u_shuffle, (size_t)B * (ndraws - 1)); if (rc) return rc; }
  rc = s.alloc(&dc, (size_t)B * 2 * N); if (rc) return rc;
  rc = s.alloc(&da, (size_t)B * 2 * ndraws); if (rc) return rc;
  rc = s.alloc(&dt, (size_t)B * 2 * ndraws); if (rc) return rc;
  multinomial_kernel<2><<<(unsigned)B, 128>>>(d1, d2, N, ndraws, de, dus, dc, dt, da);
  rc = sync_check(); if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(anc, da, sizeof(int) * (size_t)B * 2 * ndraws, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

int cpimh_indexmatching(const double* w1, const double* w2, int64_t B, int N, int ndraws, const double* u_couple,
                        const double* e_mult, const double* us_mult, const double* e_cm, const double* us_cm, int32_t* anc) {
  CPIMH_REQUIRE(w1 && w2 && u_couple && e_mult && e_cm && anc, "null argument");
  CPIMH_REQUIRE(B >= 1 && N >= 1 && ndraws >= 1 && ndraws <= N, "need 1 <= ndraws <= N (src/coupledresamplingindexmatching.cpp:27 reads uniforms(i), i < ndraws)");
  Scratch s;
  double *d1, *d2, *duc, *dem, *decm, *dusm = nullptr, *duscm = nullptr, *mu, *R1, *R2, *dc;
  int *flag, *rank, *nc, *tmp, *da;
  int rc = s.upload(&d1, w1, (size_t)B * N); if (rc) return rc;
  rc = s.upload(&d2, w2, (size_t)B * N); if (rc) return rc;
  rc = s.upload(&duc, u_couple, (size_t)B * N); if (rc) return rc;
  rc = s.upload(&dem, e_mult, (size_t)B * ndraws); if (rc) return rc;
  rc = s.upload(&decm, e_cm, (size_t)B * ndraws); if (rc) return rc;
  if (us_mult && ndraws > 1) { rc = s.upload(&dusm, us_mult, (size_t)B * (ndraws - 1)); if (rc) return rc; }
  if (us_cm && ndraws > 1) { rc = s.upload(&duscm, us_cm, (size_t)B * (ndraws - 1)); if (rc) return rc; }
  rc = s.alloc(&mu, (size_t)B * N); if (rc) return rc;
  rc = s.alloc(&R1, (size_t)B * N); if (rc) return rc;
  rc = s.alloc(&R2, (size_t)B * N); if (rc) return rc;
  rc = s.alloc(&dc, (size_t)B * 3 * N); if (rc) return rc;
  rc = s.alloc(&flag, (size_t)B * ndraws); if (rc) return rc;
  rc = s.alloc(&rank, (size_t)B * ndraws); if (rc) return rc;
  rc = s.alloc(&nc, (size_t)B); if (rc) return rc;
  rc = s.alloc(&tmp, (size_t)B * 5 * ndraws); if (rc) return rc;
  rc = s.alloc(&da, (size_t)B * 2 * ndraws); if (rc) return rc;
  indexmatching_prepare<<<(unsigned)B, 128>>>(d1, d2, N, ndraws, duc, mu, R1, R2, flag, rank, nc);
  indexmatching_draw<<<(unsigned)B, 128>>>(mu, R1, R2, N, ndraws, nc, dem, dusm, decm, duscm, flag, rank, dc, tmp, da);
  rc = sync_check(); if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(anc, da, sizeof(int) * (size_t)B * 2 * ndraws, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

int cpimh_sorted_uniforms(const double* u_raw, int64_t B, int n, double* e_sorted) {
  CPIMH_REQUIRE(u_raw && e_sorted, "null argument");
  CPIMH_REQUIRE(B >= 1 && n >= 1 && n <= 24576, "need 1 <= n <= 24576");
  Scratch s;
  double *du, *de;
  int rc = s.upload(&du, u_raw, (size_t)B * n); if (rc) return rc;
  rc = s.alloc(&de, (size_t)B * n); if (rc) return rc;
  const size_t smem = sizeof(double) * (size_t)n;
  CPIMH_CUDA_CHECK(cudaFuncSetAttribute(sorted_uniforms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sorted_uniforms_kernel<<<(unsigned)B, 256, smem>>>(du, n, de);
  rc = sync_check(); if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(e_sorted, de, sizeof(double) * (size_t)B * n, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

int cpimh_normalize_weight(const double* logw, int64_t B, int N, double* w, double* logz) {
  CPIMH_REQUIRE(logw && w, "null argument");
  CPIMH_REQUIRE(B >= 1 && N >= 1, "B and N must be >= 1");
  Scratch s;
  double *dl, *dw, *dz;
  int rc = s.upload(&dl, logw, (size_t)B * N); if (rc) return rc;
  rc = s.alloc(&dw, (size_t)B * N); if (rc) return rc;
  rc = s.alloc(&dz, (size_t)B); if (rc) return rc;
  normalize_kernel<<<(unsigned)B, 256>>>(dl, N, dw, dz);
  rc = sync_check(); if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(w, dw, sizeof(double) * (size_t)B * N, cudaMemcpyDeviceToHost));
  if (logz) CPIMH_CUDA_CHECK(cudaMemcpy(logz, dz, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

int cpimh_dmvnorm_transpose_cholesky(const double* x, int d, int64_t n, const double* mean, const double* chol, double* out) {
  CPIMH_REQUIRE(x && mean && chol && out, "null argument");
  CPIMH_REQUIRE(d >= 1 && d <= CPIMH_MVN_MAXD && n >= 1, "need 1 <= d <= %d and n >= 1", CPIMH_MVN_MAXD);
  double hl = 0;
  for (int i = 0; i < d; i++) hl += log(chol[i + (size_t)i * d]);            // src/mvnorm.cpp:120
  const double cst = -(hl) - (d * CPIMH_HALF_LOG_2PI_TRUNC);                 // :121
  Scratch s;
  double *dx, *dm, *dL, *dout;
  int rc = s.upload(&dx, x, (size_t)n * d); if (rc) return rc;
  rc = s.upload(&dm, mean, (size_t)d); if (rc) return rc;
  rc = s.upload(&dL, chol, (size_t)d * d); if (rc) return rc;
  rc = s.alloc(&dout, (size_t)n); if (rc) return rc;
  const int threads = 128;
  dmvnorm_chol_kernel<<<(unsigned)((n + threads - 1) / threads), threads, sizeof(double) * (size_t)(d * d + d)>>>(dx, d, n, dm, dL, cst, dout);
  rc = sync_check(); if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(out, dout, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}
int cpimh_dmvnorm_transpose(const double* x, int d, int64_t n, const double* mean, const double* cov, double* out) {
  CPIMH_REQUIRE(cov, "null argument");
  std::vector<double> L;
  int rc = cholesky_lower(cov, d, L);                                        // src/mvnorm.cpp:93-94
  if (rc) return rc;
  return cpimh_dmvnorm_transpose_cholesky(x, d, n, mean, L.data(), out);
}
int cpimh_dmvnorm(const double* x, int64_t n, int d, const double* mean, const double* cov, double* out) {
  CPIMH_REQUIRE(x && cov, "null argument");
  std::vector<double> xt((size_t)n * d);                                     // src/mvnorm.cpp:79 solve(xcentered.transpose())
  for (int64_t i = 0; i < n; i++) for (int j = 0; j < d; j++) xt[(size_t)i * d + j] = x[(size_t)i + (size_t)j * n];
  return cpimh_dmvnorm_transpose(xt.data(), d, n, mean, cov, out);
}

int cpimh_rmvnorm_transpose_cholesky(int64_t n, int d, const double* mean, const double* chol, uint64_t seed, const double* z, double* out) {
  CPIMH_REQUIRE(mean && chol && out, "null argument");
  CPIMH_REQUIRE(d >= 1 && d <= CPIMH_MVN_MAXD && n >= 1, "need 1 <= d <= %d and n >= 1", CPIMH_MVN_MAXD);
  Scratch s;
  double *dm, *dL, *dz = nullptr, *dout;
  int rc = s.upload(&dm, mean, (size_t)d); if (rc) return rc;
  rc = s.upload(&dL, chol, (size_t)d * d); if (rc) return rc;
  if (z) { rc = s.upload(&dz, z, (size_t)n * d); if (rc) return rc; }
  rc = s.alloc(&dout, (size_t)n * d); if (rc) return rc;
  const int threads = 128;
  rmvnorm_chol_kernel<<<(unsigned)((n + threads - 1) / threads), threads, sizeof(double) * (size_t)(d * d + d)>>>(n, d, dm, dL, seed, dz, dout);
  rc = sync_check(); if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(out, dout, sizeof(double) * (size_t)n * d, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}
int cpimh_rmvnorm_transpose(int64_t n, int d, const double* mean, const double* cov, uint64_t seed, const double* z, double* out) {
  CPIMH_REQUIRE(cov, "null argument");
  std::vector<double> L;
  int rc = cholesky_lower(cov, d, L);                                        // src/mvnorm.cpp:30-31
  if (rc) return rc;
  return cpimh_rmvnorm_transpose_cholesky(n, d, mean, L.data(), seed, z, out);
}
int cpimh_rmvnorm(int64_t n, int d, const double* mean, const double* cov, uint64_t seed, const double* z, double* out) {
  // src/mvnorm.cpp:6-22: Y (n x d) filled by column, Y = Y U with U = L^T, + mean(j)  ==  transpose of (L Z^T + mean)
  CPIMH_REQUIRE(out, "null argument");
  std::vector<double> zt, yt((size_t)n * d);
  if (z) { zt.resize((size_t)n * d); for (int64_t i = 0; i < n; i++) for (int j = 0; j < d; j++) zt[(size_t)i * d + j] = z[(size_t)i + (size_t)j * n]; }
  int rc = cpimh_rmvnorm_transpose(n, d, mean, cov, seed, z ? zt.data() : nullptr, yt.data());
  if (rc) return rc;
  for (int64_t i = 0; i < n; i++) for (int j = 0; j < d; j++) out[(size_t)i + (size_t)j * n] = yt[(size_t)i * d + j];
  return CPIMH_OK;
}

int cpimh_create_A(double alpha, int d, double* A) {
  CPIMH_REQUIRE(A && d >= 1, "bad argument");
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) A[i + (size_t)j * d] = pow(alpha, 1 + abs(i - j));   // src/create_A_.cpp:10
  return CPIMH_OK;
}

int cpimh_lorenz_transition(const double* x_in, int d, int64_t n, double tstart, double tend, double h, double F,
                            int integrator, int substeps, double* x_out) {
  CPIMH_REQUIRE(x_in && x_out && n >= 1, "bad argument");
  Scratch s;
  double *dx, *dy;
  int rc = s.upload(&dx, x_in, (size_t)n * d); if (rc) return rc;
  rc = s.alloc(&dy, (size_t)n * d); if (rc) return rc;
  const int threads = 128; const unsigned grid = (unsigned)((n + threads - 1) / threads);
  switch (d) {
    case 4: lorenz_kernel<4><<<grid, threads>>>(dx, n, tstart, tend, h, F, integrator, substeps, dy); break;
    case 8: lorenz_kernel<8><<<grid, threads>>>(dx, n, tstart, tend, h, F, integrator, substeps, dy); break;
    case 16: lorenz_kernel<16><<<grid, threads>>>(dx, n, tstart, tend, h, F, integrator, substeps, dy); break;
    default:
      CPIMH_REQUIRE((d == 32 || d == 64) && integrator == CPIMH_INTEGRATOR_RK4, "Lorenz 96: d in {4, 8, 16} (rk4 or dopri5) or d in {32, 64} with rk4; got d=%d integrator=%d", d, integrator);
      lorenz_wide_rk4_kernel<<<(unsigned)((n * 32 + 127) / 128), 128>>>(dx, n, d, tstart, tend, F, substeps, dy);
  }
  rc = sync_check(); if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(x_out, dy, sizeof(double) * (size_t)n * d, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

int cpimh_measure_fp64_peak(double* tflops) {
  CPIMH_REQUIRE(tflops, "null argument");
  int dev = 0; cudaDeviceProp prop;
  CPIMH_CUDA_CHECK(cudaGetDevice(&dev));
  CPIMH_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  const int threads = 256, grid = prop.multiProcessorCount * 8, iters = 1 << 15;
  double* out;
  CPIMH_CUDA_CHECK(cudaMalloc(&out, sizeof(double) * (size_t)threads * grid));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma_kernel<<<grid, threads>>>(out, 1024);
  double best = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    dfma_kernel<<<grid, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 8.0 * (double)iters * threads * grid / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  CPIMH_CUDA_CHECK(cudaGetLastError());
  *tflops = best;
  return CPIMH_OK;
}

int cpimh_measure_issue_peak(double* gwarp) {
  CPIMH_REQUIRE(gwarp, "null argument");
  int dev = 0; cudaDeviceProp prop;
  CPIMH_CUDA_CHECK(cudaGetDevice(&dev));
  CPIMH_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  const int threads = 256, grid = prop.multiProcessorCount * 4, iters = 1 << 12;
  unsigned* out;
  CPIMH_CUDA_CHECK(cudaMalloc(&out, sizeof(unsigned) * (size_t)threads * grid));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  issue_kernel<<<grid, threads>>>(out, 64);
  double best = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    issue_kernel<<<grid, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    // per thread and iteration: 16 x (4 FFMA + 4 LOP3 + 4 IADD) = 192 instructions (xor-add may fuse to fewer: a lower bound)
    const double winst = 192.0 * (double)iters * (threads / 32) * grid;
    const double g = winst / (ms * 1e-3) / 1e9;
    if (g > best) best = g;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  CPIMH_CUDA_CHECK(cudaGetLastError());
  *gwarp = best;
  return CPIMH_OK;
}

int cpimh_rng_transform(const double* u, int64_t n, double* lg, double* sn, double* cs) {
  CPIMH_REQUIRE(u && lg && sn && cs && n >= 1, "null argument");
  DevBuf du, dl, ds, dc;
  struct Rel { DevBuf* b[4]; ~Rel() { for (auto x : b) x->release(); } } rel{{&du, &dl, &ds, &dc}};
  int rc = du.alloc(sizeof(double) * n);
  if (!rc) rc = dl.alloc(sizeof(double) * n);
  if (!rc) rc = ds.alloc(sizeof(double) * n);
  if (!rc) rc = dc.alloc(sizeof(double) * n);
  if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(du.p, u, sizeof(double) * n, cudaMemcpyHostToDevice));
  rng_transform_kernel<<<(unsigned)((n + 255) / 256), 256>>>(du.as<double>(), n, dl.as<double>(), ds.as<double>(), dc.as<double>());
  CPIMH_CUDA_CHECK(cudaGetLastError());
  CPIMH_CUDA_CHECK(cudaMemcpy(lg, dl.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  CPIMH_CUDA_CHECK(cudaMemcpy(sn, ds.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  CPIMH_CUDA_CHECK(cudaMemcpy(cs, dc.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

}  // extern "C"
