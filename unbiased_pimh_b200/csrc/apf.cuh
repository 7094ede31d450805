// apf.cuh -- pf(y, theta, model, nparticles, ess_threshold, systematic) (R/CPF.R:106-155): the
// bootstrap filter with ESS-triggered resampling that the PMMH scripts use as likelihood evaluator.  One CTA per filter.
// A step is propagate -> weigh with the carried log-weights -> log-likelihood increment, normalised weights, filtering mean
// -> resample iff (1 / sum(w^2)) / N < ess_threshold (sorted multinomial draws or systematic) -> store the RESAMPLED
// particles and their ancestors (Tree$update, :149).  Randomness and its order: oracle/apf.c header, DESIGN.md "RNG".
//
// Layout: hist_x[ws][t][q][i] holds the resampled particles of every time (the path store), hist_a[ws][t][i] their
// ancestors, prop[ws][q][i] the propagated particles of the current step (gather source).  Shared memory: weights / CDF,
// ladder / log-weights, carried log-weights.
#pragma once
#include "common.cuh"
#include "models.cuh"

namespace cpimh {

struct ApfParams {
  int N, T, dy, Npad;
  long long B;
  unsigned long long seed, first_filter;
  double ess_threshold;
  int systematic;
  const double* obs; const double* theta; const double* pre;
  int npre, integrator, substeps, sk_M;
  double* hist_x;              // [ws][T+1][D][Npad]
  int* hist_a;                 // [ws][T][Npad]
  double* prop;                // [ws][D][Npad]
  double* loglik;              // [B]
  double* loglik_t;            // [B][T] cumulative, or null
  double* pf_means;            // [B][T][D] or null
  double* path;                // [B][T+1][D]
  int* path_index;             // [B] or null
  int* resampled;              // [B][T] or null
  int* anc_rec;                // [B][T][N] or null
  int* status;                 // [B]
  // exact ancestors (DESIGN.md "Exact ancestors"): a step whose uniform lies within margin_rel * sum(w) of a CDF knot is repeated
  // inside the kernel with cumsum(normweights) by one thread, left to right; exact = 1 takes that order from the start (test aid);
  // near[b] counts the repeated steps of filter b
  int exact; int* near; double margin_rel;
};
struct ApfSmem { double* C; double* S; double* LW; double* theta; double* pre; double* dscr; int* iscr; int theta_off, pre_off; };
__host__ __device__ inline size_t apf_smem_bytes(int Npad, int npre) {
  size_t b = (size_t)(2 * padded_len(Npad + 16) + Npad) * 8 + (size_t)(CPIMH_MAX_THETA + npre + (npre & 1) + 64) * 8 + 36 * 4;
  return (b + 15) & ~(size_t)15;
}
__device__ inline ApfSmem apf_carve(unsigned char* raw, int Npad, int npre) {
  ApfSmem s;
  double* d = reinterpret_cast<double*>(raw);
  s.C = d; d += padded_len(Npad + 16);       // C and S in the padded layout (element i at pad16(i), common.cuh)
  s.S = d; d += padded_len(Npad + 16);
  s.LW = d; d += Npad;
  s.theta = d; s.theta_off = (int)(reinterpret_cast<unsigned char*>(d) - raw); d += CPIMH_MAX_THETA;
  s.pre = d; s.pre_off = (int)(reinterpret_cast<unsigned char*>(d) - raw); d += npre + (npre & 1);
  s.dscr = d; d += 64;
  s.iscr = reinterpret_cast<int*>(d);
  return s;
}

template <class Model>
__device__ void run_apf(const ApfParams& p, const ApfSmem& sm, long long ws, const Stream& stream, long long b) {
  constexpr int D = Model::D;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int N = p.N, T = p.T, Npad = p.Npad;
  const size_t xstride = (size_t)D * Npad;
  double* hx = p.hist_x + (size_t)ws * (T + 1) * xstride;
  int* ha = p.hist_a + (size_t)ws * T * Npad;
  double* prop = p.prop + (size_t)ws * xstride;
  ModelCtx c;
  c.theta.off = sm.theta_off; c.pre.off = sm.pre_off; c.stream = stream; c.N = N; c.T = T;
  c.inj_init = nullptr; c.inj_trans = nullptr;
  c.n_trans = 0; c.sk_M = p.sk_M; c.integrator = p.integrator; c.substeps = p.substeps;
  int st = 0;
  if (tid == 0) sm.iscr[32] = 0;
  bool near = false;
  int repeats = 0;
  bool cdf_exact = false;                              // whether sm.C currently holds the reference-order CDF
  // CDF of the normalised weights: parallel scan, or (ex) the reference's running sum (src/multinomial.cpp:17, src/systematic.cpp:13-19)
  auto weight_cdf = [&](bool ex) {
    if (ex) { __syncthreads(); if (tid == 0) sequential_cumsum_p(sm.C, N); __syncthreads(); }
    else block_scan_chunked<false>(sm.C, nullptr, N, sm.dscr);
    cdf_exact = ex;
  };
  // a flagged pass: put the normalised weights (stashed in LW, which is dead while a step resamples) back
  auto restore_weights = [&]() {
    __syncthreads();
    for (int k = tid; k < N; k += NT) sm.C[pad16(k)] = sm.LW[k];
    repeats++;
  };
  // the two knots that decided ancestor a (rule `u < cdf[j-1]`, src/multinomial.cpp:25-27) against u +- the rounding margin
  auto knots_le = [&](int a, double u) {
    const double mar = p.margin_rel * sm.C[pad16(N - 1)];
    const double cur = a < N ? sm.C[pad16(a)] : INFINITY, below = a > 0 ? sm.C[pad16(a - 1)] : -INFINITY;
    if (cur - u <= mar || u - below < mar) near = true;
  };
  // rule `while (csw < u)` (src/systematic.cpp:15)
  auto knots_lt = [&](int a, double u) {
    const double mar = p.margin_rel * sm.C[pad16(N - 1)];
    const double cur = a < N ? sm.C[pad16(a)] : INFINITY, below = a > 0 ? sm.C[pad16(a - 1)] : -INFINITY;
    if (cur - u < mar || u - below <= mar) near = true;
  };
  const double lw0 = log(1.0 / (double)N);
  for (int i = tid; i < N; i += NT) {                    // R/CPF.R:113-119
    double x[D];
    Model::init(c, i, x);
#pragma unroll
    for (int q = 0; q < D; q++) hx[(size_t)q * Npad + i] = x[q];
    sm.LW[i] = lw0;
  }
  double cum = 0.0;
  __syncthreads();

  for (int t = 1; t <= T; t++) {
    const double* y = p.obs + (size_t)(t - 1) * p.dy;
    const double sc = Model::step_const(c, t);
    const double* xprev = hx + (size_t)(t - 1) * xstride;
    double* xres = hx + (size_t)t * xstride;
    int* arow = ha + (size_t)(t - 1) * Npad;
    // ---- propagate and weigh (:124-128)
    double m = -INFINITY;
    for (int k = tid; k < N; k += NT) {
      double x[D];
#pragma unroll
      for (int q = 0; q < D; q++) x[q] = xprev[(size_t)q * Npad + k];
      uint2 ush = make_uint2(0u, 0u);
      if (Model::USES_SHARED_U) { double ua; philox_u1raw(stream, k, t, CPIMH_P_RESAMPLE, 0, ua, ush); }
      st |= Model::transition(c, t, k, x, sc, ush);
#pragma unroll
      for (int q = 0; q < D; q++) prop[(size_t)q * Npad + k] = x[q];
      const double lw = Model::logw(c, x, y) + sm.LW[k];
      sm.S[pad16(k)] = lw;
      m = fmax(m, lw);
    }
    m = block_max(m, sm.dscr);
    double s = 0.0;
    for (int k = tid; k < N; k += NT) { const double w = exp(sm.S[pad16(k)] - m); sm.C[pad16(k)] = w; s += w; }
    s = block_sum(s, sm.dscr);
    cum += m + log(s);                                   // :130
    if (tid == 0 && p.loglik_t) p.loglik_t[(size_t)b * T + (t - 1)] = cum;
    double s2 = 0.0;
    for (int k = tid; k < N; k += NT) { const double w = sm.C[pad16(k)] / s; sm.C[pad16(k)] = w; s2 += w * w; }   // :131
    s2 = block_sum(s2, sm.dscr);
    if (p.pf_means) {                                    // :133
#pragma unroll
      for (int q = 0; q < D; q++) {
        double a = 0.0;
        for (int k = tid; k < N; k += NT) a += sm.C[pad16(k)] * prop[(size_t)q * Npad + k];
        a = block_sum(a, sm.dscr);
        if (tid == 0) p.pf_means[((size_t)b * T + (t - 1)) * D + q] = a;
      }
    }
    const bool resample = ((1.0 / s2) / (double)N) < p.ess_threshold;     // :135
    if (tid == 0 && p.resampled) p.resampled[(size_t)b * T + (t - 1)] = resample ? 1 : 0;
    // ---- resample (:136-147)
    if (resample) {
      for (int k = tid; k < N; k += NT) sm.LW[k] = sm.C[pad16(k)];          // stash: LW is reset below
      bool ex = p.exact != 0;
      for (;;) {
        near = false;
        if (!p.systematic) {
          for (int i = tid; i <= N; i += NT) sm.S[pad16(i)] = -rng_log(philox_u1(stream, i, t, CPIMH_P_RESAMPLE, 0));
          __syncthreads();
          weight_cdf(ex);
          block_scan_chunked<false>(sm.S, nullptr, N + 1, sm.dscr);
          const double scale = sm.C[pad16(N - 1)] / sm.S[pad16(N)];
          for (int k = tid; k < N; k += NT) {
            const double u = scale * sm.S[pad16(k)];
            int a = count_le_p(sm.C, 0, N, u);
            if (!ex) knots_le(a, u);
            if (a >= N) { a = N - 1; st |= 512; }
            arow[k] = a;
          }
        } else {                                         // systematic_resampling_n(normweights, N, u): src/systematic.cpp:7-32
          __syncthreads();
          weight_cdf(ex);
          const double u = philox_u1(stream, 0, t, CPIMH_P_AS, 0) / (double)N;
          for (int k = tid; k < N; k += NT) {
            const double uk = u + (double)k * (1.0 / (double)N);
            int a = count_lt_p(sm.C, N, uk);
            if (!ex) knots_lt(a, uk);
            if (a >= N) { a = N - 1; st |= 512; }
            arow[k] = a;
          }
        }
        if (!ex && __syncthreads_or(near ? 1 : 0)) { ex = true; restore_weights(); continue; }   // repeat this step in reference order
        break;
      }
      __syncthreads();
      if (t < T) for (int k = tid; k < N; k += NT) sm.LW[k] = lw0;          // t == T: the stash serves the final draw
    } else {
      for (int k = tid; k < N; k += NT) { arow[k] = k; sm.LW[k] = (t < T) ? log(sm.C[pad16(k)]) : sm.C[pad16(k)]; }
      __syncthreads();
      if (t == T) weight_cdf(p.exact != 0);                                  // the final draw needs the CDF
    }
    // ---- the tree keeps the resampled particles (:148-149)
    for (int k = tid; k < N; k += NT) {
      const int a = arow[k];
#pragma unroll
      for (int q = 0; q < D; q++) xres[(size_t)q * Npad + k] = prop[(size_t)q * Npad + a];
      if (p.anc_rec) p.anc_rec[((size_t)b * T + (t - 1)) * N + k] = a;
    }
    __syncthreads();
  }
  // ---- sample(1:N, 1, prob = normweights) (:152) by inversion of the last CDF, then Tree$get_path
  if (st) atomicOr(&sm.iscr[32], st);
  __syncthreads();
  const double upath = philox_u1(stream, 0, T + 1, CPIMH_P_PATH, 0);
  int kidx = count_lt_p(sm.C, N, upath);
  near = false;
  if (!cdf_exact) knots_lt(kidx, upath);
  if (!cdf_exact && __syncthreads_or(near ? 1 : 0)) {    // the final draw in reference order
    restore_weights();
    weight_cdf(true);
    kidx = count_lt_p(sm.C, N, upath);
  }
  if (kidx >= N) kidx = N - 1;
  if (tid == 0 && p.near) p.near[b] = repeats;
  const int flags = sm.iscr[32];
  if (tid == 0) {
    p.loglik[b] = cum;
    if (p.path_index) p.path_index[b] = kidx;
    p.status[b] = flags & 255;
  }
  if (tid < 32) {
    double* tr = p.path + (size_t)b * (T + 1) * D;
    int j = kidx;
    for (int step = T; step >= 0; step--) {
      for (int q = tid; q < D; q += 32) tr[(size_t)step * D + q] = hx[(size_t)step * xstride + (size_t)q * Npad + j];
      if (step > 0) j = ha[(size_t)(step - 1) * Npad + j];
    }
  }
  __syncthreads();
}

template <class Model>
__global__ void __launch_bounds__(512, 1) apf_kernel(ApfParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ApfSmem sm = apf_carve(smem_raw, p.Npad, p.npre);
  for (int i = threadIdx.x; i < CPIMH_MAX_THETA; i += blockDim.x) sm.theta[i] = p.theta[i];
  for (int i = threadIdx.x; i < p.npre; i += blockDim.x) sm.pre[i] = p.pre[i];
  __syncthreads();
  for (long long b = blockIdx.x; b < p.B; b += gridDim.x)
    run_apf<Model>(p, sm, blockIdx.x, make_stream(p.seed, p.first_filter + (unsigned long long)b), b);
}

struct ApfLaunch { int grid, threads; size_t smem; };
template <class Model>
int launch_apf(const ApfParams& p, const ApfLaunch& L, cudaStream_t stream) {
  CPIMH_CUDA_CHECK(cudaFuncSetAttribute(apf_kernel<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  apf_kernel<Model><<<L.grid, L.threads, L.smem, stream>>>(p);
  CPIMH_CUDA_CHECK(cudaGetLastError());
  return CPIMH_OK;
}
template <class Model>
int occupancy_apf(int threads, size_t smem, int* blocks_per_sm) {
  CPIMH_CUDA_CHECK(cudaFuncSetAttribute(apf_kernel<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CPIMH_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, apf_kernel<Model>, threads, smem));
  return CPIMH_OK;
}

}  // namespace cpimh
