// chains.cuh -- coupled PIMH chains + unbiased (H_bar) estimator, one persistent CTA per replicate (strategy 1 of the
// chain driver; the default round-based strategy with batched speculative proposals lives in api_chains.cu).
//
// Restates, batched on the GPU, the reference's R-level drivers:
//   pimh_init / single_kernel / coupled_kernel   R/pf_kernels.R:86-145   (one CPF proposal, one uniform,
//                                                accept iff log(u) < logz_prop - logz_i, same u for both chains)
//   coupled_pimh_chains                          R/debiasing_algorithm.R:140-265 (lag-1 coupling, stop at max(tau, K))
//   H_bar(c_chains, h = identity, k, K)          R/debiasing_algorithm.R:284-316
// PIMH proposals do not depend on the chain state, so a replicate is: init filter, then one filter per iteration
// with an O(1) accept/meet recursion.  A CTA keeps running replicates fetched from an atomic counter, so short
// (tau = 1) and long replicates balance across the 148 SMs without host round trips.  The second filter that
// pimh_init runs and coupled_pimh_chains discards (R/pf_kernels.R:140-142 vs debiasing_algorithm.R:160-161) is
// NOT run; `nfilters` counts filters actually run.
#pragma once
#include "pf_fused.cuh"

namespace cpimh {

struct ChainParams {
  PfParams pf;                 // workspace slot = blockIdx.x; per-filter outputs unused
  long long R;
  unsigned first_replicate;
  int k, K, max_iterations;    // max_iterations < 0: unbounded (R: Inf)
  double* state;               // [grid][3][T+1][D]   chain 1, chain 2, proposal (rotating)
  double* acc;                 // [grid][2][T+1][D]   MCMC sum, bias-correction sum (all state components)
  double* scal;                // [grid][4]           scratch scalars written by run_filter (logz)
  int* iscal;                  // [grid][4]
  double* hbar;                // [R][T+1][D] or null (R layout d x (T+1) column-major per replicate)
  double* bc;                  // [R][T+1][D] or null: bias-correction term alone (BC(), R/debiasing_algorithm.R:330-363)
  int* meetingtime;            // [R]
  int* iteration;              // [R]
  int* status;                 // [R]
  unsigned long long* nfilters;// [1]
  int* work_counter;           // [1]
  double* samples1; double* samples2; double* log_pdf1; double* log_pdf2;   // recording (nullable)
  int max_rec;
};

template <class Model, int NT_MAX, int MIN_BLOCKS>
__global__ void __launch_bounds__(NT_MAX, MIN_BLOCKS) chains_kernel(ChainParams cp) {
  constexpr int D = Model::D;
  const PfParams& p = cp.pf;
  PfSmem sm = pf_carve(p.lay);
  const int tid = threadIdx.x, NT = blockDim.x, P = p.T + 1;
  for (int i = tid; i < CPIMH_MAX_THETA; i += NT) sm.theta[i] = p.theta[i];
  for (int i = tid; i < p.npre; i += NT) sm.pre[i] = p.pre[i];
  __syncthreads();
  const long long ws = blockIdx.x;
  double* sbuf = cp.state + (size_t)ws * 3 * P * D;
  const int PD = P * D;
  double* accH = cp.acc + (size_t)ws * 2 * PD;
  double* accD = accH + PD;
  double* lzbuf = cp.scal + (size_t)ws * 4;
  int* ibuf = cp.iscal + (size_t)ws * 4;

  for (;;) {
    if (tid == 0) sm.iscr[33] = atomicAdd(cp.work_counter, 1);
    __syncthreads();
    const long long r = sm.iscr[33];
    __syncthreads();
    if (r >= cp.R) break;
    const unsigned rep = cp.first_replicate + (unsigned)r;
    int p1 = 0, p2 = -1, pp = 1;           // buffer indices: chain 1, chain 2 (none yet), next proposal
    int id1 = 0, id2 = -1;                  // iteration index of the proposal each chain currently holds
    FilterOut out;
    out.logz = lzbuf; out.status = ibuf;
    // pimh_init (R/pf_kernels.R:134-139)
    out.traj = sbuf + (size_t)p1 * P * D;
    run_filter<Model>(p, sm, ws, make_stream(p.seed, (unsigned long long)rep), -1, out);
    int status = ibuf[0];
    double lz1 = lzbuf[0], lz2 = -INFINITY;
    const double* s1 = sbuf + (size_t)p1 * P * D;
    for (int j = tid; j < PD; j += NT) { accH[j] = (cp.k <= 0) ? s1[j] : 0.0; accD[j] = 0.0; }
    if (cp.samples1) {
      for (int j = tid; j < P; j += NT) cp.samples1[((size_t)r * (cp.max_rec + 1)) * P + j] = s1[(size_t)j * D];
      if (tid == 0) cp.log_pdf1[(size_t)r * (cp.max_rec + 1)] = lz1;
    }
    int iter = 0, meet = 0, tau = -1, finished = 0;
    while (!finished && status == 0 && (cp.max_iterations < 0 || iter < cp.max_iterations)) {   // R/debiasing_algorithm.R:180
      iter++;
      __syncthreads();
      out.traj = sbuf + (size_t)pp * P * D;
      Stream s = make_stream(p.seed, (unsigned long long)rep | ((unsigned long long)iter << 32));
      run_filter<Model>(p, sm, ws, s, -1, out);
      status = ibuf[0];
      const double lzp = lzbuf[0];
      if (status == 0 && isnan(lzp)) status = CPIMH_ERR_INVALID;          // all weights zero / NaN: no valid proposal
      if (status) break;
      const double lu = log(philox_u1(s, 0, 0, CPIMH_P_MH, 0));          // R/pf_kernels.R:112
      const bool a1 = lu < (lzp - lz1);                                   // :113
      const bool a2 = meet ? a1 : (lu < (lzp - lz2));                     // :121 (lz2 = -Inf at iter 1: always accepts)
      if (a1) { p1 = pp; lz1 = lzp; id1 = iter; }
      if (meet) { p2 = p1; lz2 = lz1; id2 = id1; }                        // single_kernel: both chains move together
      else {
        if (a2) { p2 = pp; lz2 = lzp; id2 = iter; }
        if (id1 == id2) { meet = 1; tau = iter; }                         // debiasing_algorithm.R:216-220
      }
      pp = 0; while (pp == p1 || pp == p2) pp++;                          // a buffer neither chain references
      const double* x1 = sbuf + (size_t)p1 * P * D;
      const double* x2 = sbuf + (size_t)p2 * P * D;
      // H_bar on the fly (debiasing_algorithm.R:296-315): MCMC term for k <= iter <= K,
      // bias term t = iter-1 >= k with coefficient min(t-k+1, K-k+1) (zero once the chains have met)
      const int t = iter - 1;
      const bool mcmc = (iter >= cp.k && iter <= cp.K);
      const bool bias = (t >= cp.k) && (id1 != id2);
      const double coef = (double)min(t - cp.k + 1, cp.K - cp.k + 1);
      for (int j = tid; j < PD; j += NT) {
        double xa = x1[j], xb = x2[j];
        if (mcmc) accH[j] += xa;
        if (bias) accD[j] = accD[j] + coef * (xa - xb);
      }
      if (cp.samples1 && iter <= cp.max_rec) {                            // :235-236 chain_state[1,]: first component only
        for (int j = tid; j < P; j += NT) {
          cp.samples1[((size_t)r * (cp.max_rec + 1) + iter) * P + j] = x1[(size_t)j * D];
          cp.samples2[((size_t)r * cp.max_rec + (iter - 1)) * P + j] = x2[(size_t)j * D];
        }
      }
      if (cp.samples1 && tid == 0 && iter <= cp.max_rec) {
        cp.log_pdf1[(size_t)r * (cp.max_rec + 1) + iter] = lz1;
        cp.log_pdf2[(size_t)r * cp.max_rec + (iter - 1)] = lz2;
      }
      if (tau > 0 && iter >= max(tau, cp.K)) finished = 1;                // :245
    }
    __syncthreads();
    const double denom = (double)(cp.K - cp.k + 1);
    if (cp.hbar) for (int j = tid; j < PD; j += NT) cp.hbar[(size_t)r * PD + j] = (accH[j] + accD[j]) / denom;
    if (cp.bc) for (int j = tid; j < PD; j += NT) cp.bc[(size_t)r * PD + j] = accD[j] / denom;
    if (tid == 0) {
      cp.meetingtime[r] = tau;
      cp.iteration[r] = iter;
      cp.status[r] = status;
      atomicAdd(cp.nfilters, (unsigned long long)(1 + iter));
    }
    __syncthreads();
  }
}

template <class Model, class F>
int chains_with_kernel(int variant, F&& f) {
  switch (variant) {
    case PF_VARIANT_128x8: return f(chains_kernel<Model, 128, (Model::D > 4 ? 5 : 6)>);
    case PF_VARIANT_256x3: return f(chains_kernel<Model, 256, (Model::D > 4 ? 2 : 3)>);
    case PF_VARIANT_512x2: return f(chains_kernel<Model, 512, (Model::D > 4 ? 1 : 2)>);
    default: return f(chains_kernel<Model, 1024, 1>);
  }
}
template <class Model>
int launch_chains(const ChainParams& cp, const PfLaunch& L, cudaStream_t stream) {
  return chains_with_kernel<Model>(L.variant, [&](auto kernel) -> int {
    CPIMH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    kernel<<<L.grid, L.threads, L.smem, stream>>>(cp);
    CPIMH_CUDA_CHECK(cudaGetLastError());
    return CPIMH_OK;
  });
}
template <class Model>
int occupancy_chains(int variant, int threads, size_t smem, int* blocks_per_sm) {
  return chains_with_kernel<Model>(variant, [&](auto kernel) -> int {
    CPIMH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPIMH_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, threads, smem));
    return CPIMH_OK;
  });
}

// ------------------------------------------------------------------ single model operations (primitive API)
struct ModelOpParams {
  int op;                      // 0 rinit, 1 rtransition, 2 dmeasurement
  int N, T, time;
  const double* theta; const double* pre; int npre;
  unsigned long long seed, filter_id;
  const double* noise;         // [n][N] component-major (device) or null
  int n_trans, sk_M, integrator, substeps;
  double* x;                   // d x N column-major (R layout), device
  const double* y;             // [dim_y]
  double* logw;                // [N]
  int* status;                 // [1]
};
template <class Model>
__global__ void model_op_kernel(ModelOpParams q) {
  constexpr int D = Model::D;
  ModelCtx c;                                  // dynamic window: theta [CPIMH_MAX_THETA] then pre [npre]
  c.theta.off = 0; c.pre.off = CPIMH_MAX_THETA * (int)sizeof(double);
  for (int i = threadIdx.x; i < CPIMH_MAX_THETA; i += blockDim.x) c.theta[i] = q.theta[i];
  for (int i = threadIdx.x; i < q.npre; i += blockDim.x) c.pre[i] = q.pre[i];
  __syncthreads();
  c.stream = make_stream(q.seed, q.filter_id); c.N = q.N; c.T = q.T;
  c.inj_init = q.op == 0 ? q.noise : nullptr;
  // rtransition with injected noise: present the single step as row `time` of a [T][n][N] array
  c.inj_trans = (q.op == 1 && q.noise) ? q.noise - (size_t)(q.time - 1) * q.n_trans * q.N : nullptr;
  c.n_trans = q.n_trans; c.sk_M = q.sk_M; c.integrator = q.integrator; c.substeps = q.substeps;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q.N) return;
  double x[D];
  if (q.op != 0) {
#pragma unroll
    for (int k = 0; k < D; k++) x[k] = q.x[(size_t)i * D + k];
  }
  if (q.op == 0) Model::init(c, i, x);
  else if (q.op == 1) {
    double ua = 0.5, ub = 0.5;
    if (Model::USES_SHARED_U && !c.inj_trans) philox_u2(c.stream, i, q.time, CPIMH_P_RESAMPLE, 0, ua, ub);
    int st = Model::transition(c, q.time, i, x, Model::step_const(c, q.time), ub);
    if (st & 255) atomicOr(q.status, st & 255);
  }
  else { q.logw[i] = Model::logw(c, x, q.y); return; }
#pragma unroll
  for (int k = 0; k < D; k++) q.x[(size_t)i * D + k] = x[k];
}
template <class Model>
int launch_model_op(const ModelOpParams& q, cudaStream_t stream) {
  const int threads = 128, grid = (q.N + threads - 1) / threads;
  model_op_kernel<Model><<<grid, threads, (size_t)(CPIMH_MAX_THETA + q.npre + 1) * sizeof(double), stream>>>(q);
  CPIMH_CUDA_CHECK(cudaGetLastError());
  return CPIMH_OK;
}

}  // namespace cpimh
