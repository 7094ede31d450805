// api_ccpf.cu -- C ABI of the filters beyond the bootstrap hot path: pf() with ESS-adaptive resampling (R/CPF.R:106-155,
// kernel apf.cuh) and the conditional particle filters: CPF(ref_trajectory, with_as)
// (R/CPF.R:14-78), CPF_coupled (R/CPF_coupled.R:14-112), the CCPF kernels (R/pf_kernels.R:236-265) and
// coupled_ccpf_chains + H_bar (R/debiasing_algorithm.R:23-111, 284-316).  Kernels: ccpf.cuh.
#include <cstring>
#include <vector>
#include "host_common.h"

namespace cpimh {

struct CcpfWork {
  cpimh_config cfg;
  const ModelEntry* entry = nullptr;
  int Npad = 0, npre = 0, threads = 0, grid_max = 0;
  size_t smem = 0;
  DevBuf obs, theta, pre, hist_x, hist_a;
  int64_t launches = 0;
  ~CcpfWork() { obs.release(); theta.release(); pre.release(); hist_x.release(); hist_a.release(); }
};

static int ccpf_work_init(CcpfWork* w, const cpimh_config* cfg_in, const double* obs_host, int64_t max_filters) {
  w->cfg = *cfg_in;
  cpimh_config& c = w->cfg;
  int rc = api_validate(&c);
  if (rc) return rc;
  CPIMH_REQUIRE(c.shuffle == 0, "the conditional filters draw N-1 sorted ancestors for the N-1 free particles; cfg.shuffle must be 0");
  CPIMH_REQUIRE(c.nparticles >= 2, "a conditional filter needs at least 2 particles");
  w->entry = find_entry(c.model, c.dim);
  CPIMH_REQUIRE(w->entry != nullptr, "model %d with dimension %d has no conditional-filter kernel", c.model, c.dim);
  int dev = 0;
  CPIMH_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CPIMH_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  CPIMH_REQUIRE(prop.major >= 10, "libcpimh is built for sm_100a (Blackwell); device %d is sm_%d%d", dev, prop.major, prop.minor);
  std::vector<double> pre = api_precompute(c);
  w->npre = (int)pre.size();
  const int N = c.nparticles, T = c.nobs, D = c.dim;
  w->Npad = (N + 31) & ~31;
  w->smem = ccpf_smem_bytes(w->Npad, w->npre);
  CPIMH_REQUIRE(w->smem <= prop.sharedMemPerBlockOptin, "the conditional filter keeps 36 bytes per particle in shared memory: N=%d needs %zu bytes (max %zu)",
                N, w->smem, (size_t)prop.sharedMemPerBlockOptin);
  int threads = 32;
  while (threads < 512 && threads * 4 < N) threads <<= 1;     // about 4-8 particles per thread
  w->threads = threads;
  int bps = 1;
  rc = w->entry->occupancy_ccpf(threads, w->smem, &bps);
  if (rc) return rc;
  if (bps < 1) bps = 1;
  int64_t resident = (int64_t)bps * prop.multiProcessorCount;
  w->grid_max = (int)(max_filters < resident ? max_filters : resident);
  rc = w->obs.alloc(sizeof(double) * (size_t)T * c.dim_y);
  if (!rc) rc = w->theta.alloc(sizeof(double) * CPIMH_MAX_THETA);
  if (!rc) rc = w->pre.alloc(sizeof(double) * (pre.size() + 1));
  if (!rc) rc = w->hist_x.alloc(sizeof(double) * (size_t)w->grid_max * 2 * (T + 1) * D * w->Npad);
  if (!rc) rc = w->hist_a.alloc(sizeof(int) * (size_t)w->grid_max * 2 * T * w->Npad);
  if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(w->theta.p, c.theta, sizeof(double) * CPIMH_MAX_THETA, cudaMemcpyHostToDevice));
  if (!pre.empty()) CPIMH_CUDA_CHECK(cudaMemcpy(w->pre.p, pre.data(), sizeof(double) * pre.size(), cudaMemcpyHostToDevice));
  return api_upload_obs(c, obs_host, w->obs.p);
}

static void ccpf_fill(const CcpfWork& w, CcpfParams* p) {
  memset(p, 0, sizeof(*p));
  const cpimh_config& c = w.cfg;
  p->N = c.nparticles; p->T = c.nobs; p->dy = c.dim_y; p->Npad = w.Npad;
  p->obs = w.obs.as<double>(); p->theta = w.theta.as<double>(); p->pre = w.pre.as<double>();
  p->npre = w.npre; p->integrator = c.integrator; p->substeps = c.substeps;
  p->hist_x = w.hist_x.as<double>(); p->hist_a = w.hist_a.as<int>();
  p->n_init = api_n_init(c); p->n_trans = api_n_trans(c); p->sk_M = c.sk_max_events;
}
static int ccpf_launch(CcpfWork& w, const CcpfParams& p, cudaStream_t s) {
  CcpfLaunch L;
  L.threads = w.threads; L.smem = w.smem;
  L.grid = (int)(p.B < w.grid_max ? p.B : w.grid_max);
  w.launches++;
  return w.entry->launch_ccpf(p, L, s);
}

// ------------------------------------------------------------------ coupled_ccpf_chains state machine
struct CcRound {
  int PD, k, K, max_it, it;
  unsigned first_replicate;
  const int* active; int nA; int* next_active; int* next_count;
  unsigned long long* stream_ids;
  const double* new1; const double* new2; const double* new_logz; const int* new_status;
  double* st1; double* st2; double* lz1; double* lz2;
  int* single; int* iter; int* tau; int* status;
  double* accH; double* accD; double* hbar; double* bc;
  unsigned long long* nfilters;
  // Rao-Blackwellised twin (null when off): exp paths of the new filters / of the current states, accumulators, output
  const double* newe1; const double* newe2; double* e1; double* e2; double* accHe; double* accDe; double* hbar_rb;
};
__global__ void cc_prepare_kernel(CcRound q) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= q.nA) return;
  const int rep = q.active[a];
  q.stream_ids[a] = (unsigned long long)(q.first_replicate + (unsigned)rep) | ((unsigned long long)q.it << 32);
}
// pimh_init (R/pf_kernels.R:134-145): two independent bootstrap filters give chain_state1 / chain_state2
__global__ void cc_init_kernel(CcRound q, const int* pf_status1, const int* pf_status2) {
  const int rep = blockIdx.x;
  for (int j = threadIdx.x; j < q.PD; j += blockDim.x) {
    q.accH[(size_t)rep * q.PD + j] = (q.k <= 0) ? q.st1[(size_t)rep * q.PD + j] : 0.0;       // samples1[1,] (debiasing_algorithm.R:42)
    q.accD[(size_t)rep * q.PD + j] = 0.0;
    if (q.e1) { q.accHe[(size_t)rep * q.PD + j] = (q.k <= 0) ? q.e1[(size_t)rep * q.PD + j] : 0.0; q.accDe[(size_t)rep * q.PD + j] = 0.0; }
  }
  if (threadIdx.x == 0) {
    q.single[rep] = 1;                                   // iteration 1 is a single-kernel move of chain 1 (:50-52)
    q.iter[rep] = 0; q.tau[rep] = -1;
    int st = pf_status1[rep] | pf_status2[rep];
    if (st == 0 && (isnan(q.lz1[rep]) || isnan(q.lz2[rep]))) st = CPIMH_ERR_INVALID;
    q.status[rep] = st;
    if (st == 0) q.next_active[atomicAdd(q.next_count, 1)] = rep;
  }
}
__global__ void cc_accept_kernel(CcRound q) {
  const int rep = q.active[blockIdx.x];
  const int PD = q.PD, it = q.it;
  double* x1 = q.st1 + (size_t)rep * PD;
  double* x2 = q.st2 + (size_t)rep * PD;
  const double* n1 = q.new1 + (size_t)rep * PD;
  const double* n2 = q.new2 + (size_t)rep * PD;
  double* accH = q.accH + (size_t)rep * PD;
  double* accD = q.accD + (size_t)rep * PD;
  const int single = q.single[rep];
  int tau = q.tau[rep];
  int status = q.new_status[rep];
  if (status == 0 && isnan(q.new_logz[rep])) status = CPIMH_ERR_INVALID;
  const int t = it - 1;
  const bool mcmc = (it >= q.k && it <= q.K);
  const double coef = (double)min(t - q.k + 1, q.K - q.k + 1);
  int same = 1;
  for (int j = threadIdx.x; j < PD; j += blockDim.x) {
    const double xa = n1[j];
    // iteration 1 moves chain 1 only; after the meeting both chains take the single kernel's output (:68-73)
    const double xb = (it == 1) ? x2[j] : (single ? xa : n2[j]);
    x1[j] = xa; x2[j] = xb;
    if (xa != xb) same = 0;
  }
  same = __syncthreads_and(same);
  int finished = 0;
  if (status) finished = 1;
  if (it >= 2 && !single && same && tau < 0) tau = it;                         // :82-86 all(chain_state1 == chain_state2)
  if (tau > 0 && it >= max(tau, q.K)) finished = 1;                            // :104-106
  if (q.max_it >= 0 && it >= q.max_it) finished = 1;
  // H_bar (debiasing_algorithm.R:296-312): MCMC sum over k..K plus coef_t (X_{t+1} - Y_t) for k <= t <= tau - 1, skipped
  // altogether when tau <= k + 1 (:303).  For the states the skipped term is zero; for the weighted-average paths it is not.
  const bool bias = (t >= q.k) && !(tau > 0 && tau <= q.k + 1);
  for (int j = threadIdx.x; j < PD; j += blockDim.x) {
    const double xa = x1[j], xb = x2[j];
    if (mcmc) accH[j] += xa;
    if (bias) accD[j] = accD[j] + coef * (xa - xb);
    if (q.e1) {                                                                // the same recursion over the filters' weighted-average paths
      const size_t o = (size_t)rep * PD + j;
      const double ea = q.newe1[o];
      const double eb = (it == 1) ? q.e2[o] : (single ? ea : q.newe2[o]);
      q.e1[o] = ea; q.e2[o] = eb;
      if (mcmc) q.accHe[o] += ea;
      if (bias) q.accDe[o] = q.accDe[o] + coef * (ea - eb);
    }
  }
  if (finished) {
    const double denom = (double)(q.K - q.k + 1);
    for (int j = threadIdx.x; j < PD; j += blockDim.x) {
      q.hbar[(size_t)rep * PD + j] = (accH[j] + accD[j]) / denom;
      q.bc[(size_t)rep * PD + j] = accD[j] / denom;
      if (q.e1) q.hbar_rb[(size_t)rep * PD + j] = (q.accHe[(size_t)rep * PD + j] + q.accDe[(size_t)rep * PD + j]) / denom;
    }
  }
  if (threadIdx.x == 0) {
    if (it == 1 || single) { q.lz1[rep] = q.new_logz[rep]; if (it > 1) q.lz2[rep] = q.new_logz[rep]; }   // coupled kernel leaves log_pdf unchanged
    q.single[rep] = (tau > 0) ? 1 : 0;
    q.iter[rep] = it; q.tau[rep] = tau;
    if (status) q.status[rep] = status;
    if (finished) atomicAdd(q.nfilters, (unsigned long long)(2 + it));
    else q.next_active[atomicAdd(q.next_count, 1)] = rep;
  }
}
__global__ void cc_reduce_kernel(const double* hbar, long long R, int P, double* sums) {
  const int j = blockIdx.x;
  double s = 0.0, s2 = 0.0;
  for (long long r = threadIdx.x; r < R; r += blockDim.x) { const double v = hbar[(size_t)r * P + j]; s += v; s2 += v * v; }
  __shared__ double scr[64];
  s = block_sum(s, scr);
  s2 = block_sum(s2, scr);
  if (threadIdx.x == 0) { sums[j] = s; sums[P + j] = s2; }
}

}  // namespace cpimh

using namespace cpimh;

extern "C" {

int cpimh_ccpf_batch(const cpimh_config* cfg, const double* obs_host, int64_t B, uint64_t seed, uint64_t first_filter,
                     int nsys, int with_as, const double* ref1_host, const double* ref2_host, const cpimh_noise* noise,
                     double* traj1, double* traj2, double* logz, int32_t* path_index, int32_t* ancestors,
                     double* exp_path1, double* exp_path2) {
  CPIMH_REQUIRE(cfg && obs_host && ref1_host && traj1, "null argument");
  CPIMH_REQUIRE(B >= 1, "nfilters must be >= 1");
  CPIMH_REQUIRE(nsys == 1 || nsys == 2, "nsys must be 1 (CPF with a reference trajectory) or 2 (CPF_coupled)");
  CPIMH_REQUIRE(nsys == 1 || (ref2_host && traj2), "CPF_coupled needs two reference trajectories and two outputs");
  CPIMH_REQUIRE(!noise || !(noise->resample || noise->shuffle || noise->path_u), "conditional filters accept injected init / transition noise only");
  CcpfWork w;
  int rc = ccpf_work_init(&w, cfg, obs_host, B);
  if (rc) return rc;
  const cpimh_config& c = w.cfg;
  const size_t N = c.nparticles, T = c.nobs, PD = (size_t)(c.nobs + 1) * c.dim;
  DevBuf ref1, ref2, t1, t2, lz, pidx, status, anc, ninit, ntrans, ex1, ex2;
  struct Rel { DevBuf* b[12]; ~Rel() { for (auto x : b) x->release(); } } rel{{&ref1, &ref2, &t1, &t2, &lz, &pidx, &status, &anc, &ninit, &ntrans, &ex1, &ex2}};
  const bool want_exp = exp_path1 != nullptr || exp_path2 != nullptr;
  if (want_exp) { rc = ex1.alloc(sizeof(double) * B * PD); if (!rc) rc = ex2.alloc(sizeof(double) * B * PD); if (rc) return rc; }
  rc = ref1.alloc(sizeof(double) * B * PD);
  if (!rc) rc = t1.alloc(sizeof(double) * B * PD);
  if (!rc && nsys == 2) rc = ref2.alloc(sizeof(double) * B * PD);
  if (!rc) rc = t2.alloc(sizeof(double) * B * PD);
  if (!rc) rc = lz.alloc(sizeof(double) * B);
  if (!rc) rc = pidx.alloc(sizeof(int) * 2 * B);
  if (!rc) rc = status.alloc(sizeof(int) * B);
  if (!rc && ancestors) rc = anc.alloc(sizeof(int) * (size_t)B * 2 * T * N);
  if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(ref1.p, ref1_host, sizeof(double) * B * PD, cudaMemcpyHostToDevice));
  if (nsys == 2) CPIMH_CUDA_CHECK(cudaMemcpy(ref2.p, ref2_host, sizeof(double) * B * PD, cudaMemcpyHostToDevice));
  if (ancestors) CPIMH_CUDA_CHECK(cudaMemset(anc.p, 0xff, anc.bytes));
  CcpfParams p;
  ccpf_fill(w, &p);
  if (noise && noise->init) {
    size_t n = (size_t)B * p.n_init * N;
    rc = ninit.alloc(sizeof(double) * n);
    if (rc) return rc;
    CPIMH_CUDA_CHECK(cudaMemcpy(ninit.p, noise->init, sizeof(double) * n, cudaMemcpyHostToDevice));
    p.inj_init = ninit.as<double>();
  }
  if (noise && noise->trans) {
    size_t n = (size_t)B * T * p.n_trans * N;
    rc = ntrans.alloc(sizeof(double) * n);
    if (rc) return rc;
    CPIMH_CUDA_CHECK(cudaMemcpy(ntrans.p, noise->trans, sizeof(double) * n, cudaMemcpyHostToDevice));
    p.inj_trans = ntrans.as<double>();
  }
  p.B = B; p.nsys = nsys; p.with_as = with_as ? 1 : 0; p.seed = seed; p.first_filter = first_filter;
  p.ref1 = ref1.as<double>(); p.ref2 = nsys == 2 ? ref2.as<double>() : nullptr;
  p.traj1 = t1.as<double>(); p.traj2 = t2.as<double>(); p.logz = lz.as<double>(); p.path_index = pidx.as<int>();
  p.status = status.as<int>(); p.anc_rec = ancestors ? anc.as<int>() : nullptr;
  if (want_exp) { p.exp1 = ex1.as<double>(); p.exp2 = ex2.as<double>(); }
  rc = ccpf_launch(w, p, nullptr);
  if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaDeviceSynchronize());
  std::vector<int> st((size_t)B);
  CPIMH_CUDA_CHECK(cudaMemcpy(st.data(), status.p, sizeof(int) * B, cudaMemcpyDeviceToHost));
  for (int64_t b = 0; b < B; b++)
    if (st[(size_t)b]) { set_error("conditional filter %lld failed with status %d", (long long)b, st[(size_t)b]); return st[(size_t)b]; }
  CPIMH_CUDA_CHECK(cudaMemcpy(traj1, t1.p, sizeof(double) * B * PD, cudaMemcpyDeviceToHost));
  if (nsys == 2) CPIMH_CUDA_CHECK(cudaMemcpy(traj2, t2.p, sizeof(double) * B * PD, cudaMemcpyDeviceToHost));
  if (logz) CPIMH_CUDA_CHECK(cudaMemcpy(logz, lz.p, sizeof(double) * B, cudaMemcpyDeviceToHost));
  if (path_index) CPIMH_CUDA_CHECK(cudaMemcpy(path_index, pidx.p, sizeof(int) * 2 * B, cudaMemcpyDeviceToHost));
  if (ancestors) CPIMH_CUDA_CHECK(cudaMemcpy(ancestors, anc.p, anc.bytes, cudaMemcpyDeviceToHost));
  if (exp_path1) CPIMH_CUDA_CHECK(cudaMemcpy(exp_path1, ex1.p, sizeof(double) * B * PD, cudaMemcpyDeviceToHost));
  if (exp_path2 && nsys == 2) CPIMH_CUDA_CHECK(cudaMemcpy(exp_path2, ex2.p, sizeof(double) * B * PD, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

int cpimh_coupled_ccpf(const cpimh_config* cfg, const double* obs_host, int64_t R, uint64_t seed, uint32_t first_replicate,
                       int k, int K, int max_iterations, int with_as, cpimh_chain_out* o) {
  CPIMH_REQUIRE(cfg && obs_host && o, "null argument");
  CPIMH_REQUIRE(R >= 1, "nreplicates must be >= 1");
  CPIMH_REQUIRE(k >= 0 && K >= k, "need 0 <= k <= K (got k=%d K=%d)", k, K);
  CPIMH_REQUIRE(max_iterations < 0 || max_iterations >= (K > 1 ? K : 1), "max_iterations must be >= max(K, 1) or negative (unbounded)");
  CPIMH_REQUIRE(max_iterations < 4096, "iteration index must fit 12 bits of the Philox counter (max_iterations < 4096)");
  const bool rb = o->hbar_rb || o->sum_hbar_rb || o->sum_hbar_rb_sq;
  const int max_it = max_iterations < 0 ? 4094 : max_iterations;
  CcpfWork w;
  int rc = ccpf_work_init(&w, cfg, obs_host, R);
  if (rc) return rc;
  const cpimh_config& c = w.cfg;
  const int D = c.dim, P = c.nobs + 1, PD = P * D;
  cudaStream_t s = nullptr;
  // ---- initial states: two independent bootstrap filters per replicate (iteration indices 0 and 0xFFF)
  cpimh_config cpf = c;
  cpf.store_paths = cpf.store_paths == 0 ? 1 : cpf.store_paths;
  cpimh_pf* pf = nullptr;
  rc = cpimh_pf_create(&cpf, R, &pf);
  if (rc) return rc;
  DevBuf st1, st2, n1, n2, lz1, lz2, nlz, nst, single, iter, tau, status, accH, accD, hbar, bc, act0, act1, cnt, ids, nf, sums, pst1, pst2;
  DevBuf e1, e2, ne1, ne2, accHe, accDe, hbar_rb, sums_rb;
  DevBuf* all[] = {&st1, &st2, &n1, &n2, &lz1, &lz2, &nlz, &nst, &single, &iter, &tau, &status, &accH, &accD, &hbar, &bc, &act0, &act1, &cnt, &ids, &nf, &sums, &pst1, &pst2,
                   &e1, &e2, &ne1, &ne2, &accHe, &accDe, &hbar_rb, &sums_rb};
  struct Rel { DevBuf** b; size_t n; cpimh_pf** pf; ~Rel() { for (size_t i = 0; i < n; i++) b[i]->release(); if (*pf) cpimh_pf_destroy(*pf); } } rel{all, sizeof(all) / sizeof(all[0]), &pf};
  const size_t RPD = sizeof(double) * (size_t)R * PD;
  DevBuf* dbl[] = {&st1, &st2, &n1, &n2, &accH, &accD, &hbar, &bc};
  for (auto b : dbl) { rc = b->alloc(RPD); if (rc) return rc; }
  if (rb) {
    DevBuf* dre[] = {&e1, &e2, &ne1, &ne2, &accHe, &accDe, &hbar_rb};
    for (auto b : dre) { rc = b->alloc(RPD); if (rc) return rc; }
    rc = sums_rb.alloc(sizeof(double) * (2 * (size_t)PD)); if (rc) return rc;
    rc = cpimh_pf_all_particles(pf, 1); if (rc) return rc;
  }
  DevBuf* dR[] = {&lz1, &lz2, &nlz};
  for (auto b : dR) { rc = b->alloc(sizeof(double) * (size_t)R); if (rc) return rc; }
  DevBuf* iR[] = {&nst, &single, &iter, &tau, &status, &act0, &act1, &pst1, &pst2};
  for (auto b : iR) { rc = b->alloc(sizeof(int) * (size_t)R); if (rc) return rc; }
  rc = cnt.alloc(sizeof(int)); if (rc) return rc;
  rc = ids.alloc(sizeof(unsigned long long) * (size_t)R); if (rc) return rc;
  rc = nf.alloc(sizeof(unsigned long long)); if (rc) return rc;
  rc = sums.alloc(sizeof(double) * (2 * (size_t)PD)); if (rc) return rc;
  rc = cpimh_pf_set_observations(pf, obs_host);
  if (rc) return rc;
  double *d_lz = nullptr, *d_tr = nullptr;
  int* d_st = nullptr;
  rc = cpimh_pf_device_outputs(pf, &d_lz, &d_tr);
  if (rc) return rc;
  rc = cpimh_pf_device_status(pf, &d_st);
  if (rc) return rc;
  double* d_ex = nullptr;
  if (rb) { rc = cpimh_pf_device_exp_paths(pf, &d_ex); if (rc) return rc; }
  for (int which = 0; which < 2; which++) {
    const uint64_t first = (uint64_t)first_replicate | (which ? ((uint64_t)0xFFF << 32) : 0ull);
    rc = cpimh_pf_run(pf, R, seed, first, s);
    if (rc) return rc;
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(which ? st2.p : st1.p, d_tr, RPD, cudaMemcpyDeviceToDevice, s));
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(which ? lz2.p : lz1.p, d_lz, sizeof(double) * (size_t)R, cudaMemcpyDeviceToDevice, s));
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(which ? pst2.p : pst1.p, d_st, sizeof(int) * (size_t)R, cudaMemcpyDeviceToDevice, s));
    if (rb) CPIMH_CUDA_CHECK(cudaMemcpyAsync(which ? e2.p : e1.p, d_ex, RPD, cudaMemcpyDeviceToDevice, s));
  }
  CcRound q;
  memset(&q, 0, sizeof(q));
  q.PD = PD; q.k = k; q.K = K; q.max_it = max_it; q.first_replicate = first_replicate;
  q.stream_ids = ids.as<unsigned long long>();
  q.new1 = n1.as<double>(); q.new2 = n2.as<double>(); q.new_logz = nlz.as<double>(); q.new_status = nst.as<int>();
  q.st1 = st1.as<double>(); q.st2 = st2.as<double>(); q.lz1 = lz1.as<double>(); q.lz2 = lz2.as<double>();
  q.single = single.as<int>(); q.iter = iter.as<int>(); q.tau = tau.as<int>(); q.status = status.as<int>();
  q.accH = accH.as<double>(); q.accD = accD.as<double>(); q.hbar = hbar.as<double>(); q.bc = bc.as<double>();
  q.nfilters = nf.as<unsigned long long>(); q.next_count = cnt.as<int>();
  if (rb) {
    q.newe1 = ne1.as<double>(); q.newe2 = ne2.as<double>(); q.e1 = e1.as<double>(); q.e2 = e2.as<double>();
    q.accHe = accHe.as<double>(); q.accDe = accDe.as<double>(); q.hbar_rb = hbar_rb.as<double>();
  }
  CPIMH_CUDA_CHECK(cudaMemsetAsync(nf.p, 0, sizeof(unsigned long long), s));
  CPIMH_CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, sizeof(int), s));
  q.next_active = act0.as<int>();
  cc_init_kernel<<<(unsigned)R, 128, 0, s>>>(q, pst1.as<int>(), pst2.as<int>());
  int nA = 0, flip = 0;
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(&nA, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  CcpfParams p;
  ccpf_fill(w, &p);
  p.with_as = with_as ? 1 : 0; p.seed = seed; p.stream_ids = ids.as<unsigned long long>();
  p.single = single.as<int>();
  p.ref1 = st1.as<double>(); p.ref2 = st2.as<double>();
  p.traj1 = n1.as<double>(); p.traj2 = n2.as<double>(); p.logz = nlz.as<double>(); p.status = nst.as<int>();
  if (rb) { p.exp1 = ne1.as<double>(); p.exp2 = ne2.as<double>(); }
  int64_t filters = 2 * R;
  for (int it = 1; nA > 0; it++) {
    q.it = it;
    q.active = flip ? act1.as<int>() : act0.as<int>();
    q.next_active = flip ? act0.as<int>() : act1.as<int>();
    q.nA = nA;
    CPIMH_CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, sizeof(int), s));
    cc_prepare_kernel<<<(unsigned)((nA + 255) / 256), 256, 0, s>>>(q);
    p.B = nA; p.slot = q.active;
    rc = ccpf_launch(w, p, s);
    if (rc) return rc;
    cc_accept_kernel<<<(unsigned)nA, 128, 0, s>>>(q);
    filters += nA;
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(&nA, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
    flip ^= 1;
  }
  cc_reduce_kernel<<<PD, 256, 0, s>>>(hbar.as<double>(), R, PD, sums.as<double>());
  if (rb) cc_reduce_kernel<<<PD, 256, 0, s>>>(hbar_rb.as<double>(), R, PD, sums_rb.as<double>());
  CPIMH_CUDA_CHECK(cudaGetLastError());
  std::vector<int> stv((size_t)R);
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(stv.data(), status.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  for (int64_t r = 0; r < R; r++)
    if (stv[(size_t)r]) { set_error("replicate %lld failed with status %d", (long long)r, stv[(size_t)r]); return stv[(size_t)r]; }
  std::vector<double> hs(2 * (size_t)PD);
  CPIMH_CUDA_CHECK(cudaMemcpy(hs.data(), sums.p, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost));
  if (o->hbar) CPIMH_CUDA_CHECK(cudaMemcpy(o->hbar, hbar.p, RPD, cudaMemcpyDeviceToHost));
  if (o->bc) CPIMH_CUDA_CHECK(cudaMemcpy(o->bc, bc.p, RPD, cudaMemcpyDeviceToHost));
  if (o->meetingtime) CPIMH_CUDA_CHECK(cudaMemcpy(o->meetingtime, tau.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost));
  if (o->iteration) CPIMH_CUDA_CHECK(cudaMemcpy(o->iteration, iter.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost));
  if (o->sum_hbar) memcpy(o->sum_hbar, hs.data(), sizeof(double) * PD);
  if (o->sum_hbar_sq) memcpy(o->sum_hbar_sq, hs.data() + PD, sizeof(double) * PD);
  if (o->nfilters) *o->nfilters = filters;
  if (rb) {
    CPIMH_CUDA_CHECK(cudaMemcpy(hs.data(), sums_rb.p, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost));
    if (o->hbar_rb) CPIMH_CUDA_CHECK(cudaMemcpy(o->hbar_rb, hbar_rb.p, RPD, cudaMemcpyDeviceToHost));
    if (o->sum_hbar_rb) memcpy(o->sum_hbar_rb, hs.data(), sizeof(double) * PD);
    if (o->sum_hbar_rb_sq) memcpy(o->sum_hbar_rb_sq, hs.data() + PD, sizeof(double) * PD);
  }
  return CPIMH_OK;
}

int cpimh_pf_adaptive_batch(const cpimh_config* cfg_in, const double* obs_host, int64_t B, uint64_t seed, uint64_t first_filter,
                            double ess_threshold, int systematic, double* loglik, double* loglik_t, double* pf_means, double* path,
                            int32_t* resampled, int32_t* ancestors) {
  CPIMH_REQUIRE(cfg_in && obs_host && loglik && path, "null argument");
  CPIMH_REQUIRE(B >= 1, "nfilters must be >= 1");
  cpimh_config c = *cfg_in;
  int rc = api_validate(&c);
  if (rc) return rc;
  const ModelEntry* e = find_entry(c.model, c.dim);
  CPIMH_REQUIRE(e != nullptr, "model %d with dimension %d has no adaptive-filter kernel", c.model, c.dim);
  int dev = 0;
  CPIMH_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CPIMH_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  CPIMH_REQUIRE(prop.major >= 10, "libcpimh is built for sm_100a (Blackwell); device %d is sm_%d%d", dev, prop.major, prop.minor);
  std::vector<double> pre = api_precompute(c);
  const int N = c.nparticles, T = c.nobs, D = c.dim, Npad = (N + 31) & ~31;
  const size_t smem = apf_smem_bytes(Npad, (int)pre.size());
  CPIMH_REQUIRE(smem <= prop.sharedMemPerBlockOptin, "pf() keeps 24 bytes per particle in shared memory: N=%d needs %zu bytes (max %zu)", N, smem,
                (size_t)prop.sharedMemPerBlockOptin);
  int threads = 32;
  while (threads < 512 && threads * 8 < N) threads <<= 1;
  int bps = 1;
  rc = e->occupancy_apf(threads, smem, &bps);
  if (rc) return rc;
  if (bps < 1) bps = 1;
  const int64_t resident = (int64_t)bps * prop.multiProcessorCount;
  const int grid = (int)(B < resident ? B : resident);
  const size_t PD = (size_t)(T + 1) * D;
  DevBuf obs, theta, dpre, hx, ha, pr, ll, llt, pm, pa, rs, an, stt;
  struct Rel { DevBuf* b[13]; ~Rel() { for (auto x : b) x->release(); } } rel{{&obs, &theta, &dpre, &hx, &ha, &pr, &ll, &llt, &pm, &pa, &rs, &an, &stt}};
  rc = obs.alloc(sizeof(double) * (size_t)T * c.dim_y);
  if (!rc) rc = theta.alloc(sizeof(double) * CPIMH_MAX_THETA);
  if (!rc) rc = dpre.alloc(sizeof(double) * (pre.size() + 1));
  if (!rc) rc = hx.alloc(sizeof(double) * (size_t)grid * (T + 1) * D * Npad);
  if (!rc) rc = ha.alloc(sizeof(int) * (size_t)grid * T * Npad);
  if (!rc) rc = pr.alloc(sizeof(double) * (size_t)grid * D * Npad);
  if (!rc) rc = ll.alloc(sizeof(double) * B);
  if (!rc) rc = llt.alloc(sizeof(double) * B * T);
  if (!rc) rc = pm.alloc(sizeof(double) * B * T * D);
  if (!rc) rc = pa.alloc(sizeof(double) * B * PD);
  if (!rc) rc = rs.alloc(sizeof(int) * B * T);
  if (!rc) rc = stt.alloc(sizeof(int) * B);
  if (!rc && ancestors) rc = an.alloc(sizeof(int) * (size_t)B * T * N);
  if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpy(theta.p, c.theta, sizeof(double) * CPIMH_MAX_THETA, cudaMemcpyHostToDevice));
  if (!pre.empty()) CPIMH_CUDA_CHECK(cudaMemcpy(dpre.p, pre.data(), sizeof(double) * pre.size(), cudaMemcpyHostToDevice));
  rc = api_upload_obs(c, obs_host, obs.p);
  if (rc) return rc;
  ApfParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.T = T; p.dy = c.dim_y; p.Npad = Npad; p.B = B; p.seed = seed; p.first_filter = first_filter;
  p.ess_threshold = ess_threshold; p.systematic = systematic ? 1 : 0;
  p.obs = obs.as<double>(); p.theta = theta.as<double>(); p.pre = dpre.as<double>(); p.npre = (int)pre.size();
  p.integrator = c.integrator; p.substeps = c.substeps; p.sk_M = c.sk_max_events;
  p.hist_x = hx.as<double>(); p.hist_a = ha.as<int>(); p.prop = pr.as<double>();
  p.loglik = ll.as<double>(); p.loglik_t = llt.as<double>(); p.pf_means = pm.as<double>(); p.path = pa.as<double>();
  p.resampled = rs.as<int>(); p.anc_rec = ancestors ? an.as<int>() : nullptr; p.status = stt.as<int>();
  ApfLaunch L;
  L.grid = grid; L.threads = threads; L.smem = smem;
  rc = e->launch_apf(p, L, nullptr);
  if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaDeviceSynchronize());
  std::vector<int> st((size_t)B);
  CPIMH_CUDA_CHECK(cudaMemcpy(st.data(), stt.p, sizeof(int) * B, cudaMemcpyDeviceToHost));
  for (int64_t b = 0; b < B; b++)
    if (st[(size_t)b]) { set_error("adaptive filter %lld failed with status %d", (long long)b, st[(size_t)b]); return st[(size_t)b]; }
  CPIMH_CUDA_CHECK(cudaMemcpy(loglik, ll.p, sizeof(double) * B, cudaMemcpyDeviceToHost));
  CPIMH_CUDA_CHECK(cudaMemcpy(path, pa.p, sizeof(double) * B * PD, cudaMemcpyDeviceToHost));
  if (loglik_t) CPIMH_CUDA_CHECK(cudaMemcpy(loglik_t, llt.p, sizeof(double) * B * T, cudaMemcpyDeviceToHost));
  if (pf_means) CPIMH_CUDA_CHECK(cudaMemcpy(pf_means, pm.p, sizeof(double) * B * T * D, cudaMemcpyDeviceToHost));
  if (resampled) CPIMH_CUDA_CHECK(cudaMemcpy(resampled, rs.p, sizeof(int) * B * T, cudaMemcpyDeviceToHost));
  if (ancestors) CPIMH_CUDA_CHECK(cudaMemcpy(ancestors, an.p, an.bytes, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

}  // extern "C"
