// api_ccpf.cu -- C ABI of the filters beyond the bootstrap hot path: pf() with ESS-adaptive resampling (R/CPF.R:106-155,
// kernel apf.cuh) and the conditional particle filters: CPF(ref_trajectory, with_as)
// (R/CPF.R:14-78), CPF_coupled (R/CPF_coupled.R:14-112), the CCPF kernels (R/pf_kernels.R:236-265) and
// coupled_ccpf_chains + H_bar (R/debiasing_algorithm.R:23-111, 284-316).  Kernels: ccpf.cuh.
#include <stdlib.h>
#include <cstring>
#include <chrono>
#include <cstdlib>
#include <mutex>
#include <vector>
#include "host_common.h"

namespace cpimh {

// steps the last conditional / adaptive filter call repeated in reference order (diagnostic; cpimh_last_exact_repeats)
static int64_t g_last_exact_repeats = 0;
static double margin_scale() { const char* e = getenv("CPIMH_EXACT_MARGIN_SCALE"); const double v = e ? atof(e) : 1.0; return v > 0.0 ? v : 1.0; }   // test aid: widen the knot margin so that repeats happen on random draws

struct CcpfWork {
  cpimh_config cfg;
  const ModelEntry* entry = nullptr;
  int Npad = 0, npre = 0, threads = 0, grid_max = 0;
  size_t smem = 0;
  DevBuf obs, theta, pre, hist_x, hist_a, near, wsave;   // near: one word counting the steps repeated in reference order; wsave: the step's normalised weights
  int64_t launches = 0;
  ~CcpfWork() { obs.release(); theta.release(); pre.release(); hist_x.release(); hist_a.release(); near.release(); wsave.release(); }
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
  CPIMH_CUDA_CHECK(cached_device_properties(&prop, dev));
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
  if (!rc) rc = w->near.alloc(sizeof(int));
  if (!rc) rc = w->wsave.alloc(sizeof(double) * (size_t)w->grid_max * 2 * w->Npad);
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
  p->near = w.near.as<int>(); p->wsave = w.wsave.as<double>(); p->margin_rel = 12.0 * ((double)c.nparticles + 2.0) * 0x1p-53 * margin_scale();
  const char* force = getenv("CPIMH_CCPF_EXACT");                 // test aid: reference-order mode from the start
  p->exact = (force && atoi(force)) ? 1 : 0;
}
static int ccpf_launch(CcpfWork& w, const CcpfParams& p, cudaStream_t s) {
  CcpfLaunch L;
  L.threads = w.threads; L.smem = w.smem;
  L.grid = (int)(p.B < w.grid_max ? p.B : w.grid_max);
  w.launches++;
  return w.entry->launch_ccpf(p, L, s);
}

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
    if (st == 0 && q.next_active) q.next_active[atomicAdd(q.next_count, 1)] = rep;
  }
}
__global__ void cc_accept_kernel(CcRound q) { cc_accept(q, q.active[blockIdx.x], q.it); }
// deterministic sums over the finished replicates (status 0): sums[0..P) = sum hbar, [P..2P) = sum hbar^2, [2P] = filters run,
// [2P+1] = replicates that entered the sums (a replicate capped by max_iterations is left out: CPIMH_ERR_NOT_MET)
__global__ void cc_reduce_kernel(const double* hbar, const int* status, long long R, int P, const unsigned long long* nfilters, double* sums) {
  __shared__ double scr[64];
  const int j = blockIdx.x;
  if (j == P) {
    double n = 0;
    for (long long r = threadIdx.x; r < R; r += blockDim.x) n += (status[r] == 0) ? 1.0 : 0.0;
    n = block_sum(n, scr);
    if (threadIdx.x == 0) { sums[2 * P] = (double)(*nfilters); sums[2 * P + 1] = n; }
    return;
  }
  double s = 0.0, s2 = 0.0;
  for (long long r = threadIdx.x; r < R; r += blockDim.x) { if (status[r] != 0) continue; const double v = hbar[(size_t)r * P + j]; s += v; s2 += v * v; }
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
  // exact ancestors: a step with a uniform within rounding distance of a CDF knot (or of alpha) is repeated in reference order
  // inside the kernel (ccpf.cuh); near counts those steps
  CPIMH_CUDA_CHECK(cudaMemset(w.near.p, 0, sizeof(int)));
  rc = ccpf_launch(w, p, nullptr);
  if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaDeviceSynchronize());
  { int nrep = 0; CPIMH_CUDA_CHECK(cudaMemcpy(&nrep, w.near.p, sizeof(int), cudaMemcpyDeviceToHost)); g_last_exact_repeats = nrep; }
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

// Workspaces of the one-shot chain entry, kept between calls (grow-only buffers, the bootstrap-filter handle of the initial
// states): the R shim calls this entry once per estimator batch.  cpimh_release_cached() frees them.
struct CcChainCache {
  CcpfWork w;
  cpimh_pf* pf = nullptr;
  cpimh_config pf_cfg;
  int64_t pf_R = 0;
  DevBuf st1, st2, n1, n2, lz1, lz2, nlz, nst, single, iter, tau, status, accH, accD, hbar, bc, cnt, nf, sums, pst1, pst2, trunc;
  DevBuf e1, e2, ne1, ne2, accHe, accDe, hbar_rb, sums_rb;
  void release() {
    DevBuf* all[] = {&st1, &st2, &n1, &n2, &lz1, &lz2, &nlz, &nst, &single, &iter, &tau, &status, &accH, &accD, &hbar, &bc, &cnt, &nf, &sums, &pst1, &pst2, &trunc,
                     &e1, &e2, &ne1, &ne2, &accHe, &accDe, &hbar_rb, &sums_rb, &w.obs, &w.theta, &w.pre, &w.hist_x, &w.hist_a, &w.near, &w.wsave};
    for (auto x : all) x->release();
    if (pf) { cpimh_pf_destroy(pf); pf = nullptr; pf_R = 0; }
  }
};
static CcChainCache g_cc;
static std::recursive_mutex g_cc_mu;

int cpimh_ccpf_release_cached(void) {
  std::lock_guard<std::recursive_mutex> lk(g_cc_mu);
  g_cc.release();
  return CPIMH_OK;
}

static int coupled_ccpf_impl(const cpimh_config* cfg, const double* obs_host, int64_t R, uint64_t seed, uint32_t first_replicate,
                             int k, int K, int max_iterations, int with_as, double rg_lambda, const int32_t* truncation, cpimh_chain_out* o);
int cpimh_coupled_ccpf(const cpimh_config* cfg, const double* obs_host, int64_t R, uint64_t seed, uint32_t first_replicate,
                       int k, int K, int max_iterations, int with_as, cpimh_chain_out* o) {
  return coupled_ccpf_impl(cfg, obs_host, R, seed, first_replicate, k, K, max_iterations, with_as, 0.0, nullptr, o);
}
int cpimh_rheeglynn(const cpimh_config* cfg, const double* obs_host, int64_t R, uint64_t seed, uint32_t first_replicate,
                    double lambda, const int32_t* truncation, int max_iterations, int with_as, cpimh_chain_out* o) {
  CPIMH_REQUIRE(lambda == 0.0 || (lambda > 0.0 && lambda < 1.0 && truncation), "lambda must be 0, or in (0, 1) with one truncation variable per replicate");
  return coupled_ccpf_impl(cfg, obs_host, R, seed, first_replicate, 0, 0, max_iterations, with_as, lambda, lambda > 0.0 ? truncation : nullptr, o);
}
static int coupled_ccpf_impl(const cpimh_config* cfg, const double* obs_host, int64_t R, uint64_t seed, uint32_t first_replicate,
                             int k, int K, int max_iterations, int with_as, double rg_lambda, const int32_t* truncation, cpimh_chain_out* o) {
  CPIMH_REQUIRE(cfg && obs_host && o, "null argument");
  CPIMH_REQUIRE(R >= 1, "nreplicates must be >= 1");
  CPIMH_REQUIRE(k >= 0 && K >= k, "need 0 <= k <= K (got k=%d K=%d)", k, K);
  CPIMH_REQUIRE(max_iterations < 0 || max_iterations >= (K > 1 ? K : 1), "max_iterations must be >= max(K, 1) or negative (unbounded)");
  CPIMH_REQUIRE(max_iterations < CPIMH_MAX_ITERATION, "iteration index must fit 20 bits of the Philox counter; the last value is the second initial filter (max_iterations < %d)", CPIMH_MAX_ITERATION);
  const bool rb = o->hbar_rb || o->sum_hbar_rb || o->sum_hbar_rb_sq;
  const int max_it = max_iterations < 0 ? CPIMH_MAX_ITERATION - 1 : max_iterations;
  std::lock_guard<std::recursive_mutex> lk(g_cc_mu);
  const bool timing = getenv("CPIMH_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_start = now();
  CcChainCache& g = g_cc;
  CcpfWork& w = g.w;
  int rc = ccpf_work_init(&w, cfg, obs_host, R);
  if (rc) return rc;
  const cpimh_config& c = w.cfg;
  const int D = c.dim, P = c.nobs + 1, PD = P * D;
  cudaStream_t s = nullptr;
  // ---- initial states: two independent bootstrap filters per replicate (iteration indices 0 and CPIMH_MAX_ITERATION = 0xFFFFF, which no chain iteration can reach)
  cpimh_config cpf = c;
  cpf.store_paths = cpf.store_paths == 0 ? 1 : cpf.store_paths;
  if (!(g.pf && g.pf_R >= R && memcmp(&g.pf_cfg, &cpf, sizeof(cpf)) == 0)) {
    if (g.pf) { cpimh_pf_destroy(g.pf); g.pf = nullptr; g.pf_R = 0; }
    rc = cpimh_pf_create(&cpf, R, &g.pf);
    if (rc) { g.pf = nullptr; return rc; }
    g.pf_cfg = cpf; g.pf_R = R;
  }
  cpimh_pf* pf = g.pf;
  const size_t RPD = sizeof(double) * (size_t)R * PD;
  DevBuf* dbl[] = {&g.st1, &g.st2, &g.n1, &g.n2, &g.accH, &g.accD, &g.hbar, &g.bc};
  for (auto b : dbl) { rc = b->alloc(RPD); if (rc) return rc; }
  rc = cpimh_pf_all_particles(pf, rb ? 1 : 0); if (rc) return rc;
  if (rb) {
    DevBuf* dre[] = {&g.e1, &g.e2, &g.ne1, &g.ne2, &g.accHe, &g.accDe, &g.hbar_rb};
    for (auto b : dre) { rc = b->alloc(RPD); if (rc) return rc; }
    rc = g.sums_rb.alloc(sizeof(double) * (2 * (size_t)PD + 2)); if (rc) return rc;
  }
  DevBuf* dR[] = {&g.lz1, &g.lz2, &g.nlz};
  for (auto b : dR) { rc = b->alloc(sizeof(double) * (size_t)R); if (rc) return rc; }
  DevBuf* iR[] = {&g.nst, &g.single, &g.iter, &g.tau, &g.status, &g.pst1, &g.pst2};
  for (auto b : iR) { rc = b->alloc(sizeof(int) * (size_t)R); if (rc) return rc; }
  rc = g.cnt.alloc(sizeof(int)); if (rc) return rc;
  rc = g.nf.alloc(sizeof(unsigned long long)); if (rc) return rc;
  rc = g.sums.alloc(sizeof(double) * (2 * (size_t)PD + 2)); if (rc) return rc;
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
    const uint64_t first = (uint64_t)first_replicate | (which ? ((uint64_t)CPIMH_MAX_ITERATION << 32) : 0ull);
    rc = cpimh_pf_run(pf, R, seed, first, s);
    if (rc) return rc;
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(which ? g.st2.p : g.st1.p, d_tr, RPD, cudaMemcpyDeviceToDevice, s));
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(which ? g.lz2.p : g.lz1.p, d_lz, sizeof(double) * (size_t)R, cudaMemcpyDeviceToDevice, s));
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(which ? g.pst2.p : g.pst1.p, d_st, sizeof(int) * (size_t)R, cudaMemcpyDeviceToDevice, s));
    if (rb) CPIMH_CUDA_CHECK(cudaMemcpyAsync(which ? g.e2.p : g.e1.p, d_ex, RPD, cudaMemcpyDeviceToDevice, s));
  }
  if (timing) { cudaStreamSynchronize(s); fprintf(stderr, "[ccpf] setup + init filters %.1f ms\n", now() - t_start); }
  const double t_chain = now();
  CcRound q;
  memset(&q, 0, sizeof(q));
  q.PD = PD; q.k = k; q.K = K; q.max_it = max_it; q.first_replicate = first_replicate;
  q.new1 = g.n1.as<double>(); q.new2 = g.n2.as<double>(); q.new_logz = g.nlz.as<double>(); q.new_status = g.nst.as<int>();
  q.st1 = g.st1.as<double>(); q.st2 = g.st2.as<double>(); q.lz1 = g.lz1.as<double>(); q.lz2 = g.lz2.as<double>();
  q.single = g.single.as<int>(); q.iter = g.iter.as<int>(); q.tau = g.tau.as<int>(); q.status = g.status.as<int>();
  q.accH = g.accH.as<double>(); q.accD = g.accD.as<double>(); q.hbar = g.hbar.as<double>(); q.bc = g.bc.as<double>();
  q.nfilters = g.nf.as<unsigned long long>();
  q.next_active = nullptr; q.next_count = nullptr;       // no active lists: the persistent kernel takes replicates from a ticket counter
  if (truncation) {
    rc = g.trunc.alloc(sizeof(int) * (size_t)R); if (rc) return rc;
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(g.trunc.p, truncation, sizeof(int) * (size_t)R, cudaMemcpyHostToDevice, s));
    q.trunc = g.trunc.as<int>(); q.rg_lambda = rg_lambda;
  }
  if (rb) {
    q.newe1 = g.ne1.as<double>(); q.newe2 = g.ne2.as<double>(); q.e1 = g.e1.as<double>(); q.e2 = g.e2.as<double>();
    q.accHe = g.accHe.as<double>(); q.accDe = g.accDe.as<double>(); q.hbar_rb = g.hbar_rb.as<double>();
  }
  CPIMH_CUDA_CHECK(cudaMemsetAsync(g.nf.p, 0, sizeof(unsigned long long), s));
  CPIMH_CUDA_CHECK(cudaMemsetAsync(g.cnt.p, 0, sizeof(int), s));
  cc_init_kernel<<<(unsigned)R, 128, 0, s>>>(q, g.pst1.as<int>(), g.pst2.as<int>());
  CcpfParams p;
  ccpf_fill(w, &p);
  p.with_as = with_as ? 1 : 0; p.seed = seed;
  p.single = g.single.as<int>();
  p.ref1 = g.st1.as<double>(); p.ref2 = g.st2.as<double>();
  p.traj1 = g.n1.as<double>(); p.traj2 = g.n2.as<double>(); p.logz = g.nlz.as<double>(); p.status = g.nst.as<int>();
  if (rb) { p.exp1 = g.ne1.as<double>(); p.exp2 = g.ne2.as<double>(); }
  CPIMH_CUDA_CHECK(cudaMemsetAsync(w.near.p, 0, sizeof(int), s));
  // ---- the chains: one persistent kernel, every resident CTA runs whole replicates (ccpf.cuh ccpf_chains_kernel)
  CcpfLaunch L;
  L.threads = w.threads; L.smem = w.smem; L.grid = (int)(R < w.grid_max ? R : w.grid_max);
  rc = w.entry->launch_ccpf_chains(p, q, R, g.cnt.as<int>(), L, s);
  if (rc) return rc;
  w.launches++;
  cc_reduce_kernel<<<PD + 1, 256, 0, s>>>(g.hbar.as<double>(), g.status.as<int>(), R, PD, g.nf.as<unsigned long long>(), g.sums.as<double>());
  if (rb) cc_reduce_kernel<<<PD + 1, 256, 0, s>>>(g.hbar_rb.as<double>(), g.status.as<int>(), R, PD, g.nf.as<unsigned long long>(), g.sums_rb.as<double>());
  CPIMH_CUDA_CHECK(cudaGetLastError());
  if (timing) { cudaStreamSynchronize(s); fprintf(stderr, "[ccpf] chain kernel + reduce %.1f ms\n", now() - t_chain); }
  const double t_out = now();
  std::vector<int> stv((size_t)R);
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(stv.data(), g.status.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost, s));
  int nrep = 0;
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(&nrep, w.near.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  g_last_exact_repeats = nrep;
  for (int64_t r = 0; r < R; r++)
    if (stv[(size_t)r] && stv[(size_t)r] != CPIMH_ERR_NOT_MET) { set_error("replicate %lld failed with status %d", (long long)r, stv[(size_t)r]); return stv[(size_t)r]; }
  std::vector<double> hs(2 * (size_t)PD + 2);
  CPIMH_CUDA_CHECK(cudaMemcpy(hs.data(), g.sums.p, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost));
  if (o->hbar) CPIMH_CUDA_CHECK(cudaMemcpy(o->hbar, g.hbar.p, RPD, cudaMemcpyDeviceToHost));
  if (o->bc) CPIMH_CUDA_CHECK(cudaMemcpy(o->bc, g.bc.p, RPD, cudaMemcpyDeviceToHost));
  if (o->meetingtime) CPIMH_CUDA_CHECK(cudaMemcpy(o->meetingtime, g.tau.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost));
  if (o->iteration) CPIMH_CUDA_CHECK(cudaMemcpy(o->iteration, g.iter.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost));
  if (o->sum_hbar) memcpy(o->sum_hbar, hs.data(), sizeof(double) * PD);
  if (o->sum_hbar_sq) memcpy(o->sum_hbar_sq, hs.data() + PD, sizeof(double) * PD);
  if (o->nfilters) *o->nfilters = (int64_t)hs[2 * (size_t)PD];
  if (o->nfinished) *o->nfinished = (int64_t)hs[2 * (size_t)PD + 1];
  if (o->finished) for (int64_t r = 0; r < R; r++) o->finished[r] = stv[(size_t)r] == 0 ? 1 : 0;
  if (rb) {
    CPIMH_CUDA_CHECK(cudaMemcpy(hs.data(), g.sums_rb.p, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost));
    if (o->hbar_rb) CPIMH_CUDA_CHECK(cudaMemcpy(o->hbar_rb, g.hbar_rb.p, RPD, cudaMemcpyDeviceToHost));
    if (o->sum_hbar_rb) memcpy(o->sum_hbar_rb, hs.data(), sizeof(double) * PD);
    if (o->sum_hbar_rb_sq) memcpy(o->sum_hbar_rb_sq, hs.data() + PD, sizeof(double) * PD);
  }
  if (timing) fprintf(stderr, "[ccpf] outputs %.1f ms, total %.1f ms\n", now() - t_out, now() - t_start);
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
  CPIMH_CUDA_CHECK(cached_device_properties(&prop, dev));
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
  DevBuf obs, theta, dpre, hx, ha, pr, ll, llt, pm, pa, rs, an, stt, nr;
  struct Rel { DevBuf* b[14]; ~Rel() { for (auto x : b) x->release(); } } rel{{&obs, &theta, &dpre, &hx, &ha, &pr, &ll, &llt, &pm, &pa, &rs, &an, &stt, &nr}};
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
  if (!rc) rc = nr.alloc(sizeof(int) * B);
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
  // exact ancestors: a step with a uniform within rounding distance of a CDF knot is repeated in reference order inside the kernel
  p.near = nr.as<int>();
  p.margin_rel = 12.0 * ((double)N + 2.0) * 0x1p-53 * margin_scale();
  CPIMH_CUDA_CHECK(cudaMemset(nr.p, 0, sizeof(int) * B));
  const char* force = getenv("CPIMH_APF_EXACT");                    // test aid: reference-order mode from the start
  p.exact = (force && atoi(force)) ? 1 : 0;
  rc = e->launch_apf(p, L, nullptr);
  if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaDeviceSynchronize());
  std::vector<int> st((size_t)B);
  CPIMH_CUDA_CHECK(cudaMemcpy(st.data(), nr.p, sizeof(int) * B, cudaMemcpyDeviceToHost));
  g_last_exact_repeats = 0;
  for (int64_t b = 0; b < B; b++) g_last_exact_repeats += st[(size_t)b];
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

int64_t cpimh_last_exact_repeats(void) { return g_last_exact_repeats; }

}  // extern "C"
