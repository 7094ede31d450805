// api_chains.cu -- C ABI for the coupled-PIMH driver (cpimh_chains_*, cpimh_coupled_pimh), the single model
// operations and the stand-alone ancestor tree (cpimh_tree_*).  See include/cpimh.h for the reference citations.
#include <math.h>
#include <string.h>
#include <vector>
#include "host_common.h"

namespace cpimh {

// deterministic reduction of the per-replicate estimators into the all-reduce buffer:
// sums[0..P) = sum_r hbar[r][j], sums[P..2P) = sum_r hbar[r][j]^2, sums[2P] = filters run, sums[2P+1] = replicates
__global__ void reduce_hbar_kernel(const double* hbar, const int* status, long long R, int P, const unsigned long long* nfilters, double* sums) {
  __shared__ double scr[32];
  const int j = blockIdx.x;
  if (j == P) {          // replicates that entered the sums: those that finished (status 0); capped ones (CPIMH_ERR_NOT_MET) are left out
    double n = 0;
    for (long long r = threadIdx.x; r < R; r += blockDim.x) n += (status[r] == 0) ? 1.0 : 0.0;
    n = block_sum(n, scr);
    if (threadIdx.x == 0) { sums[2 * P] = (double)(*nfilters); sums[2 * P + 1] = n; }
    return;
  }
  double s = 0, s2 = 0;
  for (long long r = threadIdx.x; r < R; r += blockDim.x) { if (status[r] != 0) continue; double v = hbar[(size_t)r * P + j]; s += v; s2 += v * v; }
  s = block_sum(s, scr);
  s2 = block_sum(s2, scr);
  if (threadIdx.x == 0) { sums[j] = s; sums[P + j] = s2; }
}

// ------------------------------------------------------------------ round-based coupled-PIMH driver
// PIMH proposals are independent of the chain state (R/pf_kernels.R:104-131: one CPF(), then compare log Z), so the
// proposals of iteration i+1..i+W of every unfinished replicate can be run as ONE batch of filters; a small kernel then
// replays the accept / meet / H_bar recursion (R/debiasing_algorithm.R:180-248, 296-315) over them in order and
// discards what lies beyond the stopping time max(tau, K).  W = 1 while the batch fills the GPU, then grows, so a
// replicate with a long meeting time no longer serialises its filters on one CTA.
struct RoundParams {
  long long R; int W, nA, PD, P, D, k, K, max_it;
  unsigned long long seed; unsigned first_replicate;
  const int* active; int* next_active; int* next_count;
  unsigned long long* stream_ids;
  const double* pool_logz; const double* pool_traj; const int* pool_status;
  double* lz1; double* lz2; int* id1; int* id2; int* meet; int* iter; int* tau; int* status;
  double* st1; double* st2; double* accH; double* accD; double* hbar; double* bc;
  unsigned long long* nfilters;
  // all_particles = TRUE (R/pf_kernels.R:147-215): the weighted-average path travels with the chain state; null when off
  const double* pool_exp; double* e1; double* e2; double* accHe; double* accDe; double* hbar_rb;
};
__global__ void rounds_prepare_kernel(RoundParams q) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= (long long)q.nA * q.W) return;
  const int rep = q.active ? q.active[s / q.W] : (int)(s / q.W);
  const unsigned long long it = (unsigned long long)(q.iter[rep] + 1 + (int)(s % q.W));
  q.stream_ids[s] = (unsigned long long)(q.first_replicate + (unsigned)rep) | (it << 32);
}
// pimh_init (R/pf_kernels.R:134-139): filter with iteration index 0 becomes chain 1; chain 2 does not exist yet
__global__ void rounds_init_kernel(RoundParams q) {
  const int rep = blockIdx.x;
  const double* prop = q.pool_traj + (size_t)rep * q.PD;
  for (int j = threadIdx.x; j < q.PD; j += blockDim.x) {
    const double v = prop[j];
    q.st1[(size_t)rep * q.PD + j] = v;
    q.accH[(size_t)rep * q.PD + j] = (q.k <= 0) ? v : 0.0;
    q.accD[(size_t)rep * q.PD + j] = 0.0;
    if (q.pool_exp) {                                                          // exp_samples1[1,] (debiasing_algorithm.R:166-170)
      const double e = q.pool_exp[(size_t)rep * q.PD + j];
      q.e1[(size_t)rep * q.PD + j] = e;
      q.accHe[(size_t)rep * q.PD + j] = (q.k <= 0) ? e : 0.0;
      q.accDe[(size_t)rep * q.PD + j] = 0.0;
    }
  }
  if (threadIdx.x == 0) {
    q.lz1[rep] = q.pool_logz[rep]; q.lz2[rep] = -INFINITY; q.id1[rep] = 0; q.id2[rep] = -1; q.meet[rep] = 0; q.iter[rep] = 0; q.tau[rep] = -1;
    int st = q.pool_status[rep];
    if (st == 0 && isnan(q.pool_logz[rep])) st = CPIMH_ERR_INVALID;
    q.status[rep] = st;
    if (st == 0) q.next_active[atomicAdd(q.next_count, 1)] = rep;
  }
}
__global__ void rounds_accept_kernel(RoundParams q) {
  const int a = blockIdx.x;
  const int rep = q.active[a];
  const int PD = q.PD;
  double* x1 = q.st1 + (size_t)rep * PD;
  double* x2 = q.st2 + (size_t)rep * PD;
  double* accH = q.accH + (size_t)rep * PD;
  double* accD = q.accD + (size_t)rep * PD;
  double lz1 = q.lz1[rep], lz2 = q.lz2[rep];
  int id1 = q.id1[rep], id2 = q.id2[rep], meet = q.meet[rep], iter = q.iter[rep], tau = q.tau[rep], status = 0, finished = 0;
  for (int w = 0; w < q.W && !finished; w++) {
    const size_t s = (size_t)a * q.W + w;
    iter++;
    status = q.pool_status[s];
    const double lzp = q.pool_logz[s];
    if (status == 0 && isnan(lzp)) status = CPIMH_ERR_INVALID;
    if (status) { finished = 1; break; }
    const Stream st = make_stream(q.seed, (unsigned long long)(q.first_replicate + (unsigned)rep) | ((unsigned long long)iter << 32));
    const double lu = log(philox_u1(st, 0, 0, CPIMH_P_MH, 0));               // R/pf_kernels.R:112
    const bool a1 = lu < (lzp - lz1);                                          // :113
    const bool a2 = meet ? a1 : (lu < (lzp - lz2));                            // :121
    const double* prop = q.pool_traj + s * PD;
    if (a1) { lz1 = lzp; id1 = iter; }
    if (meet) { lz2 = lz1; id2 = id1; }
    else {
      if (a2) { lz2 = lzp; id2 = iter; }
      if (id1 == id2) { meet = 1; tau = iter; }                                // debiasing_algorithm.R:216-220
    }
    const int t = iter - 1;
    const bool mcmc = (iter >= q.k && iter <= q.K);
    const bool bias = (t >= q.k) && (id1 != id2);
    const double coef = (double)min(t - q.k + 1, q.K - q.k + 1);
    for (int j = threadIdx.x; j < PD; j += blockDim.x) {
      const double pv = prop[j];
      const double xa = a1 ? pv : x1[j];
      const double xb = a2 ? pv : x2[j];
      if (a1) x1[j] = pv;
      if (a2) x2[j] = pv;
      if (mcmc) accH[j] += xa;                                                 // :296-302
      if (bias) accD[j] = accD[j] + coef * (xa - xb);                          // :308-312
    }
    if (q.pool_exp) {                                                          // exp_path follows the same accept decisions (pf_kernels.R:176-190)
      const double* pe = q.pool_exp + s * PD;
      double* e1 = q.e1 + (size_t)rep * PD; double* e2 = q.e2 + (size_t)rep * PD;
      double* aHe = q.accHe + (size_t)rep * PD; double* aDe = q.accDe + (size_t)rep * PD;
      for (int j = threadIdx.x; j < PD; j += blockDim.x) {
        const double pv = pe[j];
        const double xa = a1 ? pv : e1[j];
        const double xb = a2 ? pv : e2[j];
        if (a1) e1[j] = pv;
        if (a2) e2[j] = pv;
        if (mcmc) aHe[j] += xa;
        if (bias) aDe[j] = aDe[j] + coef * (xa - xb);
      }
    }
    if (tau > 0 && iter >= max(tau, q.K)) finished = 1;                        // :245
    if (!finished && q.max_it >= 0 && iter >= q.max_it) { finished = 1; status = CPIMH_ERR_NOT_MET; }   // R: finished = FALSE
  }
  __syncthreads();
  if (finished) {
    const double denom = (double)(q.K - q.k + 1);
    for (int j = threadIdx.x; j < PD; j += blockDim.x) {
      q.hbar[(size_t)rep * PD + j] = (accH[j] + accD[j]) / denom;   // :315
      q.bc[(size_t)rep * PD + j] = accD[j] / denom;                 // BC(), :330-363
      if (q.pool_exp) q.hbar_rb[(size_t)rep * PD + j] = (q.accHe[(size_t)rep * PD + j] + q.accDe[(size_t)rep * PD + j]) / denom;
    }
  }
  if (threadIdx.x == 0) {
    q.lz1[rep] = lz1; q.lz2[rep] = lz2; q.id1[rep] = id1; q.id2[rep] = id2; q.meet[rep] = meet; q.iter[rep] = iter; q.tau[rep] = tau;
    if (status) q.status[rep] = status;
    if (finished) atomicAdd(q.nfilters, (unsigned long long)(1 + iter));
    else q.next_active[atomicAdd(q.next_count, 1)] = rep;
  }
}

// ------------------------------------------------------------------ stand-alone tree kernels (class Tree, src/tree.cpp)
__global__ void tree_init_kernel(int N, int d, const double* x0, int* a_star, double* x_star, int* l_star) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  a_star[i] = -2;                                                  // :30
  for (int k = 0; k < d; k++) x_star[(size_t)i * d + k] = x0[(size_t)i * d + k];   // :31
  l_star[i] = i;                                                   // :32
}
// update part 1: offspring counts (:90-94), prune (:70-86), b = l_star[a] (:39-42)
__global__ void tree_prune_kernel(int N, const int* a, int* o, int* a_star, int* o_star, const int* l_star, int* b) {
  // single CTA: grid-wide ordering between the histogram, the scatter and the walks comes from __syncthreads
  for (int i = threadIdx.x; i < N; i += blockDim.x) o[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) { atomicAdd(&o[a[i]], 1); b[i] = l_star[a[i]]; }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) o_star[l_star[i]] = o[i];
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    if (o[i] != 0) continue;
    int j = a_star[l_star[i]];
    while (j >= 0) {
      int old = atomicSub(&o_star[j], 1);
      if (old != 1) break;
      j = a_star[j];
    }
  }
}
// update part 2: first N slots with o_star == 0 in index order (:45-53); found < N => caller doubles (:54-61)
__global__ void tree_find_slots_kernel(int N, int M, const int* o_star, int* l_new, int* found_out) {
  __shared__ int scr[32];
  int found = 0;
  for (int base = 0; base < M && found < N; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int isfree = (i < M && o_star[i] == 0) ? 1 : 0;
    int tot;
    const int r = found + block_excl_scan_int(isfree, scr, &tot);
    if (isfree && r < N) l_new[r] = i;
    found += tot;
  }
  if (threadIdx.x == 0) *found_out = found < N ? found : N;
}
__global__ void tree_fill_overflow_kernel(int N, int found, int old_M, int* l_new) {
  const int i = found + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) l_new[i] = old_M + i - found;                         // :58-60
}
__global__ void tree_scatter_kernel(int N, int d, const double* x, const int* b, const int* l_new, int* a_star, double* x_star, int* l_star) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int s = l_new[i];
  a_star[s] = b[i];                                                // :65
  for (int k = 0; k < d; k++) x_star[(size_t)s * d + k] = x[(size_t)i * d + k];   // :66
  l_star[i] = s;
}
__global__ void tree_get_path_kernel(int d, int nsteps, int n, const int* a_star, const double* x_star, const int* l_star, double* path) {
  int j = l_star[n];
  for (int step = nsteps; step >= 0; step--) {                     // :120-127
    for (int k = threadIdx.x; k < d; k += blockDim.x) path[(size_t)step * d + k] = x_star[(size_t)j * d + k];
    j = a_star[j];
  }
}
__global__ void tree_xgeneration_kernel(int N, int d, int lag, const int* a_star, const double* x_star, const int* l_star, double* xgen) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int j = l_star[i];
  for (int s = 0; s < lag; s++) j = a_star[j];                     // :134-137
  for (int k = 0; k < d; k++) xgen[(size_t)i * d + k] = x_star[(size_t)j * d + k];
}
// all-particle outputs of a finished filter (R/CPF.R:73-76, R/particlefilter.R:39-42): every leaf walks to the root
// in lock step; per generation the weighted average (pf_get_weighted_average, R/CPF.R:84-92) is a block reduction.
__global__ void all_paths_kernel(AllParticlesArgs a, double* exp_path /* [T+1][D] or null */, double* paths /* [N][T+1][D] or null */, int* cur) {
  __shared__ double scr[32];
  const int N = a.N, D = a.D, T = a.T;
  const bool tree = a.path_mode == PATH_TREE;
  for (int n = threadIdx.x; n < N; n += blockDim.x) cur[n] = tree ? a.leaf[n] : n;
  __syncthreads();
  for (int step = T; step >= 0; step--) {
    for (int q = 0; q < D; q++) {
      double acc = 0;
      for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const double v = tree ? a.x_star[(size_t)q * a.M + cur[n]] : a.hist_x[((size_t)step * D + q) * a.Npad + cur[n]];
        if (paths) paths[((size_t)n * (T + 1) + step) * D + q] = v;
        acc += v * a.normw[n];
      }
      if (exp_path) { acc = block_sum(acc, scr); if (threadIdx.x == 0) exp_path[(size_t)step * D + q] = acc; }
    }
    __syncthreads();
    if (step > 0) for (int n = threadIdx.x; n < N; n += blockDim.x) cur[n] = tree ? a.a_star[cur[n]] : a.hist_a[(size_t)(step - 1) * a.Npad + cur[n]];
    __syncthreads();
  }
}

}  // namespace cpimh

using namespace cpimh;

// ============================================================================ chains handle
struct cpimh_chains {
  cpimh_config cfg;
  const ModelEntry* entry;
  Geometry geo;
  int64_t maxR, lastR, launches;
  int record, max_rec, grid, npre;
  DevBuf hist_x, hist_a, ascr, wscr, jq;
  DevBuf bc;
  int all_particles;
  DevBuf r_exp, r_e1, r_e2, r_accHe, r_accDe, hbar_rb, sums_rb;
  DevBuf obs, theta, pre, xbuf, a_star, o_star, x_star, state, acc, estate, eacc, hbar, mt, it, status, nfilters, counter, sums;
  DevBuf ring_traj, ring_exp, ring_logz, ring_status, ring_tag, next_it, it_done, slot_rep, unfinished;   // persistent kernel: proposal rings per resident CTA
  DevBuf samples1, samples2, log_pdf1, log_pdf2;
  // round-based driver
  int mode; int64_t filters_launched;
  Geometry geo_pf; int grid_pf;
  DevBuf r_logz, r_traj, r_status, r_ids, r_lz1, r_lz2, r_id1, r_id2, r_meet, r_iter, r_tau, r_st1, r_st2, r_accH, r_accD, r_act0, r_act1, r_cnt;
};

struct cpimh_tree {
  int N, M, dimx, nsteps;
  DevBuf a_star, o_star, x_star, l_star, l_new, b, o, x, a, found, out;
};

extern "C" {

int cpimh_chains_destroy(cpimh_chains* h) {
  if (!h) return CPIMH_OK;
  DevBuf* all[] = {&h->bc, &h->r_exp, &h->r_e1, &h->r_e2, &h->r_accHe, &h->r_accDe, &h->hbar_rb, &h->sums_rb, &h->hist_x, &h->hist_a, &h->ascr, &h->wscr, &h->jq, &h->obs, &h->theta, &h->pre, &h->xbuf, &h->a_star, &h->o_star, &h->x_star, &h->state, &h->acc, &h->estate, &h->eacc, &h->ring_traj, &h->ring_exp, &h->ring_logz, &h->ring_status, &h->ring_tag, &h->next_it, &h->it_done, &h->slot_rep, &h->unfinished, &h->hbar, &h->mt,
                   &h->it, &h->status, &h->nfilters, &h->counter, &h->sums, &h->samples1, &h->samples2, &h->log_pdf1, &h->log_pdf2,
                   &h->r_logz, &h->r_traj, &h->r_status, &h->r_ids, &h->r_lz1, &h->r_lz2, &h->r_id1, &h->r_id2, &h->r_meet, &h->r_iter, &h->r_tau,
                   &h->r_st1, &h->r_st2, &h->r_accH, &h->r_accD, &h->r_act0, &h->r_act1, &h->r_cnt};
  for (auto b : all) b->release();
  delete h;
  return CPIMH_OK;
}

int cpimh_chains_create(const cpimh_config* cfg_in, int64_t max_replicates, int record_samples, cpimh_chains** out) {
  CPIMH_REQUIRE(cfg_in && out, "null argument");
  CPIMH_REQUIRE(max_replicates >= 1, "max_replicates must be >= 1");
  cpimh_config cfg = *cfg_in;
  if (cfg.store_paths == 0) cfg.store_paths = 1;   // chain states are trajectories
  int rc = api_validate(&cfg);
  if (rc) return rc;
  const ModelEntry* e = find_entry(cfg.model, cfg.dim);
  std::vector<double> pre = api_precompute(cfg);
  Geometry g;
  rc = api_plan_geometry(cfg, e, true, max_replicates, (int)pre.size(), &g);
  if (rc) return rc;
  cpimh_chains* h = new cpimh_chains();
  h->cfg = cfg; h->entry = e; h->geo = g; h->maxR = max_replicates; h->lastR = 0; h->launches = 0;
  h->record = record_samples > 0; h->max_rec = record_samples > 0 ? record_samples : 0;
  h->mode = 0; h->filters_launched = 0; h->grid_pf = 0; h->all_particles = 0;
  const int64_t resident = (int64_t)g.blocks_per_sm * g.num_sms;
  h->grid = (int)(max_replicates < resident ? max_replicates : resident);
  const size_t G = (size_t)h->grid, D = cfg.dim, T = cfg.nobs, P = T + 1, R = (size_t)max_replicates;
  h->npre = (int)pre.size();
  rc = h->obs.alloc(sizeof(double) * T * cfg.dim_y);
  if (!rc) rc = h->theta.alloc(sizeof(double) * CPIMH_MAX_THETA);
  if (!rc) rc = h->pre.alloc(sizeof(double) * (pre.size() + 1));
  if (!rc && g.path_mode == PATH_DENSE) {
    rc = h->hist_x.alloc(sizeof(double) * G * g.hist_x_elems);
    if (!rc) rc = h->hist_a.alloc(sizeof(int) * G * g.hist_a_elems);
  }
  if (!rc && g.path_mode == PATH_TREE) {
    rc = h->xbuf.alloc(sizeof(double) * G * 2 * D * g.Npad);
    if (!rc) rc = h->ascr.alloc(sizeof(int) * G * g.Npad);
    if (!rc) rc = h->a_star.alloc(sizeof(int) * G * g.M);
    if (!rc) rc = h->o_star.alloc(sizeof(int) * G * g.M);
    if (!rc) rc = h->x_star.alloc(sizeof(double) * G * D * g.M);
  }
  if (!rc) rc = h->wscr.alloc(sizeof(double) * G * g.Npad);
  if (!rc && cfg.model == CPIMH_MODEL_LEVY_SV) rc = h->jq.alloc(sizeof(uint4) * G * pf_jq_stride(g.Npad));
  if (!rc) rc = h->state.alloc(sizeof(double) * G * 2 * P * D);
  if (!rc) rc = h->acc.alloc(sizeof(double) * G * 2 * P * D);
  if (!rc) rc = h->ring_traj.alloc(sizeof(double) * G * CPIMH_RING * P * D);
  if (!rc) rc = h->ring_logz.alloc(sizeof(double) * G * CPIMH_RING);
  if (!rc) rc = h->ring_status.alloc(sizeof(int) * G * CPIMH_RING);
  if (!rc) rc = h->ring_tag.alloc(sizeof(int) * G * CPIMH_RING);
  if (!rc) rc = h->next_it.alloc(sizeof(int) * G);
  if (!rc) rc = h->it_done.alloc(sizeof(int) * G);
  if (!rc) rc = h->slot_rep.alloc(sizeof(int) * G);
  if (!rc) rc = h->unfinished.alloc(sizeof(int));
  if (!rc) rc = h->hbar.alloc(sizeof(double) * R * P * D);
  if (!rc) rc = h->bc.alloc(sizeof(double) * R * P * D);
  if (!rc) rc = h->mt.alloc(sizeof(int) * R);
  if (!rc) rc = h->it.alloc(sizeof(int) * R);
  if (!rc) rc = h->status.alloc(sizeof(int) * R);
  if (!rc) rc = h->nfilters.alloc(2 * sizeof(unsigned long long));
  if (!rc) rc = h->counter.alloc(sizeof(int));
  if (!rc) rc = h->sums.alloc(sizeof(double) * (2 * P * D + 2));
  if (!rc && h->record) {
    rc = h->samples1.alloc(sizeof(double) * R * (h->max_rec + 1) * P);
    if (!rc) rc = h->samples2.alloc(sizeof(double) * R * h->max_rec * P);
    if (!rc) rc = h->log_pdf1.alloc(sizeof(double) * R * (h->max_rec + 1));
    if (!rc) rc = h->log_pdf2.alloc(sizeof(double) * R * h->max_rec);
  }
  if (rc) { cpimh_chains_destroy(h); return rc; }
  cudaError_t ce = cudaMemcpy(h->theta.p, cfg.theta, sizeof(double) * CPIMH_MAX_THETA, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess && !pre.empty()) ce = cudaMemcpy(h->pre.p, pre.data(), sizeof(double) * pre.size(), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { set_error("cudaMemcpy failed: %s", cudaGetErrorString(ce)); cpimh_chains_destroy(h); return CPIMH_ERR_CUDA; }
  *out = h;
  return CPIMH_OK;
}

int cpimh_chains_set_observations(cpimh_chains* h, const double* obs_host) {
  CPIMH_REQUIRE(h && obs_host, "null argument");
  return api_upload_obs(h->cfg, obs_host, h->obs.p);
}

static int chains_rounds_alloc(cpimh_chains* h) {
  if (h->grid_pf > 0) return CPIMH_OK;
  std::vector<double> pre = api_precompute(h->cfg);
  Geometry g;
  int rc = api_plan_geometry(h->cfg, h->entry, false, h->maxR, (int)pre.size(), &g);
  if (rc) return rc;
  CPIMH_REQUIRE(g.path_mode == h->geo.path_mode && g.M == h->geo.M, "internal: filter and chain kernels planned different path storage");
  const int64_t resident = (int64_t)g.blocks_per_sm * g.num_sms;
  const int grid = (int)(h->maxR < resident ? h->maxR : resident);
  const size_t G = (size_t)grid, D = h->cfg.dim, T = h->cfg.nobs, PD = (T + 1) * D, R = (size_t)h->maxR;
  // the filter kernel may be resident with more CTAs than the chain kernel: grow the per-CTA workspaces
  if (g.path_mode == PATH_DENSE) {
    rc = h->hist_x.alloc(sizeof(double) * G * g.hist_x_elems);
    if (!rc) rc = h->hist_a.alloc(sizeof(int) * G * g.hist_a_elems);
  } else {
    rc = h->xbuf.alloc(sizeof(double) * G * 2 * D * g.Npad);
    if (!rc) rc = h->ascr.alloc(sizeof(int) * G * g.Npad);
    if (!rc) rc = h->a_star.alloc(sizeof(int) * G * g.M);
    if (!rc) rc = h->o_star.alloc(sizeof(int) * G * g.M);
    if (!rc) rc = h->x_star.alloc(sizeof(double) * G * D * g.M);
  }
  if (!rc) rc = h->wscr.alloc(sizeof(double) * G * g.Npad);
  if (!rc && h->cfg.model == CPIMH_MODEL_LEVY_SV) rc = h->jq.alloc(sizeof(uint4) * G * pf_jq_stride(g.Npad));
  if (!rc) rc = h->r_logz.alloc(sizeof(double) * R);
  if (!rc) rc = h->r_traj.alloc(sizeof(double) * R * PD);
  if (!rc) rc = h->r_status.alloc(sizeof(int) * R);
  if (!rc) rc = h->r_ids.alloc(sizeof(unsigned long long) * R);
  if (!rc) rc = h->r_lz1.alloc(sizeof(double) * R);
  if (!rc) rc = h->r_lz2.alloc(sizeof(double) * R);
  if (!rc) rc = h->r_id1.alloc(sizeof(int) * R);
  if (!rc) rc = h->r_id2.alloc(sizeof(int) * R);
  if (!rc) rc = h->r_meet.alloc(sizeof(int) * R);
  if (!rc) rc = h->r_iter.alloc(sizeof(int) * R);
  if (!rc) rc = h->r_tau.alloc(sizeof(int) * R);
  if (!rc) rc = h->r_st1.alloc(sizeof(double) * R * PD);
  if (!rc) rc = h->r_st2.alloc(sizeof(double) * R * PD);
  if (!rc) rc = h->r_accH.alloc(sizeof(double) * R * PD);
  if (!rc) rc = h->r_accD.alloc(sizeof(double) * R * PD);
  if (!rc) rc = h->r_act0.alloc(sizeof(int) * R);
  if (!rc) rc = h->r_act1.alloc(sizeof(int) * R);
  if (!rc) rc = h->r_cnt.alloc(sizeof(int));
  if (rc) return rc;
  h->geo_pf = g; h->grid_pf = grid;
  return CPIMH_OK;
}

static int chains_run_rounds(cpimh_chains* h, int64_t R, uint64_t seed, uint32_t first_replicate, int k, int K, int max_it, cudaStream_t s) {
  int rc = chains_rounds_alloc(h);
  if (rc) return rc;
  h->filters_launched = 0;
  const cpimh_config& c = h->cfg;
  const Geometry& g = h->geo_pf;
  const int D = c.dim, P = c.nobs + 1, PD = P * D;
  PfParams p;
  api_fill_pf_params(c, g, h->npre, &p);
  p.seed = seed; p.first_filter = 0; p.b0 = 0;
  p.obs = h->obs.as<double>(); p.theta = h->theta.as<double>(); p.pre = h->pre.as<double>();
  p.hist_x = h->hist_x.as<double>(); p.hist_a = h->hist_a.as<int>(); p.ascr = h->ascr.as<int>(); p.wscr = h->wscr.as<double>(); p.jq = h->jq.as<uint4>();
  p.xbuf = h->xbuf.as<double>(); p.a_star = h->a_star.as<int>(); p.o_star = h->o_star.as<int>(); p.x_star = h->x_star.as<double>();
  p.logz = h->r_logz.as<double>(); p.traj = h->r_traj.as<double>(); p.status = h->r_status.as<int>();
  p.stream_ids = h->r_ids.as<unsigned long long>();
  p.exp_path = h->all_particles ? h->r_exp.as<double>() : nullptr;
  RoundParams q;
  memset(&q, 0, sizeof(q));
  q.R = R; q.PD = PD; q.P = P; q.D = D; q.k = k; q.K = K; q.max_it = max_it; q.seed = seed; q.first_replicate = first_replicate;
  q.stream_ids = h->r_ids.as<unsigned long long>();
  q.pool_logz = h->r_logz.as<double>(); q.pool_traj = h->r_traj.as<double>(); q.pool_status = h->r_status.as<int>();
  q.lz1 = h->r_lz1.as<double>(); q.lz2 = h->r_lz2.as<double>(); q.id1 = h->r_id1.as<int>(); q.id2 = h->r_id2.as<int>();
  q.meet = h->r_meet.as<int>(); q.iter = h->r_iter.as<int>(); q.tau = h->r_tau.as<int>(); q.status = h->status.as<int>();
  q.st1 = h->r_st1.as<double>(); q.st2 = h->r_st2.as<double>(); q.accH = h->r_accH.as<double>(); q.accD = h->r_accD.as<double>();
  q.hbar = h->hbar.as<double>(); q.bc = h->bc.as<double>(); q.nfilters = h->nfilters.as<unsigned long long>(); q.next_count = h->r_cnt.as<int>();
  if (h->all_particles) {
    q.pool_exp = h->r_exp.as<double>(); q.e1 = h->r_e1.as<double>(); q.e2 = h->r_e2.as<double>();
    q.accHe = h->r_accHe.as<double>(); q.accDe = h->r_accDe.as<double>(); q.hbar_rb = h->hbar_rb.as<double>();
  }
  auto launch_filters = [&](int64_t B) -> int {
    PfLaunch L;
    L.threads = g.threads; L.smem = g.smem; L.variant = g.variant;
    L.grid = (int)(B < h->grid_pf ? B : h->grid_pf);
    p.B = B;
    h->filters_launched += B; h->launches += 1;
    return h->entry->launch_pf(p, L, s);
  };
  // iteration 0: one filter per replicate (stream ids = replicate ids with iteration index 0)
  CPIMH_CUDA_CHECK(cudaMemsetAsync(h->r_iter.p, 0xff, sizeof(int) * (size_t)R, s));     // iter = -1 so that prepare() yields iteration 0
  CPIMH_CUDA_CHECK(cudaMemsetAsync(h->r_cnt.p, 0, sizeof(int), s));
  q.active = nullptr; q.next_active = h->r_act0.as<int>(); q.nA = (int)R; q.W = 1;
  rounds_prepare_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(q);
  rc = launch_filters(R);
  if (rc) return rc;
  rounds_init_kernel<<<(unsigned)R, 128, 0, s>>>(q);
  h->launches += 2;
  int nA = 0, flip = 0;
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(&nA, h->r_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  while (nA > 0) {
    // window: one proposal per replicate while that fills the GPU, more (speculatively) once few replicates remain
    int64_t W = h->grid_pf / nA;
    if (W < 1) W = 1;
    if (W > 32) W = 32;
    if (W * nA > R) W = R / nA;
    q.active = flip ? h->r_act1.as<int>() : h->r_act0.as<int>();
    q.next_active = flip ? h->r_act0.as<int>() : h->r_act1.as<int>();
    q.nA = nA; q.W = (int)W;
    CPIMH_CUDA_CHECK(cudaMemsetAsync(h->r_cnt.p, 0, sizeof(int), s));
    rounds_prepare_kernel<<<(unsigned)((nA * W + 255) / 256), 256, 0, s>>>(q);
    rc = launch_filters(nA * W);
    if (rc) return rc;
    rounds_accept_kernel<<<(unsigned)nA, 128, 0, s>>>(q);
    h->launches += 2;
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(&nA, h->r_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
    flip ^= 1;
  }
  // per-replicate scalars in the persistent kernel's output arrays
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(h->mt.p, h->r_tau.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToDevice, s));
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(h->it.p, h->r_iter.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToDevice, s));
  CPIMH_CUDA_CHECK(cudaGetLastError());
  return CPIMH_OK;
}

int cpimh_chains_run(cpimh_chains* h, int64_t R, uint64_t seed, uint32_t first_replicate, int k, int K, int max_iterations,
                     double* d_sums, void* stream) {
  CPIMH_REQUIRE(h, "null handle");
  CPIMH_REQUIRE(R >= 1 && R <= h->maxR, "nreplicates %lld outside [1, %lld]", (long long)R, (long long)h->maxR);
  CPIMH_REQUIRE(k >= 0 && K >= k, "need 0 <= k <= K (got k=%d K=%d)", k, K);
  CPIMH_REQUIRE(max_iterations < 0 || max_iterations >= K, "max_iterations must be >= K or negative (unbounded)");
  CPIMH_REQUIRE(max_iterations < CPIMH_MAX_ITERATION, "iteration index must fit 20 bits of the Philox counter (max_iterations < %d)", CPIMH_MAX_ITERATION);
  cudaStream_t s = (cudaStream_t)stream;
  const bool rounds = h->mode == 2;       // default (0) and 1: the persistent owner / helper kernel (chains.cuh)
  if (rounds) {
    CPIMH_CUDA_CHECK(cudaMemsetAsync(h->nfilters.p, 0, 2 * sizeof(unsigned long long), s));
    CPIMH_CUDA_CHECK(cudaMemsetAsync(h->status.p, 0, sizeof(int) * (size_t)R, s));
    int rcr = chains_run_rounds(h, R, seed, first_replicate, k, K, max_iterations < 0 ? CPIMH_MAX_ITERATION - 1 : max_iterations, s);
    if (rcr) return rcr;
    const int Pr = (h->cfg.nobs + 1) * h->cfg.dim;
    double* sums_r = d_sums ? d_sums : h->sums.as<double>();
    reduce_hbar_kernel<<<Pr + 1, 256, 0, s>>>(h->hbar.as<double>(), h->status.as<int>(), R, Pr, h->nfilters.as<unsigned long long>(), sums_r);
    CPIMH_CUDA_CHECK(cudaGetLastError());
    if (d_sums) CPIMH_CUDA_CHECK(cudaMemcpyAsync(h->sums.p, d_sums, sizeof(double) * (2 * Pr + 2), cudaMemcpyDeviceToDevice, s));
    if (h->all_particles) {
      reduce_hbar_kernel<<<Pr + 1, 256, 0, s>>>(h->hbar_rb.as<double>(), h->status.as<int>(), R, Pr, h->nfilters.as<unsigned long long>(), h->sums_rb.as<double>());
      CPIMH_CUDA_CHECK(cudaGetLastError());
      h->launches += 1;
    }
    h->launches += 1; h->lastR = R;
    return CPIMH_OK;
  }
  ChainParams cp;
  memset(&cp, 0, sizeof(cp));
  api_fill_pf_params(h->cfg, h->geo, h->npre, &cp.pf);
  cp.pf.B = 0; cp.pf.seed = seed; cp.pf.first_filter = 0;
  cp.pf.obs = h->obs.as<double>(); cp.pf.theta = h->theta.as<double>(); cp.pf.pre = h->pre.as<double>();
  cp.pf.hist_x = h->hist_x.as<double>(); cp.pf.hist_a = h->hist_a.as<int>(); cp.pf.ascr = h->ascr.as<int>(); cp.pf.wscr = h->wscr.as<double>(); cp.pf.jq = h->jq.as<uint4>();
  cp.pf.xbuf = h->xbuf.as<double>(); cp.pf.a_star = h->a_star.as<int>(); cp.pf.o_star = h->o_star.as<int>(); cp.pf.x_star = h->x_star.as<double>();
  // the proposal rings are the filter's per-filter outputs (entry = slot * RING + iteration mod RING)
  cp.pf.traj = h->ring_traj.as<double>(); cp.pf.logz = h->ring_logz.as<double>(); cp.pf.status = h->ring_status.as<int>();
  cp.pf.exp_path = h->all_particles ? h->ring_exp.as<double>() : nullptr;
  cp.R = R; cp.first_replicate = first_replicate; cp.k = k; cp.K = K; cp.max_iterations = max_iterations < 0 ? CPIMH_MAX_ITERATION - 1 : max_iterations;
  cp.state = h->state.as<double>(); cp.acc = h->acc.as<double>();
  cp.estate = h->all_particles ? h->estate.as<double>() : nullptr; cp.eacc = h->all_particles ? h->eacc.as<double>() : nullptr;
  cp.ring_tag = h->ring_tag.as<int>(); cp.next_iter = h->next_it.as<int>(); cp.iter_done = h->it_done.as<int>(); cp.slot_rep = h->slot_rep.as<int>();
  cp.unfinished = h->unfinished.as<int>();
  cp.hbar = h->hbar.as<double>(); cp.bc = h->bc.as<double>(); cp.hbar_rb = h->all_particles ? h->hbar_rb.as<double>() : nullptr;
  cp.meetingtime = h->mt.as<int>(); cp.iteration = h->it.as<int>(); cp.status = h->status.as<int>();
  cp.nfilters = h->nfilters.as<unsigned long long>(); cp.work_counter = h->counter.as<int>();
  if (h->record) {
    cp.samples1 = h->samples1.as<double>(); cp.samples2 = h->samples2.as<double>();
    cp.log_pdf1 = h->log_pdf1.as<double>(); cp.log_pdf2 = h->log_pdf2.as<double>(); cp.max_rec = h->max_rec;
  }
  const int Ri = (int)R;
  CPIMH_CUDA_CHECK(cudaMemsetAsync(h->nfilters.p, 0, 2 * sizeof(unsigned long long), s));
  CPIMH_CUDA_CHECK(cudaMemsetAsync(h->counter.p, 0, sizeof(int), s));
  CPIMH_CUDA_CHECK(cudaMemsetAsync(h->slot_rep.p, 0xff, sizeof(int) * (size_t)h->grid, s));      // -1: no slot owns a replicate yet
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(h->unfinished.p, &Ri, sizeof(int), cudaMemcpyHostToDevice, s));  // pageable source: copied before the call returns
  PfLaunch L;
  L.threads = h->geo.threads; L.smem = h->geo.smem; L.variant = h->geo.variant;
  L.grid = h->grid;        // every resident CTA takes part: owners while replicates last, then helpers
  int rc = h->entry->launch_chains(cp, L, s);
  if (rc) return rc;
  const int P = (h->cfg.nobs + 1) * h->cfg.dim;     // estimator entries per replicate: d x (T+1)
  double* sums = d_sums ? d_sums : h->sums.as<double>();
  reduce_hbar_kernel<<<P + 1, 256, 0, s>>>(h->hbar.as<double>(), h->status.as<int>(), R, P, h->nfilters.as<unsigned long long>(), sums);
  CPIMH_CUDA_CHECK(cudaGetLastError());
  if (d_sums) CPIMH_CUDA_CHECK(cudaMemcpyAsync(h->sums.p, d_sums, sizeof(double) * (2 * P + 2), cudaMemcpyDeviceToDevice, s));
  if (h->all_particles) {
    reduce_hbar_kernel<<<P + 1, 256, 0, s>>>(h->hbar_rb.as<double>(), h->status.as<int>(), R, P, h->nfilters.as<unsigned long long>(), h->sums_rb.as<double>());
    CPIMH_CUDA_CHECK(cudaGetLastError());
    h->launches += 1;
  }
  h->launches += 2; h->lastR = R; h->filters_launched = -1;
  return CPIMH_OK;
}

int cpimh_chains_set_all_particles(cpimh_chains* h, int on) {
  CPIMH_REQUIRE(h, "null handle");
  h->all_particles = on ? 1 : 0;
  if (!on) return CPIMH_OK;
  const size_t PD = (size_t)(h->cfg.nobs + 1) * h->cfg.dim, G = (size_t)h->grid;
  int rc = h->hbar_rb.alloc(sizeof(double) * (size_t)h->maxR * PD);
  if (!rc) rc = h->estate.alloc(sizeof(double) * G * 2 * PD);
  if (!rc) rc = h->eacc.alloc(sizeof(double) * G * 2 * PD);
  if (!rc) rc = h->ring_exp.alloc(sizeof(double) * G * CPIMH_RING * PD);
  if (!rc) rc = h->sums_rb.alloc(sizeof(double) * (2 * PD + 2));
  if (!rc && h->mode == 2) {                  // the round-based driver keeps per-replicate copies
    DevBuf* bufs[] = {&h->r_exp, &h->r_e1, &h->r_e2, &h->r_accHe, &h->r_accDe};
    for (auto b : bufs) { rc = b->alloc(sizeof(double) * (size_t)h->maxR * PD); if (rc) break; }
  }
  return rc;
}

int cpimh_chains_set_mode(cpimh_chains* h, int mode) {
  CPIMH_REQUIRE(h && mode >= 0 && mode <= 2, "mode must be 0 (default: persistent kernel), 1 (the same) or 2 (host-driven rounds, round 1's driver)");
  CPIMH_REQUIRE(!(mode == 2 && h->record), "recording samples needs the persistent kernel (mode 0 / 1)");
  h->mode = mode;
  if (mode == 2 && h->all_particles) return cpimh_chains_set_all_particles(h, 1);
  return CPIMH_OK;
}
int64_t cpimh_chains_filters_launched(const cpimh_chains* h) {     // including discarded speculative proposals
  if (!h) return 0;
  if (h->filters_launched >= 0) return h->filters_launched;          // round-based driver: counted on the host
  unsigned long long n[2] = {0, 0};                                  // persistent kernel: counted on the device
  if (cudaMemcpy(n, h->nfilters.p, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)n[1];
}

int cpimh_chains_fetch(cpimh_chains* h, int64_t R, void* stream, cpimh_chain_out* o) {
  CPIMH_REQUIRE(h && o, "null argument");
  CPIMH_REQUIRE(R >= 1 && R <= h->lastR, "nreplicates %lld exceeds the last run (%lld)", (long long)R, (long long)h->lastR);
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<int> st((size_t)R);
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(st.data(), h->status.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  for (int64_t r = 0; r < R; r++)
    if (st[(size_t)r] && st[(size_t)r] != CPIMH_ERR_NOT_MET) {     // a capped replicate is reported through `finished`, not as a failure
      if (st[(size_t)r] == CPIMH_ERR_TREE_OVERFLOW) set_error("replicate %lld: ancestor tree capacity exhausted; raise cfg.tree_capacity", (long long)r);
      else set_error("replicate %lld failed with status %d", (long long)r, st[(size_t)r]);
      return st[(size_t)r];
    }
  const size_t P = ((size_t)h->cfg.nobs + 1) * h->cfg.dim;
  std::vector<double> sums(2 * P + 2);
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(sums.data(), h->sums.p, sizeof(double) * sums.size(), cudaMemcpyDeviceToHost, s));
  if (o->hbar) CPIMH_CUDA_CHECK(cudaMemcpyAsync(o->hbar, h->hbar.p, sizeof(double) * (size_t)R * P, cudaMemcpyDeviceToHost, s));
  if (o->bc) CPIMH_CUDA_CHECK(cudaMemcpyAsync(o->bc, h->bc.p, sizeof(double) * (size_t)R * P, cudaMemcpyDeviceToHost, s));
  CPIMH_REQUIRE(h->all_particles || !(o->hbar_rb || o->sum_hbar_rb || o->sum_hbar_rb_sq), "the *_rb outputs need cpimh_chains_set_all_particles(h, 1) before the run");
  std::vector<double> sums_rb(2 * P + 2);
  if (h->all_particles) CPIMH_CUDA_CHECK(cudaMemcpyAsync(sums_rb.data(), h->sums_rb.p, sizeof(double) * sums_rb.size(), cudaMemcpyDeviceToHost, s));
  if (o->hbar_rb) CPIMH_CUDA_CHECK(cudaMemcpyAsync(o->hbar_rb, h->hbar_rb.p, sizeof(double) * (size_t)R * P, cudaMemcpyDeviceToHost, s));
  if (o->meetingtime) CPIMH_CUDA_CHECK(cudaMemcpyAsync(o->meetingtime, h->mt.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost, s));
  if (o->iteration) CPIMH_CUDA_CHECK(cudaMemcpyAsync(o->iteration, h->it.p, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  if (o->sum_hbar) memcpy(o->sum_hbar, sums.data(), sizeof(double) * P);
  if (o->sum_hbar_sq) memcpy(o->sum_hbar_sq, sums.data() + P, sizeof(double) * P);
  if (o->nfilters) *o->nfilters = (int64_t)sums[2 * P];
  if (o->nfinished) *o->nfinished = (int64_t)sums[2 * P + 1];
  if (o->finished) for (int64_t r = 0; r < R; r++) o->finished[r] = st[(size_t)r] == 0 ? 1 : 0;
  if (o->sum_hbar_rb) memcpy(o->sum_hbar_rb, sums_rb.data(), sizeof(double) * P);
  if (o->sum_hbar_rb_sq) memcpy(o->sum_hbar_rb_sq, sums_rb.data() + P, sizeof(double) * P);
  return CPIMH_OK;
}

int cpimh_chains_fetch_samples(cpimh_chains* h, int64_t r, void* stream, int max_rows, double* samples1, double* samples2,
                               double* log_pdf1, double* log_pdf2) {
  CPIMH_REQUIRE(h, "null handle");
  CPIMH_REQUIRE(h->record, "handle was created without record_samples");
  CPIMH_REQUIRE(r >= 0 && r < h->lastR, "replicate index out of range");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t P = (size_t)h->cfg.nobs + 1;
  int rows = max_rows < h->max_rec ? max_rows : h->max_rec;
  if (samples1) CPIMH_CUDA_CHECK(cudaMemcpyAsync(samples1, h->samples1.as<double>() + (size_t)r * (h->max_rec + 1) * P, sizeof(double) * (size_t)(rows + 1) * P, cudaMemcpyDeviceToHost, s));
  if (samples2) CPIMH_CUDA_CHECK(cudaMemcpyAsync(samples2, h->samples2.as<double>() + (size_t)r * h->max_rec * P, sizeof(double) * (size_t)rows * P, cudaMemcpyDeviceToHost, s));
  if (log_pdf1) CPIMH_CUDA_CHECK(cudaMemcpyAsync(log_pdf1, h->log_pdf1.as<double>() + (size_t)r * (h->max_rec + 1), sizeof(double) * (size_t)(rows + 1), cudaMemcpyDeviceToHost, s));
  if (log_pdf2) CPIMH_CUDA_CHECK(cudaMemcpyAsync(log_pdf2, h->log_pdf2.as<double>() + (size_t)r * h->max_rec, sizeof(double) * (size_t)rows, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  return CPIMH_OK;
}
int64_t cpimh_chains_launch_count(const cpimh_chains* h) { return h ? h->launches : 0; }

int cpimh_coupled_pimh(const cpimh_config* cfg, const double* obs_host, int64_t R, uint64_t seed, uint32_t first_replicate,
                       int k, int K, int max_iterations, cpimh_chain_out* out_host) {
  CPIMH_REQUIRE(cfg && obs_host && out_host, "null argument");
  cpimh_config c = *cfg;
  for (int attempt = 0; attempt < 4; attempt++) {
    cpimh_chains* h = nullptr;
    int rc = cpimh_chains_create(&c, R, 0, &h);
    if (rc) return rc;
    rc = cpimh_chains_set_observations(h, obs_host);
    if (!rc) rc = cpimh_chains_run(h, R, seed, first_replicate, k, K, max_iterations, nullptr, nullptr);
    if (!rc) rc = cpimh_chains_fetch(h, R, nullptr, out_host);
    int M = h->geo.M;
    cpimh_chains_destroy(h);
    if (rc != CPIMH_ERR_TREE_OVERFLOW) return rc;
    c.tree_capacity = 2 * M;
    c.store_paths = 3;
  }
  return CPIMH_ERR_TREE_OVERFLOW;
}

// ============================================================================ single model operations
static int model_op(const cpimh_config* cfg_in, int op, int time, const double* noise, size_t noise_rows, uint64_t seed, uint64_t filter_id,
                    double* x, const double* y, double* logw) {
  CPIMH_REQUIRE(cfg_in, "null config");
  cpimh_config c = *cfg_in;
  if (c.nobs < 1) c.nobs = 1;
  int rc = api_validate(&c);
  if (rc) return rc;
  const ModelEntry* e = find_entry(c.model, c.dim);
  const size_t N = c.nparticles, D = c.dim;
  std::vector<double> pre = api_precompute(c);
  DevBuf dtheta, dpre, dnoise, dx, dy, dlw, dst;
  auto cleanup = [&]() { dtheta.release(); dpre.release(); dnoise.release(); dx.release(); dy.release(); dlw.release(); dst.release(); };
  rc = dtheta.alloc(sizeof(double) * CPIMH_MAX_THETA);
  if (!rc) rc = dpre.alloc(sizeof(double) * (pre.size() + 1));
  if (!rc) rc = dx.alloc(sizeof(double) * N * D);
  if (!rc) rc = dy.alloc(sizeof(double) * (c.dim_y + 1));
  if (!rc) rc = dlw.alloc(sizeof(double) * N);
  if (!rc) rc = dst.alloc(sizeof(int));
  if (!rc && noise && noise_rows) rc = dnoise.alloc(sizeof(double) * noise_rows * N);
  if (rc) { cleanup(); return rc; }
  cudaMemcpy(dtheta.p, c.theta, sizeof(double) * CPIMH_MAX_THETA, cudaMemcpyHostToDevice);
  if (!pre.empty()) cudaMemcpy(dpre.p, pre.data(), sizeof(double) * pre.size(), cudaMemcpyHostToDevice);
  if (op != 0) cudaMemcpy(dx.p, x, sizeof(double) * N * D, cudaMemcpyHostToDevice);
  if (y) cudaMemcpy(dy.p, y, sizeof(double) * c.dim_y, cudaMemcpyHostToDevice);
  if (noise && noise_rows) cudaMemcpy(dnoise.p, noise, sizeof(double) * noise_rows * N, cudaMemcpyHostToDevice);
  cudaMemset(dst.p, 0, sizeof(int));
  ModelOpParams q;
  memset(&q, 0, sizeof(q));
  q.op = op; q.N = c.nparticles; q.T = c.nobs; q.time = time;
  q.theta = dtheta.as<double>(); q.pre = dpre.as<double>(); q.npre = (int)pre.size();
  q.seed = seed; q.filter_id = filter_id; q.noise = (noise && noise_rows) ? dnoise.as<double>() : nullptr;
  q.n_trans = (int)noise_rows; q.sk_M = c.sk_max_events; q.integrator = c.integrator; q.substeps = c.substeps;
  q.x = dx.as<double>(); q.y = dy.as<double>(); q.logw = dlw.as<double>(); q.status = dst.as<int>();
  rc = e->launch_model_op(q, nullptr);
  int st = 0;
  cudaError_t ce = cudaDeviceSynchronize();
  if (!rc && ce != cudaSuccess) { set_error("CUDA error in model op: %s", cudaGetErrorString(ce)); rc = CPIMH_ERR_CUDA; }
  if (!rc) {
    cudaMemcpy(&st, dst.p, sizeof(int), cudaMemcpyDeviceToHost);
    if (op == 2) cudaMemcpy(logw, dlw.p, sizeof(double) * N, cudaMemcpyDeviceToHost);
    else cudaMemcpy(x, dx.p, sizeof(double) * N * D, cudaMemcpyDeviceToHost);
    if (st) { set_error("model transition failed with status %d (injected noise exhausted?)", st); rc = st; }
  }
  cleanup();
  return rc;
}
int cpimh_model_rinit(const cpimh_config* cfg, const double* noise, uint64_t seed, uint64_t filter_id, double* x) {
  CPIMH_REQUIRE(cfg && x, "null argument");
  return model_op(cfg, 0, 0, noise, noise ? (size_t)cpimh_model_n_init(cfg) : 0, seed, filter_id, x, nullptr, nullptr);
}
int cpimh_model_rtransition(const cpimh_config* cfg, int time, const double* noise, uint64_t seed, uint64_t filter_id, double* x) {
  CPIMH_REQUIRE(cfg && x, "null argument");
  CPIMH_REQUIRE(time >= 1, "time is 1-based");
  return model_op(cfg, 1, time, noise, noise ? (size_t)cpimh_model_n_trans(cfg) : 0, seed, filter_id, x, nullptr, nullptr);
}
int cpimh_model_dmeasurement(const cpimh_config* cfg, const double* x, const double* y, double* logw) {
  CPIMH_REQUIRE(cfg && x && y && logw, "null argument");
  return model_op(cfg, 2, 1, nullptr, 0, 0, 0, const_cast<double*>(x), y, logw);
}

// ============================================================================ stand-alone tree (class Tree)
int cpimh_tree_destroy(cpimh_tree* t) {
  if (!t) return CPIMH_OK;
  DevBuf* all[] = {&t->a_star, &t->o_star, &t->x_star, &t->l_star, &t->l_new, &t->b, &t->o, &t->x, &t->a, &t->found, &t->out};
  for (auto b : all) b->release();
  delete t;
  return CPIMH_OK;
}
int cpimh_tree_reset(cpimh_tree* t) {                                          // src/tree.cpp:22-26
  CPIMH_REQUIRE(t, "null tree");
  t->nsteps = 0;
  CPIMH_CUDA_CHECK(cudaMemset(t->a_star.p, 0, sizeof(int) * (size_t)t->M));
  CPIMH_CUDA_CHECK(cudaMemset(t->o_star.p, 0, sizeof(int) * (size_t)t->M));
  return CPIMH_OK;
}
int cpimh_tree_create(int N, int M, int dimx, cpimh_tree** out) {              // :6-16
  CPIMH_REQUIRE(out && N >= 1 && dimx >= 1, "bad argument");
  cpimh_tree* t = new cpimh_tree();
  t->N = N; t->M = M < 3 * N ? 3 * N : M; t->dimx = dimx; t->nsteps = 0;
  int rc = t->a_star.alloc(sizeof(int) * (size_t)t->M);
  if (!rc) rc = t->o_star.alloc(sizeof(int) * (size_t)t->M);
  if (!rc) rc = t->x_star.alloc(sizeof(double) * (size_t)t->M * dimx);
  if (!rc) rc = t->l_star.alloc(sizeof(int) * (size_t)N);
  if (!rc) rc = t->l_new.alloc(sizeof(int) * (size_t)N);
  if (!rc) rc = t->b.alloc(sizeof(int) * (size_t)N);
  if (!rc) rc = t->o.alloc(sizeof(int) * (size_t)N);
  if (!rc) rc = t->a.alloc(sizeof(int) * (size_t)N);
  if (!rc) rc = t->x.alloc(sizeof(double) * (size_t)N * dimx);
  if (!rc) rc = t->found.alloc(sizeof(int));
  if (!rc) { cudaMemset(t->x_star.p, 0, sizeof(double) * (size_t)t->M * dimx); rc = cpimh_tree_reset(t); }
  if (rc) { cpimh_tree_destroy(t); return rc; }
  *out = t;
  return CPIMH_OK;
}
int cpimh_tree_init(cpimh_tree* t, const double* x0) {                         // :28-34
  CPIMH_REQUIRE(t && x0, "null argument");
  CPIMH_CUDA_CHECK(cudaMemcpy(t->x.p, x0, sizeof(double) * (size_t)t->N * t->dimx, cudaMemcpyHostToDevice));
  tree_init_kernel<<<(t->N + 127) / 128, 128>>>(t->N, t->dimx, t->x.as<double>(), t->a_star.as<int>(), t->x_star.as<double>(), t->l_star.as<int>());
  CPIMH_CUDA_CHECK(cudaGetLastError());
  CPIMH_CUDA_CHECK(cudaDeviceSynchronize());
  return CPIMH_OK;
}
static int tree_double_size(cpimh_tree* t) {                                   // :101-116
  const int old_M = t->M, new_M = 2 * old_M;
  DevBuf na, no, nx;
  int rc = na.alloc(sizeof(int) * (size_t)new_M);
  if (!rc) rc = no.alloc(sizeof(int) * (size_t)new_M);
  if (!rc) rc = nx.alloc(sizeof(double) * (size_t)new_M * t->dimx);
  if (rc) { na.release(); no.release(); nx.release(); return rc; }
  cudaMemset(na.p, 0, sizeof(int) * (size_t)new_M);
  cudaMemset(no.p, 0, sizeof(int) * (size_t)new_M);
  cudaMemset(nx.p, 0, sizeof(double) * (size_t)new_M * t->dimx);
  cudaMemcpy(na.p, t->a_star.p, sizeof(int) * (size_t)old_M, cudaMemcpyDeviceToDevice);
  cudaMemcpy(no.p, t->o_star.p, sizeof(int) * (size_t)old_M, cudaMemcpyDeviceToDevice);
  cudaMemcpy(nx.p, t->x_star.p, sizeof(double) * (size_t)old_M * t->dimx, cudaMemcpyDeviceToDevice);
  t->a_star.release(); t->o_star.release(); t->x_star.release();
  t->a_star = na; t->o_star = no; t->x_star = nx; t->M = new_M;
  CPIMH_CUDA_CHECK(cudaGetLastError());
  return CPIMH_OK;
}
int cpimh_tree_update(cpimh_tree* t, const double* x, const int32_t* a) {      // :88-99
  CPIMH_REQUIRE(t && x && a, "null argument");
  const int N = t->N;
  for (int i = 0; i < N; i++) CPIMH_REQUIRE(a[i] >= 0 && a[i] < N, "ancestor index %d out of range at position %d (0-based expected)", a[i], i);
  CPIMH_CUDA_CHECK(cudaMemcpy(t->x.p, x, sizeof(double) * (size_t)N * t->dimx, cudaMemcpyHostToDevice));
  CPIMH_CUDA_CHECK(cudaMemcpy(t->a.p, a, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice));
  tree_prune_kernel<<<1, 1024>>>(N, t->a.as<int>(), t->o.as<int>(), t->a_star.as<int>(), t->o_star.as<int>(), t->l_star.as<int>(), t->b.as<int>());
  t->nsteps++;                                                                 // :37
  tree_find_slots_kernel<<<1, 1024>>>(N, t->M, t->o_star.as<int>(), t->l_new.as<int>(), t->found.as<int>());
  int found = 0;
  CPIMH_CUDA_CHECK(cudaMemcpy(&found, t->found.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (found < N) {
    const int old_M = t->M;
    int rc = tree_double_size(t);
    if (rc) return rc;
    tree_fill_overflow_kernel<<<(N - found + 127) / 128, 128>>>(N, found, old_M, t->l_new.as<int>());
  }
  tree_scatter_kernel<<<(N + 127) / 128, 128>>>(N, t->dimx, t->x.as<double>(), t->b.as<int>(), t->l_new.as<int>(), t->a_star.as<int>(),
                                                t->x_star.as<double>(), t->l_star.as<int>());
  CPIMH_CUDA_CHECK(cudaGetLastError());
  CPIMH_CUDA_CHECK(cudaDeviceSynchronize());
  return CPIMH_OK;
}
int cpimh_tree_get_path(cpimh_tree* t, int n, double* path) {                  // :118-129
  CPIMH_REQUIRE(t && path, "null argument");
  CPIMH_REQUIRE(n >= 0 && n < t->N, "particle index %d out of range (0-based)", n);
  const size_t sz = (size_t)t->dimx * (t->nsteps + 1);
  int rc = t->out.alloc(sizeof(double) * sz);
  if (rc) return rc;
  tree_get_path_kernel<<<1, 32>>>(t->dimx, t->nsteps, n, t->a_star.as<int>(), t->x_star.as<double>(), t->l_star.as<int>(), t->out.as<double>());
  CPIMH_CUDA_CHECK(cudaGetLastError());
  CPIMH_CUDA_CHECK(cudaMemcpy(path, t->out.p, sizeof(double) * sz, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}
int cpimh_tree_retrieve_xgeneration(cpimh_tree* t, int lag, double* xgen) {    // :131-141
  CPIMH_REQUIRE(t && xgen, "null argument");
  CPIMH_REQUIRE(lag >= 0 && lag <= t->nsteps, "lag %d outside [0, nsteps=%d]", lag, t->nsteps);
  const size_t sz = (size_t)t->dimx * t->N;
  int rc = t->out.alloc(sizeof(double) * sz);
  if (rc) return rc;
  tree_xgeneration_kernel<<<(t->N + 127) / 128, 128>>>(t->N, t->dimx, lag, t->a_star.as<int>(), t->x_star.as<double>(), t->l_star.as<int>(), t->out.as<double>());
  CPIMH_CUDA_CHECK(cudaGetLastError());
  CPIMH_CUDA_CHECK(cudaMemcpy(xgen, t->out.p, sizeof(double) * sz, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}
int cpimh_tree_fields(cpimh_tree* t, int* N, int* M, int* dimx, int* nsteps) {
  CPIMH_REQUIRE(t, "null tree");
  if (N) *N = t->N; if (M) *M = t->M; if (dimx) *dimx = t->dimx; if (nsteps) *nsteps = t->nsteps;
  return CPIMH_OK;
}
int cpimh_tree_arrays(cpimh_tree* t, int32_t* a_star, int32_t* o_star, int32_t* l_star, double* x_star) {
  CPIMH_REQUIRE(t, "null tree");
  if (a_star) CPIMH_CUDA_CHECK(cudaMemcpy(a_star, t->a_star.p, sizeof(int) * (size_t)t->M, cudaMemcpyDeviceToHost));
  if (o_star) CPIMH_CUDA_CHECK(cudaMemcpy(o_star, t->o_star.p, sizeof(int) * (size_t)t->M, cudaMemcpyDeviceToHost));
  if (l_star) CPIMH_CUDA_CHECK(cudaMemcpy(l_star, t->l_star.p, sizeof(int) * (size_t)t->N, cudaMemcpyDeviceToHost));
  if (x_star) CPIMH_CUDA_CHECK(cudaMemcpy(x_star, t->x_star.p, sizeof(double) * (size_t)t->M * t->dimx, cudaMemcpyDeviceToHost));
  return CPIMH_OK;
}

}  // extern "C"

// ============================================================================ all-particle outputs of the filter handle
namespace cpimh {
int api_all_particles(const AllParticlesArgs& a, cudaStream_t s, double* h_exp_path, double* h_paths, double* h_normw) {
  DevBuf dexp, dpaths, dcur;
  const size_t psz = (size_t)a.D * (a.T + 1);
  int rc = dcur.alloc(sizeof(int) * (size_t)a.N);
  if (!rc && h_exp_path) rc = dexp.alloc(sizeof(double) * psz);
  if (!rc && h_paths) rc = dpaths.alloc(sizeof(double) * psz * a.N);
  if (!rc && (h_exp_path || h_paths)) {
    all_paths_kernel<<<1, 512, 0, s>>>(a, h_exp_path ? dexp.as<double>() : nullptr, h_paths ? dpaths.as<double>() : nullptr, dcur.as<int>());
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && h_exp_path) e = cudaMemcpyAsync(h_exp_path, dexp.p, sizeof(double) * psz, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && h_paths) e = cudaMemcpyAsync(h_paths, dpaths.p, sizeof(double) * psz * a.N, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) { set_error("CUDA error in all_particles: %s", cudaGetErrorString(e)); rc = CPIMH_ERR_CUDA; }
  }
  if (!rc && h_normw) {
    cudaError_t e = cudaMemcpyAsync(h_normw, a.normw, sizeof(double) * (size_t)a.N, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) { set_error("CUDA error: %s", cudaGetErrorString(e)); rc = CPIMH_ERR_CUDA; }
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (!rc && e != cudaSuccess) { set_error("CUDA error: %s", cudaGetErrorString(e)); rc = CPIMH_ERR_CUDA; }
  dexp.release(); dpaths.release(); dcur.release();
  return rc;
}
}  // namespace cpimh
