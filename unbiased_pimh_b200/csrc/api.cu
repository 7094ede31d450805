// api.cu -- C ABI (include/cpimh.h) for the filter and the coupled-chain driver: configuration checks, HBM
// workspaces, launches.  No CPU fallback: every compute entry point needs a CUDA device.
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "host_common.h"
#include "wide.h"

namespace cpimh {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const ModelEntry* find_entry(int model, int dim) {
  const ModelEntry* (*tables[])(int*) = {entries_gss, entries_ar1, entries_ar2, entries_ar3, entries_ar4, entries_ar5, entries_ar6, entries_ar7, entries_ar8, entries_ar9, entries_ar10, entries_ar11, entries_ar12, entries_ar13, entries_ar14, entries_ar15, entries_ar16,
                                          entries_sv, entries_sk, entries_lorenz4, entries_lorenz8, entries_lorenz16};
  for (auto tf : tables) {
    int n = 0;
    const ModelEntry* e = tf(&n);
    for (int i = 0; i < n; i++)
      if (e[i].model == model && e[i].dim == dim) return &e[i];
  }
  return nullptr;
}

int api_n_init(const cpimh_config& c) {
  switch (c.model) {
    case CPIMH_MODEL_GSS: return 1;
    case CPIMH_MODEL_AR: case CPIMH_MODEL_LORENZ: return c.dim;
    case CPIMH_MODEL_LEVY_SV: return 3;
    default: return 0;
  }
}
int api_n_trans(const cpimh_config& c) {
  switch (c.model) {
    case CPIMH_MODEL_GSS: return 1;
    case CPIMH_MODEL_AR: case CPIMH_MODEL_LORENZ: return c.dim;
    case CPIMH_MODEL_LEVY_SV: return 2;
    case CPIMH_MODEL_SK: return 2 * c.sk_max_events;
    default: return 0;
  }
}

// host-side model precompute: model$precompute(theta) (R/model_get_ar.R:41-44) and the density constant of
// dmvnorm_transpose_cholesky (src/mvnorm.cpp:120-121), evaluated the reference's way (a loop of log(L_ii)).
std::vector<double> api_precompute(const cpimh_config& c) {
  std::vector<double> pre;
  if (c.model == CPIMH_MODEL_AR) {
    const int d = c.dim;
    pre.resize((size_t)d * d + 1);
    for (int i = 0; i < d; i++)
      for (int j = 0; j < d; j++) pre[i + (size_t)j * d] = pow(c.theta[0], 1 + abs(i - j));   // src/create_A_.cpp:10
    double hl = 0;
    for (int i = 0; i < d; i++) hl += log(c.theta[1]);
    pre[(size_t)d * d] = -(hl) - (d * CPIMH_HALF_LOG_2PI_TRUNC);
    pre.push_back(1.0 / c.theta[1]);
  } else if (c.model == CPIMH_MODEL_LEVY_SV) {
    const double xi = c.theta[2], w2 = c.theta[3], lambda = c.theta[4];
    pre.push_back(exp(-lambda));                       // src/SVLevy_functions.cpp:66 exp(-lambda)
    pre.push_back(1 / lambda);                         // :67 (1/lambda)
    pre.push_back(w2 / xi);                            // rexp(k, xi/w2) = exp_rand() * (w2/xi) (:63)
    // CDF of rpois(lambda xi^2 / w2) (:61) by the sequential-search recurrence, until it stops increasing
    const double rate = lambda * xi * xi / w2;
    std::vector<double> F, invp;
    double p = exp(-rate), acc = p;
    F.push_back(acc); invp.push_back(1.0 / p);
    for (int k = 1; k < 1000; k++) {
      p = p * rate / k;
      const double nxt = acc + p;
      if (nxt == acc) break;
      acc = nxt;
      F.push_back(acc); invp.push_back(1.0 / p);
    }
    pre.push_back((double)F.size());
    pre.insert(pre.end(), F.begin(), F.end());
    for (int k = 0; k < 3; k++) pre.push_back(2.0);    // guard knots for the unrolled head of the table search (models.cuh poisson_tab)
    pre.insert(pre.end(), invp.begin(), invp.end());   // 1 / p_k: position of the Poisson uniform inside its cell -> first jump time
  } else if (c.model == CPIMH_MODEL_SK) {
    const double sd = sqrt(c.theta[0]);                 // R/model_stochastic_kinetic.R:151 dnorm(..., sd = sqrt(s2_obs), log = TRUE)
    pre.push_back(1.0 / sd);
    pre.push_back(log(sd));
  } else if (c.model == CPIMH_MODEL_LORENZ) {
    double hl = 0;
    for (int i = 0; i < c.dim; i++) hl += log(c.theta[2]);
    pre.push_back(-(hl) - (c.dim * CPIMH_HALF_LOG_2PI_TRUNC));
    pre.push_back(1.0 / c.theta[2]);
  }
  return pre;
}

int api_validate(cpimh_config* cp) {
  cpimh_config& c = *cp;
  CPIMH_REQUIRE(c.model >= CPIMH_MODEL_GSS && c.model <= CPIMH_MODEL_LORENZ, "unknown model id %d", c.model);
  if (c.model == CPIMH_MODEL_GSS) { if (c.dim == 0) c.dim = 1; if (c.dim_y == 0) c.dim_y = 1; }
  if (c.model == CPIMH_MODEL_LEVY_SV) { if (c.dim == 0) c.dim = 2; if (c.dim_y == 0) c.dim_y = 1; }
  if (c.model == CPIMH_MODEL_SK) { if (c.dim == 0) c.dim = 4; if (c.dim_y == 0) c.dim_y = 2; }
  if ((c.model == CPIMH_MODEL_AR || c.model == CPIMH_MODEL_LORENZ) && c.dim_y == 0) c.dim_y = c.dim;
  CPIMH_REQUIRE(c.nparticles >= 1, "nparticles must be >= 1 (got %d)", c.nparticles);
  CPIMH_REQUIRE(c.nobs >= 1 && c.nobs < (1 << 20), "nobs must be in [1, 2^20) (got %d)", c.nobs);
  CPIMH_REQUIRE(wide_supported(c) || find_entry(c.model, c.dim) != nullptr,
                "model %d with dimension %d has no compiled kernel (see csrc/model_*.cu for the instantiated dimensions)", c.model, c.dim);
  if (c.model == CPIMH_MODEL_SK) CPIMH_REQUIRE(c.dim_y == 2, "stochastic kinetic model has dim_y = 2");
  if (c.model == CPIMH_MODEL_AR || c.model == CPIMH_MODEL_LORENZ) CPIMH_REQUIRE(c.dim_y == c.dim, "dim_y must equal dim for this model");
  if (c.model == CPIMH_MODEL_LEVY_SV) {
    double rate = c.theta[4] * c.theta[2] * c.theta[2] / c.theta[3];
    CPIMH_REQUIRE(rate > 0 && rate <= 64.0, "Levy SV jump rate lambda*xi^2/omega^2 = %g outside (0, 64]", rate);
  }
  if (c.sk_max_events <= 0) c.sk_max_events = 64;
  if (c.substeps <= 0) c.substeps = 1;
  return CPIMH_OK;
}

// launch geometry shared by the filter and the chain kernels
static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
int api_plan_geometry(const cpimh_config& c, const ModelEntry* e, bool chains, int64_t max_items, int npre, Geometry* g) {
  int dev = 0;
  CPIMH_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CPIMH_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  CPIMH_REQUIRE(prop.major >= 10, "libcpimh is built for sm_100a (Blackwell); device %d is sm_%d%d", dev, prop.major, prop.minor);
  const int N = c.nparticles, T = c.nobs, D = c.dim;
  g->num_sms = prop.multiProcessorCount;
  g->Npad = (N + 31) & ~31;
  CPIMH_REQUIRE(N <= 32768, "the CTA-per-filter kernel supports N <= 32768 particles (got %d)", N);
  // CTA size: about 8 particles per thread, snapped to a compiled variant (128/256/512/1024 threads)
  int threads = c.threads_per_filter > 0 ? c.threads_per_filter : next_pow2((N + 7) / 8);
  // 2048 < N <= 4096: three 256-thread CTAs per SM (80 registers per thread, 16 particles per thread) beat two 512-thread
  // CTAs (64 registers, spills) by ~10 % on levy_sv; larger N keeps 8 per thread (token lists and event lists hold 16)
  if (c.threads_per_filter <= 0 && threads == 512) threads = 256;
  g->variant = pf_variant_for(threads);
  if (threads >= 128) threads = pf_variant_threads(g->variant);
  else threads = threads <= 32 ? 32 : 64;                      // small filters: one or two warps (no cross-warp barriers to pay for)
  CPIMH_REQUIRE((long long)threads * CPIMH_MAX_E >= N, "N=%d needs at least %d threads per filter", N, (N + CPIMH_MAX_E - 1) / CPIMH_MAX_E);
  g->threads = threads;
  const size_t max_smem = prop.sharedMemPerBlockOptin;
  g->hist_x_elems = (size_t)(T + 1) * D * g->Npad;
  g->hist_a_elems = (size_t)T * g->Npad;

  auto tree_capacity = [&]() -> long long {
    long long Mwant;
    if (c.tree_capacity > 0) Mwant = c.tree_capacity < 3LL * N ? 3LL * N : c.tree_capacity;          // src/tree.cpp:7-9
    else { Mwant = 10LL * N * D; if (Mwant < 3LL * N) Mwant = 3LL * N; if (Mwant > 16LL * N) Mwant = 16LL * N; }   // R/CPF.R:17, capped
    return (Mwant + 31) & ~31LL;
  };
  auto occupancy = [&](int mode, int M, size_t* smem, int* bps) -> int {
    *smem = pf_smem_bytes(g->Npad, M, npre, mode, c.shuffle);
    if (*smem > max_smem) { *bps = 0; return CPIMH_OK; }
    return chains ? e->occupancy_chains(g->variant, threads, *smem, bps) : e->occupancy_pf(g->variant, threads, *smem, bps);
  };
  // path storage: 0 none, 1 auto (dense when the histories of all resident CTAs fit in HBM, else tree), 2 dense, 3 tree
  int mode = c.store_paths == 0 ? PATH_NONE : (c.store_paths == 3 ? PATH_TREE : PATH_DENSE);
  int M = 32, bps = 0, rc;
  if (mode != PATH_TREE) {
    rc = occupancy(mode, M, &g->smem, &bps);
    if (rc) return rc;
    if (mode == PATH_DENSE) {
      size_t free_b = 0, total_b = 0;
      CPIMH_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
      long long grid = (long long)bps * g->num_sms;
      if (max_items < grid) grid = max_items;
      const double need = (double)grid * ((double)g->hist_x_elems * 8.0 + (double)g->hist_a_elems * 4.0);
      if (need > 0.85 * (double)free_b) {
        CPIMH_REQUIRE(c.store_paths != 2, "dense path history needs %.1f GB but only %.1f GB of HBM is free; use store_paths = 3 (tree)", need / 1e9, free_b / 1e9);
        mode = PATH_TREE;
      }
    }
  }
  if (mode == PATH_TREE) {
    M = (int)tree_capacity();
    rc = occupancy(mode, M, &g->smem, &bps);
    if (rc) return rc;
    CPIMH_REQUIRE(g->smem <= max_smem, "filter with N=%d, tree capacity %d needs %zu bytes of shared memory per CTA (limit %zu); lower tree_capacity or N",
                  N, M, g->smem, max_smem);
  }
  CPIMH_REQUIRE(g->smem <= max_smem, "filter with N=%d needs %zu bytes of shared memory per CTA (limit %zu)", N, g->smem, max_smem);
  CPIMH_REQUIRE(bps >= 1, "kernel cannot be resident with %d threads and %zu bytes of shared memory", threads, g->smem);
  g->M = M; g->path_mode = mode; g->blocks_per_sm = bps;
  return CPIMH_OK;
}

// Tail launch: when the last wave of a batch would leave the SMs with one CTA instead of two or three, those filters run faster
// as ONE wide CTA per SM.  Measured on B200 (tools/tail_geometry.py: levy_sv, N = 4096, 136 filters on 148 SMs, T = 300):
// 128 / 256 / 512 / 1024 threads per filter 20.9 / 8.7 / 6.6 / 7.0 ms -- 512 threads (128 registers per thread, half the
// barrier width) beat 1024 up to N = 4096; above that the 1024-thread CTA wins on particles per thread.
// Returns tail->threads = 0 when no such variant applies.
int api_plan_tail_geometry(const cpimh_config& c, const ModelEntry* e, int npre, const Geometry& main, Geometry* tail) {
  *tail = main;
  tail->threads = 0;
  if (main.variant == PF_VARIANT_1024x1 || c.threads_per_filter > 0 || main.blocks_per_sm < 2 || c.nparticles < 2048) return CPIMH_OK;
  const int candidates[2] = {(c.nparticles <= 4096 && main.variant < PF_VARIANT_512x2) ? PF_VARIANT_512x2 : PF_VARIANT_1024x1, PF_VARIANT_1024x1};
  for (int k = 0; k < 2; k++) {
    Geometry t = main;
    t.variant = candidates[k]; t.threads = pf_variant_threads(t.variant);
    t.smem = pf_smem_bytes(t.Npad, t.M, npre, t.path_mode, c.shuffle);
    int bps = 0;
    int rc = e->occupancy_pf(t.variant, t.threads, t.smem, &bps);
    if (rc) return rc;
    if (bps < 1) continue;
    t.blocks_per_sm = 1;
    *tail = t;
    return CPIMH_OK;
  }
  return CPIMH_OK;
}

}  // namespace cpimh

using namespace cpimh;

// ============================================================================ filter handle
struct cpimh_pf {
  WideFilter* wide = nullptr;   // d = 32/64 Lorenz: one filter across the GPU (wide.cu); all other fields unused then
  cpimh_config cfg;
  const ModelEntry* entry;
  Geometry geo, geo_tail;
  int64_t maxB, lastB;
  int record_anc, all_particles;
  int64_t launches, bytes;
  int grid;
  DevBuf hist_x, hist_a, ascr, wscr, exact, jq;
  DevBuf obs, theta, pre, xbuf, a_star, o_star, x_star, logz, logz_t, traj, path_index, status, normw, leaf_out, anc_rec, exp_path;
  DevBuf inj_init, inj_trans, inj_resample, inj_shuffle, inj_path;
  bool have_inj[5];
  int npre;
};

extern "C" {

const char* cpimh_last_error(void) { return g_err; }
int cpimh_version(void) { return CPIMH_VERSION_MAJOR * 100 + CPIMH_VERSION_MINOR; }
void cpimh_config_default(cpimh_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->store_paths = 1; cfg->substeps = 1; cfg->sk_max_events = 64;
}
int cpimh_model_n_init(const cpimh_config* cfg) { return api_n_init(*cfg); }
int cpimh_model_n_trans(const cpimh_config* cfg) { cpimh_config c = *cfg; if (c.sk_max_events <= 0) c.sk_max_events = 64; return api_n_trans(c); }
int cpimh_device_count(int* count) { CPIMH_CUDA_CHECK(cudaGetDeviceCount(count)); return CPIMH_OK; }
int cpimh_set_device(int device) { CPIMH_CUDA_CHECK(cudaSetDevice(device)); return CPIMH_OK; }

int cpimh_pf_create(const cpimh_config* cfg_in, int64_t max_filters, cpimh_pf** out) {
  CPIMH_REQUIRE(cfg_in && out, "null argument");
  CPIMH_REQUIRE(max_filters >= 1, "max_filters must be >= 1");
  cpimh_config cfg = *cfg_in;
  int rc = api_validate(&cfg);
  if (rc) return rc;
  if (wide_supported(cfg)) {
    WideFilter* w = nullptr;
    rc = wide_create(cfg, max_filters, &w);
    if (rc) return rc;
    cpimh_pf* hw = new cpimh_pf();
    hw->wide = w; hw->cfg = cfg; hw->maxB = max_filters; hw->lastB = 0;
    *out = hw;
    return CPIMH_OK;
  }
  const ModelEntry* e = find_entry(cfg.model, cfg.dim);
  std::vector<double> pre = api_precompute(cfg);
  Geometry g;
  rc = api_plan_geometry(cfg, e, false, max_filters, (int)pre.size(), &g);
  if (rc) return rc;
  cpimh_pf* h = new cpimh_pf();
  h->cfg = cfg; h->entry = e; h->geo = g; h->maxB = max_filters;
  rc = api_plan_tail_geometry(cfg, e, (int)pre.size(), g, &h->geo_tail);
  if (rc) { delete h; return rc; } h->lastB = 0; h->record_anc = 0; h->all_particles = 0; h->launches = 0; h->bytes = 0;
  memset(h->have_inj, 0, sizeof(h->have_inj));
  const int D = cfg.dim, N = cfg.nparticles, T = cfg.nobs;
  h->npre = (int)pre.size();
  const size_t B = (size_t)max_filters;
  auto A = [&](DevBuf& b, size_t n) { return b.alloc(n, &h->bytes); };
  rc = A(h->obs, sizeof(double) * (size_t)T * cfg.dim_y);
  if (!rc) rc = A(h->theta, sizeof(double) * CPIMH_MAX_THETA);
  if (!rc) rc = A(h->pre, sizeof(double) * (pre.size() + 1));
  const int64_t resident = (int64_t)g.blocks_per_sm * g.num_sms;
  h->grid = (int)(max_filters < resident ? max_filters : resident);
  const size_t G = (size_t)h->grid;          // HBM workspaces belong to resident CTAs, not to filters
  if (!rc && g.path_mode != PATH_DENSE) rc = A(h->xbuf, sizeof(double) * G * 2 * D * g.Npad);
  if (!rc && g.path_mode != PATH_DENSE) rc = A(h->ascr, sizeof(int) * G * g.Npad);
  if (!rc) rc = A(h->wscr, sizeof(double) * G * g.Npad);
  if (!rc && cfg.model == CPIMH_MODEL_LEVY_SV) rc = A(h->jq, sizeof(uint4) * G * pf_jq_stride(g.Npad));   // per-warp jump queues
  if (!rc) rc = A(h->exact, sizeof(int) * B);
  if (!rc && g.path_mode == PATH_DENSE) {
    rc = A(h->hist_x, sizeof(double) * G * g.hist_x_elems);
    if (!rc) rc = A(h->hist_a, sizeof(int) * G * g.hist_a_elems);
  }
  if (!rc && g.path_mode == PATH_TREE) {
    rc = A(h->a_star, sizeof(int) * G * g.M);
    if (!rc) rc = A(h->o_star, sizeof(int) * G * g.M);
    if (!rc) rc = A(h->x_star, sizeof(double) * G * D * g.M);
    if (!rc) rc = A(h->leaf_out, sizeof(int) * G * N);
  }
  if (!rc && g.path_mode != PATH_NONE) rc = A(h->traj, sizeof(double) * B * (T + 1) * D);
  if (!rc) rc = A(h->logz, sizeof(double) * B);
  if (!rc) rc = A(h->logz_t, sizeof(double) * B * T);
  if (!rc) rc = A(h->path_index, sizeof(int) * B);
  if (!rc) rc = A(h->status, sizeof(int) * B);
  if (!rc) rc = A(h->normw, sizeof(double) * B * N);
  if (rc) { cpimh_pf_destroy(h); return rc; }
  cudaError_t ce = cudaMemcpy(h->theta.p, cfg.theta, sizeof(double) * CPIMH_MAX_THETA, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess && !pre.empty()) ce = cudaMemcpy(h->pre.p, pre.data(), sizeof(double) * pre.size(), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { set_error("cudaMemcpy failed: %s", cudaGetErrorString(ce)); cpimh_pf_destroy(h); return CPIMH_ERR_CUDA; }
  *out = h;
  return CPIMH_OK;
}

int cpimh_pf_destroy(cpimh_pf* h) {
  if (!h) return CPIMH_OK;
  if (h->wide) { wide_destroy(h->wide); delete h; return CPIMH_OK; }
  DevBuf* all[] = {&h->obs, &h->theta, &h->pre, &h->xbuf, &h->a_star, &h->o_star, &h->x_star, &h->logz, &h->logz_t, &h->traj, &h->path_index,
                   &h->status, &h->normw, &h->leaf_out, &h->anc_rec, &h->exp_path, &h->hist_x, &h->hist_a, &h->ascr, &h->wscr, &h->exact, &h->jq, &h->inj_init, &h->inj_trans, &h->inj_resample, &h->inj_shuffle, &h->inj_path};
  for (auto b : all) b->release();
  delete h;
  return CPIMH_OK;
}

int cpimh_pf_all_particles(cpimh_pf* h, int on) {
  CPIMH_REQUIRE(h, "null handle");
  if (h->wide) return wide_all_particles(h->wide, on);
  CPIMH_REQUIRE(!on || h->geo.path_mode != PATH_NONE, "all-particle outputs need store_paths != 0");
  h->all_particles = on ? 1 : 0;
  if (on) return h->exp_path.alloc(sizeof(double) * (size_t)h->maxB * (h->cfg.nobs + 1) * h->cfg.dim, &h->bytes);
  return CPIMH_OK;
}


}  // extern "C"
namespace cpimh {
int api_upload_obs(const cpimh_config& cfg, const double* obs_host, void* dst) {
  const int T = cfg.nobs, dy = cfg.dim_y;
  std::vector<double> rm((size_t)T * dy);
  for (int t = 0; t < T; t++)
    for (int j = 0; j < dy; j++) rm[(size_t)t * dy + j] = obs_host[t + (size_t)T * j];   // T x dy column-major -> row-major
  CPIMH_CUDA_CHECK(cudaMemcpy(dst, rm.data(), sizeof(double) * rm.size(), cudaMemcpyHostToDevice));
  return CPIMH_OK;
}
}  // namespace cpimh
extern "C" {

int cpimh_pf_set_observations(cpimh_pf* h, const double* obs_host) {
  CPIMH_REQUIRE(h && obs_host, "null argument");
  if (h->wide) return wide_set_observations(h->wide, obs_host);
  return api_upload_obs(h->cfg, obs_host, h->obs.p);
}

int cpimh_pf_set_noise(cpimh_pf* h, int64_t B, const cpimh_noise* nz) {
  CPIMH_REQUIRE(h, "null handle");
  if (h->wide) return wide_set_noise(h->wide, B, nz);
  memset(h->have_inj, 0, sizeof(h->have_inj));
  if (!nz) return CPIMH_OK;
  CPIMH_REQUIRE(B >= 1 && B <= h->maxB, "nfilters %lld outside [1, %lld]", (long long)B, (long long)h->maxB);
  const cpimh_config& c = h->cfg;
  const size_t N = c.nparticles, T = c.nobs;
  struct Item { const double* src; DevBuf* dst; size_t n; } items[5] = {
      {nz->init, &h->inj_init, (size_t)B * api_n_init(c) * N},
      {nz->trans, &h->inj_trans, (size_t)B * T * api_n_trans(c) * N},
      {nz->resample, &h->inj_resample, (size_t)B * T * N},
      {nz->shuffle, &h->inj_shuffle, (size_t)B * T * (N - 1)},
      {nz->path_u, &h->inj_path, (size_t)B}};
  for (int i = 0; i < 5; i++) {
    if (!items[i].src || items[i].n == 0) continue;
    int rc = items[i].dst->alloc(sizeof(double) * items[i].n, &h->bytes);
    if (rc) return rc;
    CPIMH_CUDA_CHECK(cudaMemcpy(items[i].dst->p, items[i].src, sizeof(double) * items[i].n, cudaMemcpyHostToDevice));
    h->have_inj[i] = true;
  }
  return CPIMH_OK;
}

int cpimh_pf_record_ancestors(cpimh_pf* h, int on) {
  CPIMH_REQUIRE(h, "null handle");
  if (h->wide) return wide_record_ancestors(h->wide, on);
  h->record_anc = on ? 1 : 0;
  if (on) return h->anc_rec.alloc(sizeof(int) * (size_t)h->maxB * h->cfg.nobs * h->cfg.nparticles, &h->bytes);
  return CPIMH_OK;
}

}  // extern "C"
namespace cpimh {
void api_fill_pf_params(const cpimh_config& c, const Geometry& g, int npre, PfParams* p) {
  memset(p, 0, sizeof(*p));
  p->N = c.nparticles; p->T = c.nobs; p->dy = c.dim_y; p->Npad = g.Npad; p->M = g.M;
  p->shuffle = c.shuffle; p->path_mode = g.path_mode;
  p->npre = npre; p->integrator = c.integrator; p->substeps = c.substeps;
  p->n_init = api_n_init(c); p->n_trans = api_n_trans(c); p->sk_M = c.sk_max_events;
  pf_layout(g.Npad, g.M, npre, g.path_mode, c.shuffle, &p->lay);
}
}  // namespace cpimh
extern "C" {

// launch parameters of one run of the handle's batch (filters are addressed by their position b in the batch: Philox stream
// first_filter + b, outputs at index b)
static void pf_run_params(cpimh_pf* h, uint64_t seed, uint64_t first_filter, PfParams* pp) {
  PfParams& p = *pp;
  api_fill_pf_params(h->cfg, h->geo, h->npre, &p);
  p.seed = seed; p.first_filter = first_filter;
  p.obs = h->obs.as<double>(); p.theta = h->theta.as<double>(); p.pre = h->pre.as<double>();
  p.hist_x = h->hist_x.as<double>(); p.hist_a = h->hist_a.as<int>(); p.ascr = h->ascr.as<int>();
  p.wscr = h->wscr.as<double>(); p.exact_repeats = h->exact.as<int>(); p.jq = h->jq.as<uint4>();
  p.xbuf = h->xbuf.as<double>(); p.a_star = h->a_star.as<int>(); p.o_star = h->o_star.as<int>(); p.x_star = h->x_star.as<double>();
  p.logz = h->logz.as<double>(); p.logz_t = h->logz_t.as<double>(); p.traj = h->traj.as<double>();
  p.path_index = h->path_index.as<int>(); p.status = h->status.as<int>(); p.normw = h->normw.as<double>();
  p.leaf_out = h->leaf_out.as<int>(); p.anc_rec = h->record_anc ? h->anc_rec.as<int>() : nullptr;
  p.exp_path = h->all_particles ? h->exp_path.as<double>() : nullptr;
  p.inj_init = h->have_inj[0] ? h->inj_init.as<double>() : nullptr;
  p.inj_trans = h->have_inj[1] ? h->inj_trans.as<double>() : nullptr;
  p.inj_resample = h->have_inj[2] ? h->inj_resample.as<double>() : nullptr;
  p.inj_shuffle = h->have_inj[3] ? h->inj_shuffle.as<double>() : nullptr;
  p.inj_path = h->have_inj[4] ? h->inj_path.as<double>() : nullptr;
}
// whole waves run with the main geometry; a short last wave (at most one filter per SM) as one wide CTA per SM
static void pf_split_tail(const cpimh_pf* h, int64_t B, int64_t* main_B, int64_t* tail_B) {
  *main_B = B; *tail_B = 0;
  if (h->geo_tail.threads > 0 && B > h->grid) {
    const int64_t rem = B % h->grid;
    if (rem > 0 && rem <= h->geo.num_sms) { *tail_B = rem; *main_B = B - rem; }
  }
}
// filters [b0, b0 + nb) of the batch as one launch (tail: the wide-CTA geometry)
static int pf_launch_range(cpimh_pf* h, PfParams p, int64_t b0, int64_t nb, bool tail, cudaStream_t stream) {
  const Geometry& g = tail ? h->geo_tail : h->geo;
  PfLaunch L;
  L.threads = g.threads; L.smem = g.smem; L.variant = g.variant;
  L.grid = (int)(tail ? nb : (nb < h->grid ? nb : h->grid));
  p.B = nb; p.b0 = b0;
  int rc = h->entry->launch_pf(p, L, stream);
  if (!rc) h->launches += 1;
  return rc;
}

int cpimh_pf_run(cpimh_pf* h, int64_t B, uint64_t seed, uint64_t first_filter, void* stream) {
  CPIMH_REQUIRE(h, "null handle");
  if (h->wide) { int rcw = wide_run(h->wide, B, seed, first_filter, (cudaStream_t)stream); if (!rcw) h->lastB = B; return rcw; }
  CPIMH_REQUIRE(B >= 1 && B <= h->maxB, "nfilters %lld outside [1, %lld]", (long long)B, (long long)h->maxB);
  PfParams p;
  pf_run_params(h, seed, first_filter, &p);
  int64_t main_B, tail_B;
  pf_split_tail(h, B, &main_B, &tail_B);
  int rc = pf_launch_range(h, p, 0, main_B, false, (cudaStream_t)stream);
  if (!rc && tail_B > 0) rc = pf_launch_range(h, p, main_B, tail_B, true, (cudaStream_t)stream);
  if (rc) return rc;
  h->lastB = B;
  return CPIMH_OK;
}

static int check_status(const int* d_status, int64_t B, cudaStream_t s) {
  std::vector<int> st((size_t)B);
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(st.data(), d_status, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  for (int64_t b = 0; b < B; b++)
    if (st[(size_t)b] != 0) {
      const int code = st[(size_t)b];
      if (code == CPIMH_ERR_TREE_OVERFLOW) set_error("filter %lld: ancestor tree capacity exhausted (Tree::double_size needed); raise cfg.tree_capacity", (long long)b);
      else if (code == CPIMH_ERR_NOISE_OVERFLOW) set_error("filter %lld: injected stochastic-kinetic noise exhausted; raise cfg.sk_max_events", (long long)b);
      else set_error("filter %lld failed with status %d", (long long)b, code);
      return code;
    }
  return CPIMH_OK;
}

int cpimh_pf_fetch(cpimh_pf* h, int64_t B, void* stream, double* logz, double* logz_t, double* trajectory, int32_t* path_index) {
  CPIMH_REQUIRE(h, "null handle");
  if (h->wide) return wide_fetch(h->wide, B, (cudaStream_t)stream, logz, logz_t, trajectory, path_index);
  CPIMH_REQUIRE(B >= 1 && B <= h->lastB, "nfilters %lld exceeds the last run (%lld)", (long long)B, (long long)h->lastB);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = check_status(h->status.as<int>(), B, s);
  if (rc) return rc;
  const cpimh_config& c = h->cfg;
  if (logz) CPIMH_CUDA_CHECK(cudaMemcpyAsync(logz, h->logz.p, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, s));
  if (logz_t) CPIMH_CUDA_CHECK(cudaMemcpyAsync(logz_t, h->logz_t.p, sizeof(double) * (size_t)B * c.nobs, cudaMemcpyDeviceToHost, s));
  if (trajectory) {
    CPIMH_REQUIRE(h->geo.path_mode != PATH_NONE, "trajectories need store_paths != 0");
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(trajectory, h->traj.p, sizeof(double) * (size_t)B * (c.nobs + 1) * c.dim, cudaMemcpyDeviceToHost, s));
  }
  if (path_index) CPIMH_CUDA_CHECK(cudaMemcpyAsync(path_index, h->path_index.p, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  return CPIMH_OK;
}

int cpimh_pf_fetch_ancestors(cpimh_pf* h, int64_t b, void* stream, int32_t* ancestors) {
  CPIMH_REQUIRE(h && ancestors, "null argument");
  if (h->wide) return wide_fetch_ancestors(h->wide, b, (cudaStream_t)stream, ancestors);
  CPIMH_REQUIRE(h->record_anc, "call cpimh_pf_record_ancestors(h, 1) before cpimh_pf_run");
  CPIMH_REQUIRE(b >= 0 && b < h->lastB, "filter index out of range");
  const size_t n = (size_t)h->cfg.nobs * h->cfg.nparticles;
  cudaStream_t s = (cudaStream_t)stream;
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(ancestors, h->anc_rec.as<int>() + (size_t)b * n, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  return CPIMH_OK;
}

int cpimh_pf_fetch_exp_paths(cpimh_pf* h, int64_t B, void* stream, double* exp_path) {
  CPIMH_REQUIRE(h && exp_path, "null argument");
  if (h->wide) return wide_fetch_exp_paths(h->wide, B, (cudaStream_t)stream, exp_path);
  CPIMH_REQUIRE(h->all_particles, "call cpimh_pf_all_particles(h, 1) before cpimh_pf_run");
  CPIMH_REQUIRE(B >= 0 && B <= h->lastB, "filter count out of range");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = check_status(h->status.as<int>(), B, s);
  if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(exp_path, h->exp_path.p, sizeof(double) * (size_t)B * (h->cfg.nobs + 1) * h->cfg.dim, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  return CPIMH_OK;
}

int cpimh_pf_fetch_all_particles(cpimh_pf* h, int64_t b, void* stream, double* exp_path, double* trajectories, double* normweights) {
  CPIMH_REQUIRE(h, "null handle");
  CPIMH_REQUIRE(!h->wide, "all-particle outputs are not available for the wide (d = 32/64) filter");
  CPIMH_REQUIRE(b >= 0 && b < h->lastB, "filter index out of range");
  CPIMH_REQUIRE(h->geo.path_mode != PATH_NONE, "all-particle outputs need store_paths != 0");
  CPIMH_REQUIRE(h->lastB <= h->grid, "all-particle outputs read the path store of the CTA that ran the filter: run at most %d filters per launch "
                "(the last run had %lld)", h->grid, (long long)h->lastB);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = check_status(h->status.as<int>() + b, 1, s);
  if (rc) return rc;
  const size_t N = h->cfg.nparticles, D = h->cfg.dim, M = h->geo.M;
  AllParticlesArgs a;
  a.N = (int)N; a.D = (int)D; a.M = (int)M; a.T = h->cfg.nobs; a.Npad = h->geo.Npad; a.path_mode = h->geo.path_mode;
  a.a_star = h->a_star.as<int>() ? h->a_star.as<int>() + (size_t)b * M : nullptr;
  a.x_star = h->x_star.as<double>() ? h->x_star.as<double>() + (size_t)b * D * M : nullptr;
  a.leaf = h->leaf_out.as<int>() ? h->leaf_out.as<int>() + (size_t)b * N : nullptr;
  a.hist_x = h->hist_x.as<double>() ? h->hist_x.as<double>() + (size_t)b * h->geo.hist_x_elems : nullptr;
  a.hist_a = h->hist_a.as<int>() ? h->hist_a.as<int>() + (size_t)b * h->geo.hist_a_elems : nullptr;
  a.normw = h->normw.as<double>() + (size_t)b * N;
  return api_all_particles(a, s, exp_path, trajectories, normweights);
}

int cpimh_pf_exact_mode(cpimh_pf* h, int on) {
  CPIMH_REQUIRE(h, "null handle");
  if (h->wide) wide_exact_mode(h->wide, on);        // the fused filter is always exact (in-kernel repeat of a flagged step)
  return CPIMH_OK;
}

int cpimh_pf_fetch_exact_repeats(cpimh_pf* h, int64_t B, void* stream, int32_t* repeats) {
  CPIMH_REQUIRE(h && repeats, "null argument");
  if (h->wide) return wide_fetch_exact_repeats(h->wide, B, (cudaStream_t)stream, repeats);
  CPIMH_REQUIRE(B >= 1 && B <= h->lastB, "nfilters %lld exceeds the last run (%lld)", (long long)B, (long long)h->lastB);
  cudaStream_t s = (cudaStream_t)stream;
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(repeats, h->exact.p, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  return CPIMH_OK;
}

int cpimh_pf_device_status(cpimh_pf* h, int32_t** d_status) {
  CPIMH_REQUIRE(h && d_status, "null argument");
  CPIMH_REQUIRE(!h->wide, "not available for the wide (d = 32/64) filter");
  *d_status = h->status.as<int>();
  return CPIMH_OK;
}

int cpimh_pf_device_exp_paths(cpimh_pf* h, double** d_exp_path) {
  CPIMH_REQUIRE(h && d_exp_path, "null argument");
  if (h->wide) { *d_exp_path = wide_device_exp_paths(h->wide); CPIMH_REQUIRE(*d_exp_path, "call cpimh_pf_all_particles(h, 1) first"); return CPIMH_OK; }
  CPIMH_REQUIRE(h->all_particles, "call cpimh_pf_all_particles(h, 1) first");
  *d_exp_path = h->exp_path.as<double>();
  return CPIMH_OK;
}

int cpimh_pf_device_outputs(cpimh_pf* h, double** d_logz, double** d_trajectory) {
  CPIMH_REQUIRE(h, "null handle");
  if (h->wide) { wide_device_outputs(h->wide, d_logz, d_trajectory); return CPIMH_OK; }
  if (d_logz) *d_logz = h->logz.as<double>();
  if (d_trajectory) *d_trajectory = h->traj.as<double>();
  return CPIMH_OK;
}
int64_t cpimh_pf_launch_count(const cpimh_pf* h) { return !h ? 0 : (h->wide ? wide_launch_count(h->wide) : h->launches); }
int64_t cpimh_pf_device_bytes(const cpimh_pf* h) { return !h ? 0 : (h->wide ? wide_device_bytes(h->wide) : h->bytes); }
int cpimh_pf_effective_config(const cpimh_pf* h, cpimh_config* out) {
  CPIMH_REQUIRE(h && out, "null argument");
  if (h->wide) { wide_effective(h->wide, out); return CPIMH_OK; }
  *out = h->cfg;
  out->tree_capacity = h->geo.path_mode == PATH_TREE ? h->geo.M : 0;
  out->threads_per_filter = h->geo.threads;
  out->store_paths = h->geo.path_mode == PATH_NONE ? 0 : (h->geo.path_mode == PATH_DENSE ? 2 : 3);
  out->reserved = h->geo.blocks_per_sm;
  return CPIMH_OK;
}

// The one-shot entry is what the R shim calls once per PIMH sweep (r-pkg/src/shim.c cpimh_R_pf_batch): its HBM workspaces
// (24 GB for BASELINE configs[1]) stay cached in a handle between calls with the same configuration instead of being
// allocated and freed every time.  cpimh_release_cached() frees them; so does any allocation that runs out of memory.
static std::recursive_mutex g_cache_mu;   // recursive: an allocation inside the cached call may ask for the cache to be dropped
static cpimh_pf* g_cached = nullptr;
static cpimh_config g_cached_cfg;
static int64_t g_cached_B = 0;

int cpimh_ccpf_release_cached(void);
int cpimh_release_cached(void) {
  {
    std::lock_guard<std::recursive_mutex> lk(g_cache_mu);
    if (g_cached) { cpimh_pf_destroy(g_cached); g_cached = nullptr; g_cached_B = 0; }
  }
  return cpimh_ccpf_release_cached();
}

// One-shot entry, several waves: the batch runs as up to four groups of whole waves (+ the tail launch) on one stream while the
// finished groups' results are copied to the caller's (pageable) buffers on a second one, so that only the last group's copy
// is exposed (ar(10), 16 384 filters: 132 MB of trajectories, 22 ms of a 140 ms call).
static cudaStream_t g_pipe_run = nullptr, g_pipe_copy = nullptr;
static int g_pipe_device = -1;
static int pf_run_fetch_pipelined(cpimh_pf* h, int64_t B, uint64_t seed, uint64_t first_filter,
                                  double* logz, double* logz_t, double* trajectory, int32_t* path_index) {
  CPIMH_REQUIRE(B >= 1 && B <= h->maxB, "nfilters %lld outside [1, %lld]", (long long)B, (long long)h->maxB);
  CPIMH_REQUIRE(!trajectory || h->geo.path_mode != PATH_NONE, "trajectories need store_paths != 0");
  int dev = 0;
  CPIMH_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev != g_pipe_device) {                                      // streams belong to the device that was current when they were made
    CPIMH_CUDA_CHECK(cudaStreamCreateWithFlags(&g_pipe_run, cudaStreamNonBlocking));
    CPIMH_CUDA_CHECK(cudaStreamCreateWithFlags(&g_pipe_copy, cudaStreamNonBlocking));
    g_pipe_device = dev;
  }
  const cpimh_config& c = h->cfg;
  PfParams p;
  pf_run_params(h, seed, first_filter, &p);
  int64_t main_B, tail_B;
  pf_split_tail(h, B, &main_B, &tail_B);
  struct Group { int64_t b0, nb; bool tail; cudaEvent_t done; };
  std::vector<Group> groups;
  const int64_t waves = (main_B + h->grid - 1) / h->grid;
  // a group boundary costs a partly idle wave end, so groups are only formed where there is something to hide: one per 16 MB
  // of results, at most four
  const size_t per_filter = sizeof(double) * ((logz ? 1 : 0) + (logz_t ? (size_t)c.nobs : 0) + (trajectory ? (size_t)(c.nobs + 1) * c.dim : 0)) + (path_index ? sizeof(int) : 0);
  int64_t ngroups = (int64_t)((per_filter * (size_t)main_B) >> 24);
  if (ngroups > 4) ngroups = 4;
  if (ngroups > waves) ngroups = waves;
  if (ngroups < 1) ngroups = 1;
  for (int64_t g = 0, w0 = 0; g < ngroups; g++) {
    const int64_t w1 = waves * (g + 1) / ngroups;
    const int64_t b0 = w0 * h->grid, b1 = (w1 * h->grid < main_B) ? w1 * h->grid : main_B;
    if (b1 > b0) groups.push_back({b0, b1 - b0, false, nullptr});
    w0 = w1;
  }
  if (tail_B > 0) groups.push_back({main_B, tail_B, true, nullptr});
  CPIMH_CUDA_CHECK(cudaDeviceSynchronize());                       // observations / noise uploads of the legacy stream are complete
  int rc = CPIMH_OK;
  for (auto& g : groups) {
    if (!rc) rc = pf_launch_range(h, p, g.b0, g.nb, g.tail, g_pipe_run);
    if (!rc && cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate failed"); rc = CPIMH_ERR_CUDA; }
    if (!rc && cudaEventRecord(g.done, g_pipe_run) != cudaSuccess) { set_error("cudaEventRecord failed"); rc = CPIMH_ERR_CUDA; }
  }
  const size_t T = (size_t)c.nobs, D = (size_t)c.dim;
  auto copy = [&](void* dst, const void* src, size_t bytes) {
    if (!rc && cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_pipe_copy) != cudaSuccess) { set_error("device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = CPIMH_ERR_CUDA; }
  };
  for (auto& g : groups) {
    if (rc) break;
    if (cudaStreamWaitEvent(g_pipe_copy, g.done, 0) != cudaSuccess) { set_error("cudaStreamWaitEvent failed"); rc = CPIMH_ERR_CUDA; break; }
    const size_t b0 = (size_t)g.b0, nb = (size_t)g.nb;
    if (logz) copy(logz + b0, h->logz.as<double>() + b0, sizeof(double) * nb);
    if (logz_t) copy(logz_t + b0 * T, h->logz_t.as<double>() + b0 * T, sizeof(double) * nb * T);
    if (trajectory) copy(trajectory + b0 * (T + 1) * D, h->traj.as<double>() + b0 * (T + 1) * D, sizeof(double) * nb * (T + 1) * D);
    if (path_index) copy(path_index + b0, h->path_index.as<int>() + b0, sizeof(int) * nb);
  }
  cudaStreamSynchronize(g_pipe_run);
  cudaStreamSynchronize(g_pipe_copy);
  for (auto& g : groups) if (g.done) cudaEventDestroy(g.done);
  if (rc) return rc;
  CPIMH_CUDA_CHECK(cudaGetLastError());
  h->lastB = B;
  return check_status(h->status.as<int>(), B, g_pipe_copy);      // a failed filter invalidates the call (its outputs were copied but are not returned as valid)
}

static int pf_batch_cached(const cpimh_config* cfg, const double* obs_host, int64_t B, uint64_t seed, uint64_t first_filter,
                           const cpimh_noise* noise_host, double* logz, double* logz_t, double* trajectory, int32_t* path_index, double* exp_path) {
  CPIMH_REQUIRE(cfg && obs_host, "null argument");
  CPIMH_REQUIRE(B >= 1, "nfilters must be >= 1");
  std::lock_guard<std::recursive_mutex> lk(g_cache_mu);
  cpimh_config c = *cfg;
  for (int attempt = 0; attempt < 4; attempt++) {
    int rc = CPIMH_OK;
    if (!(g_cached && g_cached_B >= B && memcmp(&g_cached_cfg, &c, sizeof(c)) == 0)) {
      if (g_cached) { cpimh_pf_destroy(g_cached); g_cached = nullptr; g_cached_B = 0; }
      rc = cpimh_pf_create(&c, B, &g_cached);
      if (rc) { g_cached = nullptr; return rc; }
      g_cached_cfg = c; g_cached_B = B;
    }
    cpimh_pf* h = g_cached;
    rc = cpimh_pf_set_observations(h, obs_host);
    if (!rc) rc = cpimh_pf_set_noise(h, B, noise_host);                 // NULL clears noise left by an earlier call
    if (!rc) rc = cpimh_pf_record_ancestors(h, 0);
    if (!rc) rc = cpimh_pf_all_particles(h, exp_path ? 1 : 0);
    if (!rc && !h->wide && !exp_path && B > h->grid) rc = pf_run_fetch_pipelined(h, B, seed, first_filter, logz, logz_t, trajectory, path_index);
    else {
      if (!rc) rc = cpimh_pf_run(h, B, seed, first_filter, nullptr);
      if (!rc) rc = cpimh_pf_fetch(h, B, nullptr, logz, logz_t, trajectory, path_index);
    }
    if (!rc && exp_path) rc = cpimh_pf_fetch_exp_paths(h, B, nullptr, exp_path);
    if (rc != CPIMH_ERR_TREE_OVERFLOW) {
      if (rc) { cpimh_pf_destroy(g_cached); g_cached = nullptr; g_cached_B = 0; }   // do not keep a handle in an unknown state
      return rc;
    }
    const int M = h->wide ? 0 : h->geo.M;
    cpimh_pf_destroy(g_cached); g_cached = nullptr; g_cached_B = 0;
    c.tree_capacity = 2 * M;
    c.store_paths = 3;          // Tree::double_size (src/tree.cpp:101-116): rerun with twice the capacity (same streams => same draws)
  }
  return CPIMH_ERR_TREE_OVERFLOW;
}

int cpimh_pf_batch(const cpimh_config* cfg, const double* obs_host, int64_t B, uint64_t seed, uint64_t first_filter,
                   const cpimh_noise* noise_host, double* logz, double* logz_t, double* trajectory, int32_t* path_index) {
  return pf_batch_cached(cfg, obs_host, B, seed, first_filter, noise_host, logz, logz_t, trajectory, path_index, nullptr);
}
int cpimh_pf_batch_all_particles(const cpimh_config* cfg, const double* obs_host, int64_t B, uint64_t seed, uint64_t first_filter,
                                 double* logz, double* trajectory, double* exp_path) {
  CPIMH_REQUIRE(exp_path, "null argument");
  return pf_batch_cached(cfg, obs_host, B, seed, first_filter, nullptr, logz, nullptr, trajectory, nullptr, exp_path);
}

}  // extern "C"
