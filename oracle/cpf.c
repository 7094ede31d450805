/* oracle/cpf.c -- restatement of the bootstrap filter loop CPF(ref_trajectory=NULL) (R/CPF.R:14-78),
 * particlefilter() (R/particlefilter.R:11-44), pf_get_weighted_average (R/CPF.R:84-92), the PIMH kernels
 * (R/pf_kernels.R:86-145), coupled_pimh_chains (R/debiasing_algorithm.R:140-265) and H_bar (:284-316).
 * TEST INFRASTRUCTURE (see oracle.h). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle.h"
#include "rngmath.h"

int orc_model_rtransition_ex(const orc_config* c, int time, const double* noise, uint64_t seed, uint64_t fid, double* x, int* clamp);

static inline uint32_t c1_of(uint64_t filter_id, int t) { return (uint32_t)t | ((uint32_t)(filter_id >> 32) << 20); }
/* bits 12..19 of the iteration (high word of the filter id) go to the top byte of counter word 3 (DESIGN.md "RNG") */
static inline uint32_t c3_of(uint64_t filter_id, uint32_t c3) { return c3 | ((((uint32_t)(filter_id >> 32)) << 12) & 0xff000000u); }

int orc_cpf(const orc_config* cfg, const double* obs, const orc_noise* noise, uint64_t seed, uint64_t fid, orc_pf_out* out) {
  const int N = cfg->nparticles, d = cfg->dim, T = cfg->nobs, dy = cfg->dim_y;
  const int n_init = orc_model_n_init(cfg), n_trans = orc_model_n_trans(cfg);
  int rc = 0, clamp = 0;
  orc_tree* tree = orc_tree_new(N, 10 * N * d, d);                       /* R/CPF.R:17 */
  double* x = (double*)malloc(sizeof(double) * (size_t)d * N);
  double* xg = (double*)malloc(sizeof(double) * (size_t)d * N);
  double* normw = (double*)malloc(sizeof(double) * (size_t)N);
  double* logw = (double*)malloc(sizeof(double) * (size_t)N);
  double* nbuf = (double*)malloc(sizeof(double) * (size_t)N * (size_t)((n_trans > n_init ? n_trans : n_init) + 1));
  double* ebuf = (double*)malloc(sizeof(double) * (size_t)N * 2);
  double* ubuf = ebuf + N;
  double* yrow = (double*)malloc(sizeof(double) * (size_t)(dy > 0 ? dy : 1));
  int* anc = (int*)malloc(sizeof(int) * (size_t)N);

  const double* ninit = noise && noise->init ? noise->init : NULL;       /* :19-21 */
  if (!ninit) { orc_fill_init_noise(cfg, seed, fid, nbuf); ninit = nbuf; }
  orc_model_rinit(cfg, ninit, x);
  orc_tree_init(tree, x);                                                /* :25 */
  for (int i = 0; i < N; i++) normw[i] = 1.0 / N;                        /* :28 */
  double logz_sum = 0;

  for (int time = 1; time <= T; time++) {                                /* :35 */
    /* :36 multinomial_resampling_n(normweights, N): always evaluated in the reference; its result is
     * discarded at time 1 / after an NA observation (:38-40), so the oracle skips the arithmetic there. */
    int identity = (time == 1) || isnan(obs[(time - 2)]);
    if (identity) {
      for (int i = 0; i < N; i++) anc[i] = i;
    } else {
      if (noise && noise->resample) {
        clamp += orc_multinomial_from_sorted(normw, N, N, noise->resample + (size_t)(time - 1) * N, 1, anc);
      } else {
        /* Philox mode (DESIGN.md "RNG"): the sorted uniforms are the order statistics built from exponential spacings,
         * e_i = (E_0+..+E_i) / (E_0+..+E_N), E_i = -log(u_i); same law as the recurrence of src/multinomial.cpp:20-24 */
        double acc = 0, ua, ub;
        for (int i = 0; i < N; i++) {
          orc_philox_uniform2(seed, (uint32_t)i, c1_of(fid, time), (uint32_t)fid, c3_of(fid, ORC_P_RESAMPLE), &ua, &ub);
          acc = (i == 0) ? -orc_rng_log(ua) : acc + (-orc_rng_log(ua));
          ebuf[i] = acc;
        }
        orc_philox_uniform2(seed, (uint32_t)N, c1_of(fid, time), (uint32_t)fid, c3_of(fid, ORC_P_RESAMPLE), &ua, &ub);
        clamp += orc_multinomial_from_ladder(normw, N, N, ebuf, acc - orc_rng_log(ua), anc);
      }
      if (cfg->shuffle) {
        const double* us;
        if (noise && noise->shuffle) us = noise->shuffle + (size_t)(time - 1) * (N - 1);
        else {
          for (int i = 1; i < N; i++) { double ub; orc_philox_uniform2(seed, (uint32_t)i, c1_of(fid, time), (uint32_t)fid, c3_of(fid, ORC_P_SHUFFLE), &ubuf[i - 1], &ub); }
          us = ubuf;
        }
        orc_random_shuffle_int(anc, N, us);
      }
    }
    for (int i = 0; i < N; i++) memcpy(xg + (size_t)i * d, x + (size_t)anc[i] * d, sizeof(double) * d);   /* :41 */
    { double* t = x; x = xg; xg = t; }
    const double* ntr = NULL;                                            /* :43-44 */
    if (noise && noise->trans) ntr = noise->trans + (size_t)(time - 1) * n_trans * N;
    else if (cfg->model != ORC_MODEL_SK) { orc_fill_trans_noise(cfg, seed, fid, time, nbuf); ntr = nbuf; }
    rc = orc_model_rtransition_ex(cfg, time, ntr, seed, fid, x, &clamp);
    if (rc) goto done;
    for (int j = 0; j < dy; j++) yrow[j] = obs[(time - 1) + (size_t)T * j];
    orc_model_dmeasurement(cfg, x, yrow, logw);                          /* :60 */
    double lz = orc_logz_increment(logw, N, normw);                      /* :61-63,66 */
    orc_tree_update(tree, x, anc);                                       /* :65 */
    if (out->logz_t) out->logz_t[time - 1] = lz;
    if (out->ancestors_t) memcpy(out->ancestors_t + (size_t)(time - 1) * N, anc, sizeof(int) * (size_t)N);
    logz_sum += lz;                                                      /* :71 sum(logz): sequential here; R's sum is long double */
  }
  {
    long double s = 0;                                                   /* R's sum(logz) accumulates in long double */
    if (out->logz_t) { for (int t = 0; t < T; t++) s += out->logz_t[t]; logz_sum = (double)s; }
    double u;
    if (noise && noise->path_u) u = noise->path_u[0];
    else { double ub; orc_philox_uniform2(seed, 0, c1_of(fid, T + 1), (uint32_t)fid, c3_of(fid, ORC_P_PATH), &u, &ub); }
    int k;
    clamp += orc_systematic_resampling_n(normw, N, 1, u, &k);            /* :69 */
    orc_tree_get_path(tree, k, out->trajectory);
    if (out->path_index) *out->path_index = k;
    if (out->logz) *out->logz = logz_sum;
    if (out->normweights) memcpy(out->normweights, normw, sizeof(double) * (size_t)N);
    if (out->exp_path || out->all_paths) {                               /* :73-74, R/particlefilter.R:39-42 */
      size_t psz = (size_t)d * (T + 1);
      double* paths = out->all_paths ? out->all_paths : (double*)malloc(sizeof(double) * psz * N);
      for (int n = 0; n < N; n++) orc_tree_get_path(tree, n, paths + psz * n);
      if (out->exp_path)                                                 /* pf_get_weighted_average: colSums in long double */
        for (size_t e = 0; e < psz; e++) {
          long double acc = 0;
          for (int n = 0; n < N; n++) acc += paths[e + psz * n] * normw[n];
          out->exp_path[e] = (double)acc;
        }
      if (!out->all_paths) free(paths);
    }
    if (out->tree_M_final) *out->tree_M_final = orc_tree_M(tree);
    if (out->clamp_count) *out->clamp_count = clamp;
  }
done:
  orc_tree_free(tree);
  free(x); free(xg); free(normw); free(logw); free(nbuf); free(ebuf); free(yrow); free(anc);
  return rc;
}

int orc_cpf_batch(const orc_config* cfg, const double* obs, uint64_t seed, uint64_t first_filter, int nfilters,
                  int nthreads, double* logz, double* traj) {
  int rc = 0;
  size_t psz = (size_t)cfg->dim * (cfg->nobs + 1);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < nfilters; b++) {
    double* tr = traj ? traj + psz * b : (double*)malloc(sizeof(double) * psz);
    double* lzt = (double*)malloc(sizeof(double) * (size_t)cfg->nobs);
    orc_pf_out o; memset(&o, 0, sizeof(o));
    o.trajectory = tr; o.logz = &logz[b]; o.logz_t = lzt;
    int r = orc_cpf(cfg, obs, NULL, seed, first_filter + (uint64_t)b, &o);
    if (r) {
#pragma omp atomic write
      rc = r;
    }
    free(lzt);
    if (!traj) free(tr);
  }
  return rc;
}

/* ---- coupled PIMH chains ---- */
static int run_filter(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t replicate, uint32_t iter_index,
                      double* traj, double* logz, double* exp_path) {
  orc_pf_out o; memset(&o, 0, sizeof(o));
  double* lzt = (double*)malloc(sizeof(double) * (size_t)cfg->nobs);
  o.trajectory = traj; o.logz = logz; o.logz_t = lzt; o.exp_path = exp_path;
  int rc = orc_cpf(cfg, obs, NULL, seed, (uint64_t)replicate | ((uint64_t)iter_index << 32), &o);
  free(lzt);
  return rc;
}

int orc_coupled_pimh_chains(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t replicate,
                            int K, int max_iterations, int run_discarded_init, orc_chains* out) {
  return orc_coupled_pimh_chains_ap(cfg, obs, seed, replicate, K, max_iterations, run_discarded_init, 0, out);
}

/* all_particles != 0: the *_all_particles kernels (R/pf_kernels.R:147-215): exp_path is accepted together with the state */
int orc_coupled_pimh_chains_ap(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t replicate,
                               int K, int max_iterations, int run_discarded_init, int all_particles, orc_chains* out) {
  const int d = cfg->dim, p = cfg->nobs + 1;
  size_t psz = (size_t)d * p;
  double *s1 = (double*)malloc(sizeof(double) * psz), *s2 = (double*)malloc(sizeof(double) * psz), *prop = (double*)malloc(sizeof(double) * psz);
  double *e1 = NULL, *e2 = NULL, *eprop = NULL;
  if (all_particles) { e1 = (double*)calloc(psz, sizeof(double)); e2 = (double*)calloc(psz, sizeof(double)); eprop = (double*)calloc(psz, sizeof(double)); }
  int have2 = 0, cap = K + 10 + 1, nf = 0;
  double lz1, lz2 = -INFINITY, lzp;
  memset(out, 0, sizeof(*out));
  out->p = p;
  out->samples1 = (double*)malloc(sizeof(double) * (size_t)cap * p);
  out->samples2 = (double*)malloc(sizeof(double) * (size_t)cap * p);
  out->log_pdf1 = (double*)malloc(sizeof(double) * (size_t)cap);
  out->log_pdf2 = (double*)malloc(sizeof(double) * (size_t)cap);
  if (all_particles) {                                                   /* debiasing_algorithm.R:164-171 */
    out->exp_samples1 = (double*)malloc(sizeof(double) * (size_t)cap * p);
    out->exp_samples2 = (double*)malloc(sizeof(double) * (size_t)cap * p);
  }
  int rc = run_filter(cfg, obs, seed, replicate, 0, s1, &lz1, e1); nf++;     /* pimh_init, R/pf_kernels.R:134-139 (:200-206) */
  if (rc) return rc;
  if (run_discarded_init) { double lzd; run_filter(cfg, obs, seed, replicate, ORC_MAX_ITERATION, prop, &lzd, NULL); nf++; }  /* :140-142, discarded at debiasing_algorithm.R:160-161 */
  for (int t = 0; t < p; t++) out->samples1[t] = s1[(size_t)t * d];      /* :156 chain_state1[1,] */
  if (all_particles) for (int t = 0; t < p; t++) out->exp_samples1[t] = e1[(size_t)t * d];   /* :169 */
  out->log_pdf1[0] = lz1;
  int iter = 0, meet = 0, finished = 0, meetingtime = -1;
  while (!finished && (max_iterations < 0 || iter < max_iterations)) {   /* :180 */
    iter++;
    rc = run_filter(cfg, obs, seed, replicate, (uint32_t)iter, prop, &lzp, eprop); nf++;
    if (rc) return rc;
    double u, ub;
    orc_philox_uniform2(seed, 0, (uint32_t)iter << 20, replicate, c3_of((uint64_t)iter << 32, ORC_P_MH), &u, &ub);
    double lu = log(u);
    if (meet) {                                                          /* single_kernel R/pf_kernels.R:88-101 (:147-162) */
      if (lu < (lzp - lz1)) { memcpy(s1, prop, sizeof(double) * psz); lz1 = lzp; if (all_particles) memcpy(e1, eprop, sizeof(double) * psz); }
      memcpy(s2, s1, sizeof(double) * psz); lz2 = lz1;
      if (all_particles) memcpy(e2, e1, sizeof(double) * psz);
    } else {                                                             /* coupled_kernel :104-131 (:165-197) */
      if (lu < (lzp - lz1)) { memcpy(s1, prop, sizeof(double) * psz); lz1 = lzp; if (all_particles) memcpy(e1, eprop, sizeof(double) * psz); }
      if (lu < (lzp - lz2)) { memcpy(s2, prop, sizeof(double) * psz); lz2 = lzp; have2 = 1; if (all_particles) memcpy(e2, eprop, sizeof(double) * psz); }
      if (have2 && lz1 == lz2 && memcmp(s1, s2, sizeof(double) * psz) == 0) { meet = 1; meetingtime = iter; }  /* debiasing_algorithm.R:216-220 */
    }
    if (iter + 1 > cap) {                                                /* :222-233 */
      cap = 2 * cap;
      out->samples1 = (double*)realloc(out->samples1, sizeof(double) * (size_t)cap * p);
      out->samples2 = (double*)realloc(out->samples2, sizeof(double) * (size_t)cap * p);
      out->log_pdf1 = (double*)realloc(out->log_pdf1, sizeof(double) * (size_t)cap);
      out->log_pdf2 = (double*)realloc(out->log_pdf2, sizeof(double) * (size_t)cap);
      if (all_particles) {
        out->exp_samples1 = (double*)realloc(out->exp_samples1, sizeof(double) * (size_t)cap * p);
        out->exp_samples2 = (double*)realloc(out->exp_samples2, sizeof(double) * (size_t)cap * p);
      }
    }
    if (all_particles) for (int t = 0; t < p; t++) {                     /* :239-242 */
      out->exp_samples1[(size_t)iter * p + t] = e1[(size_t)t * d];
      out->exp_samples2[(size_t)(iter - 1) * p + t] = e2[(size_t)t * d];
    }
    for (int t = 0; t < p; t++) {                                        /* :235-236 */
      out->samples1[(size_t)iter * p + t] = s1[(size_t)t * d];
      out->samples2[(size_t)(iter - 1) * p + t] = s2[(size_t)t * d];
    }
    out->log_pdf1[iter] = lz1; out->log_pdf2[iter - 1] = lz2;
    int stop_at = meetingtime < 0 ? -1 : (meetingtime > K ? meetingtime : K);
    if (stop_at >= 0 && iter >= stop_at) finished = 1;                   /* :245 iter >= max(meetingtime, K) */
  }
  out->iteration = iter; out->meetingtime = meetingtime; out->finished = finished; out->nrows1 = iter + 1; out->nfilters = nf;
  free(s1); free(s2); free(prop); free(e1); free(e2); free(eprop);
  return 0;
}
void orc_chains_free(orc_chains* c) {
  free(c->samples1); free(c->samples2); free(c->log_pdf1); free(c->log_pdf2); free(c->exp_samples1); free(c->exp_samples2);
  memset(c, 0, sizeof(*c));
}

static int hbar_impl(const orc_chains* c, const double* samples1, const double* samples2, int k, int K, double* hbar);
/* R/debiasing_algorithm.R:284-316 with h = identity */
int orc_H_bar(const orc_chains* c, int k, int K, double* hbar) { return hbar_impl(c, c->samples1, c->samples2, k, K, hbar); }
/* the same estimator over exp_samples1/2 (what the all_particles experiments feed H_bar) */
int orc_H_bar_exp(const orc_chains* c, int k, int K, double* hbar) {
  if (!c->exp_samples1) return 2;
  return hbar_impl(c, c->exp_samples1, c->exp_samples2, k, K, hbar);
}
static int hbar_impl(const orc_chains* c, const double* samples1, const double* samples2, int k, int K, double* hbar) {
  int maxiter = c->iteration, p = c->p;
  if (k > maxiter || K > maxiter) return 1;                              /* :286-293 */
  for (int j = 0; j < p; j++) hbar[j] = 0;
  for (int t = k; t <= K; t++) for (int j = 0; j < p; j++) hbar[j] += samples1[(size_t)t * p + j];   /* :296-302 */
  int tau = c->meetingtime < 0 ? 2147483647 : c->meetingtime;
  if (!(tau <= k + 1)) {                                                 /* :303 */
    int tmax = maxiter - 1 < tau - 1 ? maxiter - 1 : tau - 1;
    double* dt = (double*)calloc((size_t)p, sizeof(double));
    for (int t = k; t <= tmax; t++) {                                    /* :308-312 */
      int coef = (t - k + 1) < (K - k + 1) ? (t - k + 1) : (K - k + 1);
      for (int j = 0; j < p; j++) dt[j] = dt[j] + coef * (samples1[(size_t)(t + 1) * p + j] - samples2[(size_t)t * p + j]);
    }
    for (int j = 0; j < p; j++) hbar[j] = hbar[j] + dt[j];
    free(dt);
  }
  for (int j = 0; j < p; j++) hbar[j] = hbar[j] / (K - k + 1);           /* :315 */
  return 0;
}

int orc_coupled_pimh_batch(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t first_replicate, int nrep,
                           int k, int K, int max_iterations, int nthreads,
                           double* hbar, int* meetingtime, int* iterations, long long* nfilters_total) {
  int rc = 0; long long nf = 0; int p = cfg->nobs + 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nf)
  for (int r = 0; r < nrep; r++) {
    orc_chains c;
    int e = orc_coupled_pimh_chains(cfg, obs, seed, first_replicate + (uint32_t)r, K, max_iterations, 0, &c);
    if (!e) {
      if (hbar) orc_H_bar(&c, k, K, hbar + (size_t)r * p);
      if (meetingtime) meetingtime[r] = c.meetingtime;
      if (iterations) iterations[r] = c.iteration;
      nf += c.nfilters;
      orc_chains_free(&c);
    } else {
#pragma omp atomic write
      rc = e;
    }
  }
  if (nfilters_total) *nfilters_total = nf;
  return rc;
}
