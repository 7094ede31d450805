/* oracle/apf.c -- restatement of pf(y, theta, model, nparticles, ess_threshold, systematic) (R/CPF.R:106-155): the bootstrap
 * filter with ESS-triggered resampling, carried log-weights, filtering means and cumulative log-likelihood that the PMMH
 * scripts use as likelihood evaluator.  SURVEY section 8(f) rank 3.  TEST INFRASTRUCTURE (see oracle.h).
 *
 * The reference's pf() talks to models through generate_randomness / row-major particles; here the same loop drives the
 * five hot-path models (oracle/models.c).  Order of a step (:123-147): propagate -> weigh (logw + log_W) -> loglik_t,
 * normalised weights, filtering mean -> resample iff (1 / sum(normweights^2)) / N < ess_threshold, by N multinomial draws
 * (R's sample(replace = TRUE, prob)) or systematic_resampling_n(normweights, N, runif(1)) -> Tree$update with the
 * RESAMPLED particles.  Randomness (DESIGN.md "RNG"): transition noise of step t as in the bootstrap filter; multinomial
 * draws are the sorted uniforms P_i / (P_{N-1} + E_N), E_i = -rng_log(ua(i, t, RESAMPLE)); the systematic uniform is
 * ua(0, t, AS); the final sample(1:N, 1, prob = normweights) (:152) uses ua(0, T+1, PATH) by inversion. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "rngmath.h"

int orc_model_rtransition_ex(const orc_config* c, int time, const double* noise, uint64_t seed, uint64_t fid, double* x, int* clamp);
static inline uint32_t c1_of(uint64_t filter_id, int t) { return (uint32_t)t | ((uint32_t)(filter_id >> 32) << 20); }
/* bits 12..19 of the iteration (high word of the filter id) go to the top byte of counter word 3 (DESIGN.md "RNG") */
static inline uint32_t c3_of(uint64_t filter_id, uint32_t c3) { return c3 | ((((uint32_t)(filter_id >> 32)) << 12) & 0xff000000u); }

int orc_pf_adaptive(const orc_config* cfg, const double* obs, uint64_t seed, uint64_t fid, double ess_threshold, int systematic,
                    orc_apf_out* out) {
  const int N = cfg->nparticles, d = cfg->dim, T = cfg->nobs, dy = cfg->dim_y;
  const int n_init = orc_model_n_init(cfg), n_trans = orc_model_n_trans(cfg);
  int rc = 0, clamp = 0;
  orc_tree* tree = orc_tree_new(N, 10 * N * d, d);                       /* :110 */
  double* x = (double*)malloc(sizeof(double) * (size_t)d * N);
  double* xg = (double*)malloc(sizeof(double) * (size_t)d * N);
  double* normw = (double*)malloc(sizeof(double) * (size_t)N);
  double* logw = (double*)malloc(sizeof(double) * (size_t)N);
  double* log_W = (double*)malloc(sizeof(double) * (size_t)N);
  double* nbuf = (double*)malloc(sizeof(double) * (size_t)N * (size_t)((n_trans > n_init ? n_trans : n_init) + 1));
  double* P = (double*)malloc(sizeof(double) * (size_t)(N + 1));
  double* yrow = (double*)malloc(sizeof(double) * (size_t)(dy > 0 ? dy : 1));
  int* anc = (int*)malloc(sizeof(int) * (size_t)N);
  orc_fill_init_noise(cfg, seed, fid, nbuf);                             /* :113-115 */
  orc_model_rinit(cfg, nbuf, x);
  orc_tree_init(tree, x);
  for (int i = 0; i < N; i++) log_W[i] = log(1.0 / N);                   /* :119 */
  double cum = 0;
  for (int time = 1; time <= T; time++) {
    const double* ntr = NULL;
    if (cfg->model != ORC_MODEL_SK) { orc_fill_trans_noise(cfg, seed, fid, time, nbuf); ntr = nbuf; }
    rc = orc_model_rtransition_ex(cfg, time, ntr, seed, fid, x, &clamp);  /* :124 */
    if (rc) goto done;
    for (int j = 0; j < dy; j++) yrow[j] = obs[(time - 1) + (size_t)T * j];
    orc_model_dmeasurement(cfg, x, yrow, logw);                          /* :125 */
    double m = -INFINITY;
    for (int i = 0; i < N; i++) { logw[i] = logw[i] + log_W[i]; if (logw[i] > m) m = logw[i]; }   /* :127-128 */
    long double s = 0;
    for (int i = 0; i < N; i++) { normw[i] = exp(logw[i] - m); s += normw[i]; }
    const double sd = (double)s;
    cum += m + log(sd);                                                  /* :130, cumsum at :154 */
    if (out->loglik_t) out->loglik_t[time - 1] = cum;
    for (int i = 0; i < N; i++) normw[i] = normw[i] / sd;                /* :131 */
    if (out->pf_means)                                                   /* :133 sum(normweights * x) per component */
      for (int q = 0; q < d; q++) {
        long double a = 0;
        for (int i = 0; i < N; i++) a += normw[i] * x[q + (size_t)i * d];
        out->pf_means[(size_t)(time - 1) * d + q] = (double)a;
      }
    long double s2 = 0;
    for (int i = 0; i < N; i++) s2 += normw[i] * normw[i];
    const int resample = ((1.0 / (double)s2) / N) < ess_threshold;       /* :135 */
    if (out->resampled) out->resampled[time - 1] = resample;
    if (resample) {
      if (!systematic) {                                                 /* :137 N iid draws, here as sorted uniforms */
        double acc = 0, ua, ub;
        for (int i = 0; i < N; i++) {
          orc_philox_uniform2(seed, (uint32_t)i, c1_of(fid, time), (uint32_t)fid, c3_of(fid, ORC_P_RESAMPLE), &ua, &ub);
          acc = (i == 0) ? -orc_rng_log(ua) : acc + (-orc_rng_log(ua));
          P[i] = acc;
        }
        orc_philox_uniform2(seed, (uint32_t)N, c1_of(fid, time), (uint32_t)fid, c3_of(fid, ORC_P_RESAMPLE), &ua, &ub);
        clamp += orc_multinomial_from_ladder(normw, N, N, P, acc - orc_rng_log(ua), anc);
      } else {                                                           /* :139 */
        double ua, ub;
        orc_philox_uniform2(seed, 0, c1_of(fid, time), (uint32_t)fid, c3_of(fid, ORC_P_AS), &ua, &ub);
        clamp += orc_systematic_resampling_n(normw, N, N, ua, anc);
      }
      for (int i = 0; i < N; i++) log_W[i] = log(1.0 / N);               /* :142 */
    } else {
      for (int i = 0; i < N; i++) { anc[i] = i; log_W[i] = log(normw[i]); }   /* :145-146 */
    }
    for (int i = 0; i < N; i++) memcpy(xg + (size_t)i * d, x + (size_t)anc[i] * d, sizeof(double) * d);   /* :148 */
    { double* t = x; x = xg; xg = t; }
    orc_tree_update(tree, x, anc);                                       /* :149 */
    if (out->ancestors) memcpy(out->ancestors + (size_t)(time - 1) * N, anc, sizeof(int) * (size_t)N);
  }
  {
    double ua, ub;
    orc_philox_uniform2(seed, 0, c1_of(fid, T + 1), (uint32_t)fid, c3_of(fid, ORC_P_PATH), &ua, &ub);
    int k;
    clamp += orc_systematic_resampling_n(normw, N, 1, ua, &k);           /* :152 one draw from normweights, by inversion */
    orc_tree_get_path(tree, k, out->path);
    out->path_index = k;
    out->loglik = cum;
    out->clamp_count = clamp;
  }
done:
  orc_tree_free(tree);
  free(x); free(xg); free(normw); free(logw); free(log_W); free(nbuf); free(P); free(yrow); free(anc);
  return rc;
}
