/* oracle/ccpf.c -- restatement of the CONDITIONAL particle filter CPF(ref_trajectory = x, with_as) (R/CPF.R:14-78, the
 * branches at :22-24 and :45-58), the coupled conditional filter CPF_coupled (R/CPF_coupled.R:14-112) with index-matching
 * coupled resampling (src/coupledresamplingindexmatching.cpp:8-57), the CCPF kernels (R/pf_kernels.R:236-265) and
 * coupled_ccpf_chains (R/debiasing_algorithm.R:23-111).  SURVEY section 8(f) rank 1.
 * TEST INFRASTRUCTURE (see oracle.h).
 *
 * Randomness (DESIGN.md "RNG", conditional filters): both systems read the SAME Philox streams (common random numbers,
 * R/CPF_coupled.R:24,29,57-59).  Per step t:
 *   E_i = -log(ua(i, t, RESAMPLE)), i = 0..nd+1, nd = N-1 (the last particle is the reference path, so N-1 are drawn)
 *   P_i = E_0 + .. + E_i
 *   single filter : sorted uniforms u_i = P_i / P_nd, i < nd                      (multinomial, src/multinomial.cpp:13-32)
 *   coupled filter: coupled_i = ua(i, t, COUPLE) < alpha, i < nd                   (coupledresamplingindexmatching.cpp:26-33)
 *                   the j-th coupled index draws from mu with u = P_j / P_nc       (:36-38)
 *                   the j-th other index draws from (R1, R2) with the common u = (P_{nc+1+j} - P_nc) / (P_{nd+1} - P_nc)  (:40-42)
 *   ancestor sampling: (uc, ud) = (ua, ub)(0, t, AS); final path: (uc, ud) = (ua, ub)(0, T+1, PATH): one index-matching draw
 * The reference's std::random_shuffle of each group is not applied (cfg->shuffle must be 0): drawing N-1 sorted
 * ancestors for the N-1 free particles is the same conditional filter up to a relabelling of the free particles.
 */
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

/* number of cdf knots <= u, clamped to n-1 (src/multinomial.cpp:25-28 read as a count) */
static int count_le(const double* cdf, int n, double u) {
  int j = n;
  while (j > 0 && u < cdf[j - 1]) j--;
  return j >= n ? n - 1 : j;
}
static void cumsum(const double* w, int n, double* cdf) {
  cdf[0] = w[0];
  for (int i = 1; i < n; i++) cdf[i] = cdf[i - 1] + w[i];
}

/* log transition density of the models that define dtransition: R/model_get_gss.R:30-33, R/model_get_ar.R:28-30 */
int orc_model_has_dtransition(const orc_config* c) { return c->model == ORC_MODEL_GSS || c->model == ORC_MODEL_AR; }
void orc_model_dtransition(const orc_config* c, const double* next_x, const double* x, int time, double* out) {
  const int N = c->nparticles, d = c->dim;
  if (c->model == ORC_MODEL_GSS) {
    double drift = 8 * cos(1.2 * (time - 1));
    for (int i = 0; i < N; i++) {
      double xi = x[i];
      double means = 0.5 * xi + 25 * xi / (1 + xi * xi) + drift;
      double z = (next_x[0] - means) / 1.0;
      out[i] = -(0.918938533204672741780329736406 + 0.5 * z * z + log(1.0));
    }
  } else {
    double* A = (double*)malloc(sizeof(double) * (size_t)d * d);
    orc_create_A(c->theta[0], d, A);
    for (int i = 0; i < N; i++) {                    /* dmvnorm_transpose_cholesky(A x, next_x, I): src/mvnorm.cpp:113-132 */
      double sq = 0;
      for (int r = 0; r < d; r++) {
        double s = 0;
        for (int l = 0; l < d; l++) s += A[r + (size_t)l * d] * x[l + (size_t)i * d];
        double e = s - next_x[r];
        sq += e * e;
      }
      out[i] = -0.5 * sq + (-0.0 - d * 0.9189385);   /* cst = -sum(log(diag(I))) - d * 0.9189385 (:120-121) */
    }
    free(A);
  }
}

/* one index-matching draw (indexmatching_cpp(w1, w2, 1)): coupled with probability alpha */
static void im_single_draw(const double* w1, const double* w2, int N, double uc, double ud, double* s0, double* s1, double* s2, int* a1, int* a2) {
  double alpha = 0;
  for (int i = 0; i < N; i++) { s0[i] = w1[i] < w2[i] ? w1[i] : w2[i]; alpha += s0[i]; }
  if (uc < alpha) {
    cumsum(s0, N, s1);
    *a1 = *a2 = count_le(s1, N, ud * s1[N - 1]);
  } else {
    for (int i = 0; i < N; i++) { s1[i] = w1[i] - s0[i]; s2[i] = w2[i] - s0[i]; }
    cumsum(s1, N, s1); cumsum(s2, N, s2);
    *a1 = count_le(s1, N, ud * s1[N - 1]);
    *a2 = count_le(s2, N, ud * s2[N - 1]);
  }
}

int orc_ccpf(const orc_config* cfg, const double* obs, const orc_noise* noise, uint64_t seed, uint64_t fid,
             int nsys, const double* ref1, const double* ref2, int with_as, orc_ccpf_out* out) {
  const int N = cfg->nparticles, d = cfg->dim, T = cfg->nobs, dy = cfg->dim_y, nd = N - 1;
  const int n_init = orc_model_n_init(cfg), n_trans = orc_model_n_trans(cfg);
  if (nsys < 1 || nsys > 2 || cfg->shuffle || N < 2) return -1;
  if (with_as && !orc_model_has_dtransition(cfg)) return -3;
  if (noise && (noise->resample || noise->shuffle || noise->path_u)) return -4;
  int rc = 0, clamp = 0;
  const double* ref[2] = {ref1, ref2};
  orc_tree* tree[2] = {0, 0};
  double *x[2], *xg[2], *xlast[2], *normw[2], *logw[2], *was[2];
  int* anc[2];
  for (int s = 0; s < 2; s++) {
    x[s] = (double*)malloc(sizeof(double) * (size_t)d * N); xg[s] = (double*)malloc(sizeof(double) * (size_t)d * N);
    xlast[s] = (double*)malloc(sizeof(double) * (size_t)d * N);
    normw[s] = (double*)malloc(sizeof(double) * (size_t)N); logw[s] = (double*)malloc(sizeof(double) * (size_t)N);
    was[s] = (double*)malloc(sizeof(double) * (size_t)N); anc[s] = (int*)malloc(sizeof(int) * (size_t)N);
  }
  double* nbuf = (double*)malloc(sizeof(double) * (size_t)N * (size_t)((n_trans > n_init ? n_trans : n_init) + 1));
  double* P = (double*)malloc(sizeof(double) * (size_t)(N + 2));
  double* s0 = (double*)malloc(sizeof(double) * (size_t)N);
  double* s1 = (double*)malloc(sizeof(double) * (size_t)N);
  double* s2 = (double*)malloc(sizeof(double) * (size_t)N);
  double* yrow = (double*)malloc(sizeof(double) * (size_t)(dy > 0 ? dy : 1));
  double* lzt = (double*)malloc(sizeof(double) * (size_t)T);

  const double* ninit = noise && noise->init ? noise->init : NULL;       /* R/CPF.R:19-24, R/CPF_coupled.R:23-32 */
  if (!ninit) { orc_fill_init_noise(cfg, seed, fid, nbuf); ninit = nbuf; }
  for (int s = 0; s < nsys; s++) {
    orc_model_rinit(cfg, ninit, x[s]);                                   /* common init_rand */
    memcpy(x[s] + (size_t)(N - 1) * d, ref[s], sizeof(double) * d);      /* xparticles[, N] <- ref[, 1] */
    tree[s] = orc_tree_new(N, 10 * N * d, d);
    orc_tree_init(tree[s], x[s]);
    for (int i = 0; i < N; i++) normw[s][i] = 1.0 / N;
    memcpy(xlast[s], x[s], sizeof(double) * (size_t)d * N);
  }

  for (int time = 1; time <= T; time++) {
    const int identity = (time == 1) || isnan(obs[(time - 2)]);
    double ua, ub;
    if (identity) {
      for (int s = 0; s < nsys; s++) for (int i = 0; i < N; i++) anc[s][i] = i;
    } else {
      double acc = 0;
      for (int i = 0; i <= nd + 1; i++) {
        orc_philox_uniform2(seed, (uint32_t)i, c1_of(fid, time), (uint32_t)fid, c3_of(fid, ORC_P_RESAMPLE), &ua, &ub);
        acc = (i == 0) ? -orc_rng_log(ua) : acc + (-orc_rng_log(ua));
        P[i] = acc;
      }
      if (nsys == 1) {
        clamp += orc_multinomial_from_ladder(normw[0], N, nd, P, P[nd], anc[0]);
      } else {
        double alpha = 0;                                                /* coupledresamplingindexmatching.cpp:13-21 */
        for (int i = 0; i < N; i++) { s0[i] = normw[0][i] < normw[1][i] ? normw[0][i] : normw[1][i]; alpha += s0[i]; }
        for (int i = 0; i < N; i++) { s1[i] = normw[0][i] - s0[i]; s2[i] = normw[1][i] - s0[i]; }
        cumsum(s0, N, s0); cumsum(s1, N, s1); cumsum(s2, N, s2);
        int nc = 0;
        for (int i = 0; i < nd; i++) {                                   /* :26-33 */
          orc_philox_uniform2(seed, (uint32_t)i, c1_of(fid, time), (uint32_t)fid, c3_of(fid, ORC_P_COUPLE), &ua, &ub);
          anc[0][i] = ua < alpha;
          nc += anc[0][i];
        }
        int jc = 0, jr = 0;
        for (int i = 0; i < nd; i++) {                                   /* :36-55, in index order */
          if (anc[0][i]) {
            double u = P[jc] / P[nc];
            anc[0][i] = anc[1][i] = count_le(s0, N, u * s0[N - 1]);
            jc++;
          } else {
            double u = (P[nc + 1 + jr] - P[nc]) / (P[nd + 1] - P[nc]);
            anc[0][i] = count_le(s1, N, u * s1[N - 1]);
            anc[1][i] = count_le(s2, N, u * s2[N - 1]);
            jr++;
          }
        }
      }
    }
    /* the reference particle: R/CPF.R:46-58, R/CPF_coupled.R:62-84 */
    if (with_as) {
      for (int s = 0; s < nsys; s++) {
        orc_model_dtransition(cfg, ref[s] + (size_t)time * d, xlast[s], time, was[s]);
        double m = -INFINITY;
        for (int i = 0; i < N; i++) { was[s][i] = log(normw[s][i]) + was[s][i]; if (was[s][i] > m) m = was[s][i]; }
        double sum = 0;
        for (int i = 0; i < N; i++) { was[s][i] = exp(was[s][i] - m); sum += was[s][i]; }
        for (int i = 0; i < N; i++) was[s][i] = was[s][i] / sum;
      }
      orc_philox_uniform2(seed, 0, c1_of(fid, time), (uint32_t)fid, c3_of(fid, ORC_P_AS), &ua, &ub);
      if (nsys == 1) { int k; clamp += orc_systematic_resampling_n(was[0], N, 1, ua, &k); anc[0][N - 1] = k; }
      else im_single_draw(was[0], was[1], N, ua, ub, s0, s1, s2, &anc[0][N - 1], &anc[1][N - 1]);
    } else {
      for (int s = 0; s < nsys; s++) anc[s][N - 1] = N - 1;
    }
    const double* ntr = NULL;
    if (noise && noise->trans) ntr = noise->trans + (size_t)(time - 1) * n_trans * N;
    else if (cfg->model != ORC_MODEL_SK) { orc_fill_trans_noise(cfg, seed, fid, time, nbuf); ntr = nbuf; }
    for (int j = 0; j < dy; j++) yrow[j] = obs[(time - 1) + (size_t)T * j];
    for (int s = 0; s < nsys; s++) {
      for (int i = 0; i < N; i++) memcpy(xg[s] + (size_t)i * d, x[s] + (size_t)anc[s][i] * d, sizeof(double) * d);
      { double* t = x[s]; x[s] = xg[s]; xg[s] = t; }
      rc = orc_model_rtransition_ex(cfg, time, ntr, seed, fid, x[s], &clamp);       /* common transition_rand */
      if (rc) goto done;
      memcpy(x[s] + (size_t)(N - 1) * d, ref[s] + (size_t)time * d, sizeof(double) * d);
      memcpy(xlast[s], x[s], sizeof(double) * (size_t)d * N);
      orc_model_dmeasurement(cfg, x[s], yrow, logw[s]);
      double lz = orc_logz_increment(logw[s], N, normw[s]);
      if (s == 0) lzt[time - 1] = lz;
      orc_tree_update(tree[s], x[s], anc[s]);
      if (out->ancestors[s]) memcpy(out->ancestors[s] + (size_t)(time - 1) * N, anc[s], sizeof(int) * (size_t)N);
    }
  }
  {
    double ua, ub;
    orc_philox_uniform2(seed, 0, c1_of(fid, T + 1), (uint32_t)fid, c3_of(fid, ORC_P_PATH), &ua, &ub);
    int k[2] = {0, 0};
    if (nsys == 1) clamp += orc_systematic_resampling_n(normw[0], N, 1, ua, &k[0]);          /* R/CPF.R:69 */
    else im_single_draw(normw[0], normw[1], N, ua, ub, s0, s1, s2, &k[0], &k[1]);            /* R/CPF_coupled.R:101-103 */
    for (int s = 0; s < nsys; s++) {
      orc_tree_get_path(tree[s], k[s], out->trajectory[s]);
      out->path_index[s] = k[s];
      if (out->exp_path[s]) {                        /* CPF_RB / CPF_RB_coupled with test_function = identity (R/rheeglynn_estimator_RB.R:50-56) */
        size_t psz = (size_t)d * (T + 1);
        double* path = (double*)malloc(sizeof(double) * psz);
        long double* acc = (long double*)calloc(psz, sizeof(long double));
        for (int n = 0; n < N; n++) {
          orc_tree_get_path(tree[s], n, path);
          for (size_t e = 0; e < psz; e++) acc[e] += normw[s][n] * path[e];
        }
        for (size_t e = 0; e < psz; e++) out->exp_path[s][e] = (double)acc[e];
        free(path); free(acc);
      }
    }
    long double sum = 0;
    for (int t = 0; t < T; t++) sum += lzt[t];
    out->logz = (double)sum;
    out->clamp_count = clamp;
  }
done:
  for (int s = 0; s < 2; s++) {
    if (tree[s]) orc_tree_free(tree[s]);
    free(x[s]); free(xg[s]); free(xlast[s]); free(normw[s]); free(logw[s]); free(was[s]); free(anc[s]);
  }
  free(nbuf); free(P); free(s0); free(s1); free(s2); free(yrow); free(lzt);
  return rc;
}

/* ---- coupled_ccpf_chains (R/debiasing_algorithm.R:23-111) with the kernels of get_cpf_kernels (R/pf_kernels.R:236-265)
 * and the initial states of pimh_init (R/pf_kernels.R:134-145: two independent bootstrap filters) ---- */
int orc_coupled_ccpf_chains(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t replicate,
                            int K, int max_iterations, int with_as, orc_chains* out) {
  return orc_coupled_ccpf_chains_rb(cfg, obs, seed, replicate, K, max_iterations, with_as, 0, out);
}
/* rb != 0: also record the weighted average of all N paths of every filter (exp_samples1/2): the Rao-Blackwellised
 * chains of rheeglynn_estimator_RB / RB2 (R/rheeglynn_estimator_RB.R:170-265) with test_function = identity */
int orc_coupled_ccpf_chains_rb(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t replicate,
                               int K, int max_iterations, int with_as, int rb, orc_chains* out) {
  const int d = cfg->dim, p = cfg->nobs + 1;
  const size_t psz = (size_t)d * p;
  double *s1 = (double*)malloc(sizeof(double) * psz), *s2 = (double*)malloc(sizeof(double) * psz);
  double *n1 = (double*)malloc(sizeof(double) * psz), *n2 = (double*)malloc(sizeof(double) * psz);
  double* lzt = (double*)malloc(sizeof(double) * (size_t)cfg->nobs);
  double *e1 = (double*)calloc(psz, sizeof(double)), *e2 = (double*)calloc(psz, sizeof(double));
  double *ne1 = (double*)calloc(psz, sizeof(double)), *ne2 = (double*)calloc(psz, sizeof(double));
  int cap = K + 10 + 1, nf = 0, rc = 0;
  double lz1 = 0, lz2 = 0;
  memset(out, 0, sizeof(*out));
  out->p = p;
  out->samples1 = (double*)malloc(sizeof(double) * (size_t)cap * p);
  out->samples2 = (double*)malloc(sizeof(double) * (size_t)cap * p);
  out->log_pdf1 = (double*)malloc(sizeof(double) * (size_t)cap);
  out->log_pdf2 = (double*)malloc(sizeof(double) * (size_t)cap);
  if (rb) {
    out->exp_samples1 = (double*)malloc(sizeof(double) * (size_t)cap * p);
    out->exp_samples2 = (double*)malloc(sizeof(double) * (size_t)cap * p);
  }
  orc_pf_out o; memset(&o, 0, sizeof(o));
  o.logz_t = lzt;
  o.trajectory = s1; o.logz = &lz1; o.exp_path = rb ? e1 : NULL;
  rc = orc_cpf(cfg, obs, NULL, seed, (uint64_t)replicate, &o); nf++;                               /* pimh_init: chain_state1 */
  o.trajectory = s2; o.logz = &lz2; o.exp_path = rb ? e2 : NULL;
  if (!rc) { rc = orc_cpf(cfg, obs, NULL, seed, (uint64_t)replicate | ((uint64_t)ORC_MAX_ITERATION << 32), &o); nf++; }   /* chain_state2 */
  if (rc) goto fail;
  for (int t = 0; t < p; t++) { out->samples1[t] = s1[(size_t)t * d]; out->samples2[t] = s2[(size_t)t * d]; }   /* :42-43 */
  out->log_pdf1[0] = lz1; out->log_pdf2[0] = lz2;
  if (rb) for (int t = 0; t < p; t++) { out->exp_samples1[t] = e1[(size_t)t * d]; out->exp_samples2[t] = e2[(size_t)t * d]; }
  orc_ccpf_out co;
  int iter = 1, meet = 0, finished = 0, meetingtime = -1;
  memset(&co, 0, sizeof(co)); co.trajectory[0] = n1; co.trajectory[1] = n2;
  if (rb) { co.exp_path[0] = ne1; co.exp_path[1] = ne2; }
  rc = orc_ccpf(cfg, obs, NULL, seed, (uint64_t)replicate | ((uint64_t)iter << 32), 1, s1, NULL, with_as, &co); nf++;   /* :50-52 */
  if (rc) goto fail;
  memcpy(s1, n1, sizeof(double) * psz); lz1 = co.logz;
  memcpy(e1, ne1, sizeof(double) * psz);
  if (rb) for (int t = 0; t < p; t++) out->exp_samples1[(size_t)p + t] = e1[(size_t)t * d];
  for (int t = 0; t < p; t++) out->samples1[(size_t)p + t] = s1[(size_t)t * d];                    /* :56-58 */
  out->log_pdf1[1] = lz1;
  while (!finished && (max_iterations < 0 || iter < max_iterations)) {                             /* :63 */
    iter++;
    if (meet) {                                                                                    /* :68-73 */
      rc = orc_ccpf(cfg, obs, NULL, seed, (uint64_t)replicate | ((uint64_t)iter << 32), 1, s1, NULL, with_as, &co); nf++;
      if (rc) goto fail;
      memcpy(s1, n1, sizeof(double) * psz); memcpy(s2, n1, sizeof(double) * psz);
      memcpy(e1, ne1, sizeof(double) * psz); memcpy(e2, ne1, sizeof(double) * psz);
      lz1 = lz2 = co.logz;
    } else {                                                                                       /* :75-87 */
      rc = orc_ccpf(cfg, obs, NULL, seed, (uint64_t)replicate | ((uint64_t)iter << 32), 2, s1, s2, with_as, &co); nf++;
      if (rc) goto fail;
      memcpy(s1, n1, sizeof(double) * psz); memcpy(s2, n2, sizeof(double) * psz);                  /* log_pdf states unchanged */
      memcpy(e1, ne1, sizeof(double) * psz); memcpy(e2, ne2, sizeof(double) * psz);
      if (memcmp(s1, s2, sizeof(double) * psz) == 0) { meet = 1; meetingtime = iter; }
    }
    if (iter + 1 > cap) {
      cap = 2 * cap;
      out->samples1 = (double*)realloc(out->samples1, sizeof(double) * (size_t)cap * p);
      out->samples2 = (double*)realloc(out->samples2, sizeof(double) * (size_t)cap * p);
      out->log_pdf1 = (double*)realloc(out->log_pdf1, sizeof(double) * (size_t)cap);
      out->log_pdf2 = (double*)realloc(out->log_pdf2, sizeof(double) * (size_t)cap);
      if (rb) {
        out->exp_samples1 = (double*)realloc(out->exp_samples1, sizeof(double) * (size_t)cap * p);
        out->exp_samples2 = (double*)realloc(out->exp_samples2, sizeof(double) * (size_t)cap * p);
      }
    }
    if (rb) for (int t = 0; t < p; t++) {
      out->exp_samples1[(size_t)iter * p + t] = e1[(size_t)t * d];
      out->exp_samples2[(size_t)(iter - 1) * p + t] = e2[(size_t)t * d];
    }
    for (int t = 0; t < p; t++) {                                                                  /* :98-99 */
      out->samples1[(size_t)iter * p + t] = s1[(size_t)t * d];
      out->samples2[(size_t)(iter - 1) * p + t] = s2[(size_t)t * d];
    }
    out->log_pdf1[iter] = lz1; out->log_pdf2[iter - 1] = lz2;
    int stop_at = meetingtime < 0 ? -1 : (meetingtime > K ? meetingtime : K);
    if (stop_at >= 0 && iter >= stop_at) finished = 1;                                             /* :104-106 */
  }
  out->iteration = iter; out->meetingtime = meetingtime; out->finished = finished; out->nrows1 = iter + 1; out->nfilters = nf;
fail:
  free(s1); free(s2); free(n1); free(n2); free(lzt); free(e1); free(e2); free(ne1); free(ne2);
  return rc;
}

int orc_coupled_ccpf_batch(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t first_replicate, int nrep,
                           int k, int K, int max_iterations, int with_as, int nthreads,
                           double* hbar, double* hbar_rb, int* meetingtime, int* iterations, long long* nfilters_total) {
  int rc = 0; long long nf = 0; int p = cfg->nobs + 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nf)
  for (int r = 0; r < nrep; r++) {
    orc_chains c;
    int e = orc_coupled_ccpf_chains_rb(cfg, obs, seed, first_replicate + (uint32_t)r, K, max_iterations, with_as, hbar_rb != NULL, &c);
    if (!e) {
      if (hbar) orc_H_bar(&c, k, K, hbar + (size_t)r * p);
      if (hbar_rb) orc_H_bar_exp(&c, k, K, hbar_rb + (size_t)r * p);
      if (meetingtime) meetingtime[r] = c.meetingtime;
      if (iterations) iterations[r] = c.iteration;
      nf += c.nfilters;
    } else {
#pragma omp atomic write
      rc = e;
    }
    orc_chains_free(&c);
  }
  if (nfilters_total) *nfilters_total = nf;
  return rc;
}
