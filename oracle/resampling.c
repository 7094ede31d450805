/* oracle/resampling.c -- restatement of the reference resamplers with INJECTED uniforms.
 * TEST INFRASTRUCTURE (see oracle.h).  Arithmetic order is the reference's: sequential fp64 running
 * sums, repeated addition for the stratified points, strict comparisons as written. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* src/systematic.cpp:7-32.  Returns the number of times the walk was clamped at N-1: the reference
 * walks past the end when sum(w) < u; oracle and GPU both clamp and count. */
int orc_systematic_resampling_n(const double* w, int nparticles, int ndraws, double u, int* anc) {
  int clamped = 0;
  u = u / ndraws;                                   /* :11 */
  int j = 0;                                        /* :12 */
  double csw = w[0];                                /* :13 */
  for (int k = 0; k < ndraws; k++) {                /* :14 */
    while (csw < u) {                               /* :15 strict < */
      if (j >= nparticles - 1) { clamped++; break; }
      j++;                                          /* :16 */
      csw += w[j];                                  /* :17 running fp64 sum */
    }
    u = u + 1. / ndraws;                            /* :19 repeated addition */
    anc[k] = j;                                     /* :20 */
  }
  return clamped;
}

/* src/multinomial.cpp:20-24: lnMax += log(U[i-1]) / i for i = n..1, e_(i) = exp(lnMax): ascending in i. */
void orc_sorted_uniforms(const double* u_raw, int n, double* e_sorted) {
  double lnMax = 0;
  for (int i = n; i > 0; i--) {
    lnMax += log(u_raw[i - 1]) / i;
    e_sorted[i - 1] = exp(lnMax);
  }
}

/* src/multinomial.cpp:17,19-28 (scale_by_sumw=1) and :40-41,49-57 for one column (scale_by_sumw=0).
 * cumsum is Rcpp sugar's: plain sequential double additions.  Returns #clamped (ancestor == N, H8). */
int orc_multinomial_from_sorted(const double* w, int nparticles, int ndraws, const double* e_sorted, int scale_by_sumw, int* anc) {
  double* cdf = (double*)malloc(sizeof(double) * (size_t)nparticles);
  cdf[0] = w[0];
  for (int i = 1; i < nparticles; i++) cdf[i] = cdf[i - 1] + w[i];   /* Rcpp::cumsum */
  double sumw = cdf[nparticles - 1];                                  /* :19 */
  int j = nparticles, clamped = 0;                                    /* :21 */
  for (int i = ndraws; i > 0; i--) {
    double u = scale_by_sumw ? sumw * e_sorted[i - 1] : e_sorted[i - 1];  /* :24 / :49 */
    while (j > 0 && u < cdf[j - 1]) j--;                              /* :25-27 */
    int a = j;
    if (a >= nparticles) { a = nparticles - 1; clamped++; }           /* H8 clamp */
    anc[i - 1] = a;                                                   /* :28 */
  }
  free(cdf);
  return clamped;
}

/* Multinomial resampling against an exponential-spacing ladder P_i (prefix sums) with grand total `total`:
 * u_i = (sumw / total) * P_i, then the same backward merge as src/multinomial.cpp:25-28. */
int orc_multinomial_from_ladder(const double* w, int nparticles, int ndraws, const double* P, double total, int* anc) {
  double* cdf = (double*)malloc(sizeof(double) * (size_t)nparticles);
  cdf[0] = w[0];
  for (int i = 1; i < nparticles; i++) cdf[i] = cdf[i - 1] + w[i];
  const double scale = cdf[nparticles - 1] / total;
  int j = nparticles, clamped = 0;
  for (int i = ndraws; i > 0; i--) {
    double u = scale * P[i - 1];
    while (j > 0 && u < cdf[j - 1]) j--;
    int a = j;
    if (a >= nparticles) { a = nparticles - 1; clamped++; }
    anc[i - 1] = a;
  }
  free(cdf);
  return clamped;
}

/* libstdc++ std::random_shuffle(first,last,rand) as called at src/multinomial.cpp:30,65 with
 * randWrapper(n) = floor(runif(1)*n) (:6-10): for i = 1..n-1: swap(a[i], a[rand(i+1)]). */
void orc_random_shuffle_int(int* a, int n, const double* u_shuffle) {
  for (int i = 1; i < n; i++) {
    int j = (int)floor(u_shuffle[i - 1] * (double)(i + 1));
    if (j > i) j = i;
    int tmp = a[i]; a[i] = a[j]; a[j] = tmp;
  }
}

/* src/multinomial.cpp:13-32 with raw uniforms injected (n for the draws, n-1 for the shuffle or NULL). */
int orc_multinomial_resampling_n(const double* w, int nparticles, int ndraws, const double* u_raw, const double* u_shuffle, int* anc) {
  double* e = (double*)malloc(sizeof(double) * (size_t)ndraws);
  orc_sorted_uniforms(u_raw, ndraws, e);
  int c = orc_multinomial_from_sorted(w, nparticles, ndraws, e, 1, anc);
  if (u_shuffle) orc_random_shuffle_int(anc, ndraws, u_shuffle);
  free(e);
  return c;
}

/* src/multinomial.cpp:35-71: same sorted uniforms (NOT scaled by sumw, :49) against both CDFs, one shared
 * permutation applied to both columns (:60-69).  anc is ndraws x 2 column-major, 0-based. */
int orc_coupled_multinomial_resampling_n(const double* w1, const double* w2, int nparticles, int ndraws,
                                         const double* u_raw, const double* u_shuffle, int* anc) {
  double* e = (double*)malloc(sizeof(double) * (size_t)ndraws);
  int* a1 = (int*)malloc(sizeof(int) * (size_t)ndraws);
  int* a2 = (int*)malloc(sizeof(int) * (size_t)ndraws);
  int* idx = (int*)malloc(sizeof(int) * (size_t)ndraws);
  orc_sorted_uniforms(u_raw, ndraws, e);
  int c = orc_multinomial_from_sorted(w1, nparticles, ndraws, e, 0, a1);
  c += orc_multinomial_from_sorted(w2, nparticles, ndraws, e, 0, a2);
  for (int i = 0; i < ndraws; i++) idx[i] = i;
  if (u_shuffle) orc_random_shuffle_int(idx, ndraws, u_shuffle);
  for (int i = 0; i < ndraws; i++) { anc[i] = a1[idx[i]]; anc[ndraws + i] = a2[idx[i]]; }
  free(e); free(a1); free(a2); free(idx);
  return c;
}

/* src/coupledresamplingindexmatching.cpp:8-57.  u_couple has N entries (only the first ndraws are read,
 * :27); the nested resamplers get their own injected uniforms (sized for the worst case ndraws). */
int orc_indexmatching(const double* w1, const double* w2, int nparticles, int ndraws, const double* u_couple,
                      const double* u_mult_raw, const double* u_mult_shuffle,
                      const double* u_cm_raw, const double* u_cm_shuffle, int* anc) {
  int N = nparticles, c = 0;
  double* nu = (double*)malloc(sizeof(double) * (size_t)N * 4);
  double *mu = nu + N, *R1 = nu + 2 * N, *R2 = nu + 3 * N;
  double alpha = 0;
  for (int i = 0; i < N; i++) { nu[i] = w1[i] < w2[i] ? w1[i] : w2[i]; alpha += nu[i]; }  /* :13-16 std::min */
  for (int i = 0; i < N; i++) {
    mu[i] = nu[i] / alpha;                       /* :18 */
    R1[i] = (w1[i] - nu[i]) / (1 - alpha);       /* :20 */
    R2[i] = (w2[i] - nu[i]) / (1 - alpha);       /* :21 */
  }
  int ncoupled = 0;
  char* coupled = (char*)malloc((size_t)ndraws);
  for (int i = 0; i < ndraws; i++) { coupled[i] = u_couple[i] < alpha; ncoupled += coupled[i]; }  /* :26-33 */
  int* a = (int*)malloc(sizeof(int) * (size_t)(ndraws > 0 ? ndraws : 1));
  int* aR = (int*)malloc(sizeof(int) * (size_t)(ndraws > 0 ? 2 * ndraws : 1));
  int nres = ndraws - ncoupled;
  if (ncoupled > 0) c += orc_multinomial_resampling_n(mu, N, ncoupled, u_mult_raw, u_mult_shuffle, a);          /* :36-38 */
  if (nres > 0) c += orc_coupled_multinomial_resampling_n(R1, R2, N, nres, u_cm_raw, u_cm_shuffle, aR);         /* :40-42 */
  int cc = 0, uc = 0;
  for (int i = 0; i < ndraws; i++) {                                                                            /* :45-55 */
    if (coupled[i]) { anc[i] = a[cc]; anc[ndraws + i] = a[cc]; cc++; }
    else { anc[i] = aR[uc]; anc[ndraws + i] = aR[nres + uc]; uc++; }
  }
  free(nu); free(coupled); free(a); free(aR);
  return c;
}

/* R/util_normalize_weight.R:5-9: w = exp(logw - max); w / sum(w) with R's long-double sum(). */
void orc_normalize_weight(const double* logw, int n, double* w) {
  double m = logw[0];
  for (int i = 1; i < n; i++) if (logw[i] > m) m = logw[i];
  long double s = 0;
  for (int i = 0; i < n; i++) { w[i] = exp(logw[i] - m); s += w[i]; }
  double sd = (double)s;
  for (int i = 0; i < n; i++) w[i] = w[i] / sd;
}

/* R/CPF.R:61-63,66: normweights as above and logz_t = log(mean(exp(logw - maxlw))) + maxlw, with R's
 * two-pass long-double mean() (summary.c: s = sum/n; s += sum(x - s)/n). */
double orc_logz_increment(const double* logw, int n, double* normw) {
  double m = logw[0];
  for (int i = 1; i < n; i++) if (logw[i] > m) m = logw[i];
  long double s = 0;
  for (int i = 0; i < n; i++) { normw[i] = exp(logw[i] - m); s += normw[i]; }
  long double mean = s / n, t = 0;
  for (int i = 0; i < n; i++) t += (normw[i] - mean);
  mean += t / n;
  double sd = (double)s;
  for (int i = 0; i < n; i++) normw[i] = normw[i] / sd;
  return log((double)mean) + m;
}
