/* shim.c -- binds the reference package's registered .Call entry points (src/RcppExports.cpp:346-375 of
 * particlemontecarlo/unbiased_pimh) to libcpimh.so (include/cpimh.h).  Plain R C API: no Rcpp, RcppEigen or BH.
 * Natives return 0-based indices exactly like the reference's C++; the +1 stays in the R wrappers.
 * Randomness: natives that drew from R's RNG in the reference (runif/rnorm inside the C++) draw the same number of
 * variates from R's stream here and pass them to the library as injected noise, so set.seed() semantics are kept;
 * the batched entries take one 64-bit seed from R's stream per call and use Philox streams on the device. */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Random.h>
#include <math.h>
#include <stdint.h>
#include "cpimh.h"

static void check(int rc) { if (rc) Rf_error("cpimh: %s", cpimh_last_error()); }
static uint64_t seed_from_R(void) {
  GetRNGstate();
  uint64_t hi = (uint64_t)(unif_rand() * 4294967296.0), lo = (uint64_t)(unif_rand() * 4294967296.0);
  PutRNGstate();
  return (hi << 32) | lo;
}
static double* runif_R(int n) { double* u = (double*)R_alloc(n > 0 ? n : 1, sizeof(double)); GetRNGstate(); for (int i = 0; i < n; i++) u[i] = unif_rand(); PutRNGstate(); return u; }
static double* rnorm_R(int n) { double* z = (double*)R_alloc(n > 0 ? n : 1, sizeof(double)); GetRNGstate(); for (int i = 0; i < n; i++) z[i] = norm_rand(); PutRNGstate(); return z; }
static double* sorted_R(int n) {            /* runif(n) + the recurrence of src/multinomial.cpp:20-24 (on the device) */
  double* u = runif_R(n), *e = (double*)R_alloc(n > 0 ? n : 1, sizeof(double));
  if (n > 0) check(cpimh_sorted_uniforms(u, 1, n, e));
  return e;
}

/* ---- resamplers ---- */
SEXP _CoupledPIMH_systematic_resampling_n_(SEXP w, SEXP ndraws, SEXP u) {               /* src/systematic.cpp:7 */
  int n = Rf_asInteger(ndraws); double uu = Rf_asReal(u);
  SEXP out = PROTECT(Rf_allocVector(INTSXP, n));
  check(cpimh_systematic_resampling_n(REAL(w), 1, LENGTH(w), n, &uu, INTEGER(out)));
  UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_multinomial_resampling_n_(SEXP w, SEXP ndraws) {                      /* src/multinomial.cpp:13 */
  int n = Rf_asInteger(ndraws);
  double* e = sorted_R(n); double* us = runif_R(n - 1);                                  /* RNG order: n draws, then n-1 for the shuffle */
  SEXP out = PROTECT(Rf_allocVector(INTSXP, n));
  check(cpimh_multinomial_resampling_n(REAL(w), 1, LENGTH(w), n, e, n > 1 ? us : NULL, INTEGER(out)));
  UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_coupled_multinomial_resampling_n_(SEXP w1, SEXP w2, SEXP ndraws) {    /* src/multinomial.cpp:35 */
  int n = Rf_asInteger(ndraws);
  double* e = sorted_R(n); double* us = runif_R(n - 1);
  SEXP out = PROTECT(Rf_allocMatrix(INTSXP, n, 2));
  check(cpimh_coupled_multinomial_resampling_n(REAL(w1), REAL(w2), 1, LENGTH(w1), n, e, n > 1 ? us : NULL, INTEGER(out)));
  UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_indexmatching_cpp(SEXP w1, SEXP w2, SEXP ndraws) {                    /* src/coupledresamplingindexmatching.cpp:8 */
  int n = Rf_asInteger(ndraws), N = LENGTH(w1);
  double* uc = runif_R(N);                                                                /* :10 runif(nparticles) */
  double alpha = 0; const double *a = REAL(w1), *b = REAL(w2);
  for (int i = 0; i < N; i++) alpha += a[i] < b[i] ? a[i] : b[i];                         /* :13-16, to size the nested draws */
  int nc = 0; for (int i = 0; i < n; i++) nc += uc[i] < alpha;
  double *em = (double*)R_alloc(n, sizeof(double)), *ecm = (double*)R_alloc(n, sizeof(double)), *usm = NULL, *uscm = NULL;
  for (int i = 0; i < n; i++) em[i] = ecm[i] = 0.5;
  if (nc > 0) { double* t = sorted_R(nc); for (int i = 0; i < nc; i++) em[i] = t[i]; double* s = runif_R(nc - 1); usm = (double*)R_alloc(n, sizeof(double)); for (int i = 0; i + 1 < nc; i++) usm[i] = s[i]; }
  if (nc < n) { double* t = sorted_R(n - nc); for (int i = 0; i < n - nc; i++) ecm[i] = t[i]; double* s = runif_R(n - nc - 1); uscm = (double*)R_alloc(n, sizeof(double)); for (int i = 0; i + 1 < n - nc; i++) uscm[i] = s[i]; }
  SEXP out = PROTECT(Rf_allocMatrix(INTSXP, n, 2));
  check(cpimh_indexmatching(a, b, 1, N, n, uc, em, usm, ecm, uscm, INTEGER(out)));
  UNPROTECT(1); return out;
}

/* ---- Gaussian kernels (src/mvnorm.cpp) ---- */
SEXP _CoupledPIMH_dmvnorm(SEXP x, SEXP mean, SEXP cov) {
  int n = Rf_nrows(x), d = Rf_ncols(x); SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  check(cpimh_dmvnorm(REAL(x), n, d, REAL(mean), REAL(cov), REAL(out))); UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_dmvnorm_transpose(SEXP x, SEXP mean, SEXP cov) {
  int d = Rf_nrows(x), n = Rf_ncols(x); SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  check(cpimh_dmvnorm_transpose(REAL(x), d, n, REAL(mean), REAL(cov), REAL(out))); UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_dmvnorm_transpose_cholesky(SEXP x, SEXP mean, SEXP chol) {
  int d = Rf_nrows(x), n = Rf_ncols(x); SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  check(cpimh_dmvnorm_transpose_cholesky(REAL(x), d, n, REAL(mean), REAL(chol), REAL(out))); UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_rmvnorm(SEXP nsamples, SEXP mean, SEXP cov) {                          /* rnorm(nsamples) per column (:12-14) */
  int n = Rf_asInteger(nsamples), d = Rf_ncols(cov); double* z = rnorm_R(n * d);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, d));
  check(cpimh_rmvnorm(n, d, REAL(mean), REAL(cov), 0, z, REAL(out))); UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_rmvnorm_transpose(SEXP nsamples, SEXP mean, SEXP cov) {                /* rnorm(dimension) per sample (:32-34) */
  int n = Rf_asInteger(nsamples), d = Rf_ncols(cov); double* z = rnorm_R(n * d);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, n));
  check(cpimh_rmvnorm_transpose(n, d, REAL(mean), REAL(cov), 0, z, REAL(out))); UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_rmvnorm_transpose_cholesky(SEXP nsamples, SEXP mean, SEXP chol) {
  int n = Rf_asInteger(nsamples), d = Rf_ncols(chol); double* z = rnorm_R(n * d);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, n));
  check(cpimh_rmvnorm_transpose_cholesky(n, d, REAL(mean), REAL(chol), 0, z, REAL(out))); UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_create_A_(SEXP alpha, SEXP d_) {                                       /* src/create_A_.cpp:6 */
  int d = Rf_asInteger(d_); SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, d));
  check(cpimh_create_A(Rf_asReal(alpha), d, REAL(out))); UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_one_step_lorenz_vector(SEXP x, SEXP tstart, SEXP tend, SEXP h, SEXP par) {   /* src/lorenz96.cpp:65 */
  int d = Rf_nrows(x), n = Rf_ncols(x); SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, n));
  check(cpimh_lorenz_transition(REAL(x), d, n, Rf_asReal(tstart), Rf_asReal(tend), Rf_asReal(h), REAL(par)[0], CPIMH_INTEGRATOR_DOPRI5, 1, REAL(out)));
  UNPROTECT(1); return out;
}
/* Levy-SV natives (src/SVLevy_functions.cpp:40-81): X is N x 2 in the reference (rows = particles) */
static void sv_config(cpimh_config* c, int N, double xi, double w2, double lambda, double mu, double beta) {
  cpimh_config_default(c); c->model = CPIMH_MODEL_LEVY_SV; c->nparticles = N; c->nobs = 1; c->ntheta = 5;
  c->theta[0] = mu; c->theta[1] = beta; c->theta[2] = xi; c->theta[3] = w2; c->theta[4] = lambda;
}
SEXP _CoupledPIMH_rinitial_SVLevy_cpp(SEXP N_, SEXP t, SEXP xi, SEXP w2, SEXP lambda) {
  int N = Rf_asInteger(N_); cpimh_config c; sv_config(&c, N, Rf_asReal(xi), Rf_asReal(w2), Rf_asReal(lambda), 0, 0);
  double* x = (double*)R_alloc(2 * N, sizeof(double));
  check(cpimh_model_rinit(&c, NULL, seed_from_R(), 0, x));                               /* d x N */
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, N, 2));
  for (int i = 0; i < N; i++) { REAL(out)[i] = x[2 * i]; REAL(out)[N + i] = x[2 * i + 1]; }
  UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_rtransition_SVLevy_cpp(SEXP Xold, SEXP t_1, SEXP t, SEXP xi, SEXP w2, SEXP lambda) {
  int N = Rf_nrows(Xold); cpimh_config c; sv_config(&c, N, Rf_asReal(xi), Rf_asReal(w2), Rf_asReal(lambda), 0, 0);
  int time = (int)Rf_asReal(t); c.nobs = time > 0 ? time : 1;
  double* x = (double*)R_alloc(2 * N, sizeof(double));
  for (int i = 0; i < N; i++) { x[2 * i] = REAL(Xold)[i]; x[2 * i + 1] = REAL(Xold)[N + i]; }
  check(cpimh_model_rtransition(&c, time > 0 ? time : 1, NULL, seed_from_R(), 0, x));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, N, 2));
  for (int i = 0; i < N; i++) { REAL(out)[i] = x[2 * i]; REAL(out)[N + i] = x[2 * i + 1]; }
  UNPROTECT(1); return out;
}
SEXP _CoupledPIMH_dobs_SVLevy_cpp(SEXP Yt, SEXP Xts, SEXP mu, SEXP beta, SEXP log_) {
  int N = Rf_nrows(Xts); cpimh_config c; sv_config(&c, N, 1, 1, 1, Rf_asReal(mu), Rf_asReal(beta));
  double* x = (double*)R_alloc(2 * N, sizeof(double));
  for (int i = 0; i < N; i++) { x[2 * i] = REAL(Xts)[i]; x[2 * i + 1] = REAL(Xts)[N + i]; }
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, 1, N));
  check(cpimh_model_dmeasurement(&c, x, REAL(Yt), REAL(out)));                            /* log density */
  if (!Rf_asLogical(log_)) for (int i = 0; i < N; i++) REAL(out)[i] = exp(REAL(out)[i]);  /* dobs_SVLevy_cpp(..., log = FALSE), src/SVLevy_functions.cpp:78 */
  UNPROTECT(1); return out;
}

/* ---- class Tree (src/tree.cpp:144-162): external pointer + one entry per method ---- */
static void tree_fin(SEXP p) { cpimh_tree* t = (cpimh_tree*)R_ExternalPtrAddr(p); if (t) { cpimh_tree_destroy(t); R_ClearExternalPtr(p); } }
SEXP cpimh_R_tree_new(SEXP N, SEXP M, SEXP dimx) {
  cpimh_tree* t = NULL; check(cpimh_tree_create(Rf_asInteger(N), Rf_asInteger(M), Rf_asInteger(dimx), &t));
  SEXP p = PROTECT(R_MakeExternalPtr(t, R_NilValue, R_NilValue)); R_RegisterCFinalizerEx(p, tree_fin, TRUE); UNPROTECT(1); return p;
}
SEXP cpimh_R_tree_init(SEXP p, SEXP x0) { check(cpimh_tree_init((cpimh_tree*)R_ExternalPtrAddr(p), REAL(x0))); return R_NilValue; }
SEXP cpimh_R_tree_update(SEXP p, SEXP x, SEXP a) { check(cpimh_tree_update((cpimh_tree*)R_ExternalPtrAddr(p), REAL(x), INTEGER(a))); return R_NilValue; }
SEXP cpimh_R_tree_get_path(SEXP p, SEXP n) {
  cpimh_tree* t = (cpimh_tree*)R_ExternalPtrAddr(p); int N, M, d, ns; check(cpimh_tree_fields(t, &N, &M, &d, &ns));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, ns + 1)); check(cpimh_tree_get_path(t, Rf_asInteger(n), REAL(out))); UNPROTECT(1); return out;
}
SEXP cpimh_R_tree_retrieve_xgeneration(SEXP p, SEXP lag) {
  cpimh_tree* t = (cpimh_tree*)R_ExternalPtrAddr(p); int N, M, d, ns; check(cpimh_tree_fields(t, &N, &M, &d, &ns));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, N)); check(cpimh_tree_retrieve_xgeneration(t, Rf_asInteger(lag), REAL(out))); UNPROTECT(1); return out;
}

/* fields of the module class (src/tree.cpp:146-154): which = 0 N, 1 M, 2 dimx, 3 nsteps, 4 a_star, 5 o_star, 6 x_star, 7 l_star */
SEXP cpimh_R_tree_field(SEXP p, SEXP which) {
  cpimh_tree* t = (cpimh_tree*)R_ExternalPtrAddr(p); int N, M, d, ns, w = Rf_asInteger(which);
  check(cpimh_tree_fields(t, &N, &M, &d, &ns));
  if (w < 4) { SEXP out = PROTECT(Rf_allocVector(INTSXP, 1)); INTEGER(out)[0] = w == 0 ? N : w == 1 ? M : w == 2 ? d : ns; UNPROTECT(1); return out; }
  if (w == 6) { SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, M)); check(cpimh_tree_arrays(t, NULL, NULL, NULL, REAL(out))); UNPROTECT(1); return out; }
  SEXP out = PROTECT(Rf_allocVector(INTSXP, w == 7 ? N : M));
  check(cpimh_tree_arrays(t, w == 4 ? INTEGER(out) : NULL, w == 5 ? INTEGER(out) : NULL, w == 7 ? INTEGER(out) : NULL, NULL));
  UNPROTECT(1); return out;
}
SEXP cpimh_R_tree_reset(SEXP p) { check(cpimh_tree_reset((cpimh_tree*)R_ExternalPtrAddr(p))); return R_NilValue; }

/* ---- batched entries the R closures dispatch to by model name ---- */
SEXP cpimh_R_pf_batch(SEXP cfg_raw, SEXP obs, SEXP nfilters, SEXP all_particles) {       /* CPF(ref=NULL[, all_particles]), R/CPF.R:14-78, B at once */
  const cpimh_config* cfg = (const cpimh_config*)RAW(cfg_raw);
  int B = Rf_asInteger(nfilters), d = cfg->dim, T = cfg->nobs, ap = Rf_asLogical(all_particles);
  SEXP logz = PROTECT(Rf_allocVector(REALSXP, B));
  SEXP traj = PROTECT(Rf_alloc3DArray(REALSXP, d, T + 1, B));
  SEXP expp = PROTECT(Rf_alloc3DArray(REALSXP, d, T + 1, ap ? B : 0));                  /* exp_path (R/CPF.R:73-76) */
  if (ap) check(cpimh_pf_batch_all_particles(cfg, REAL(obs), B, seed_from_R(), 0, REAL(logz), REAL(traj), REAL(expp)));
  else check(cpimh_pf_batch(cfg, REAL(obs), B, seed_from_R(), 0, NULL, REAL(logz), NULL, REAL(traj), NULL));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3)); SET_VECTOR_ELT(out, 0, logz); SET_VECTOR_ELT(out, 1, traj); SET_VECTOR_ELT(out, 2, expp);
  UNPROTECT(4); return out;
}
/* pf(y, theta, model, nparticles, ess_threshold, systematic) for B filters (R/CPF.R:106-155): log-likelihood, filtering means
 * per time step, one sampled path */
SEXP cpimh_R_pf_adaptive(SEXP cfg_raw, SEXP obs, SEXP nfilters, SEXP ess_threshold, SEXP systematic) {
  const cpimh_config* cfg = (const cpimh_config*)RAW(cfg_raw);
  int B = Rf_asInteger(nfilters), d = cfg->dim, T = cfg->nobs;
  SEXP ll = PROTECT(Rf_allocVector(REALSXP, B));
  SEXP means = PROTECT(Rf_alloc3DArray(REALSXP, d, T, B));                               /* pf_means: [B][T][d] = d x T x B column-major */
  SEXP path = PROTECT(Rf_alloc3DArray(REALSXP, d, T + 1, B));
  check(cpimh_pf_adaptive_batch(cfg, REAL(obs), B, seed_from_R(), 0, Rf_asReal(ess_threshold), Rf_asLogical(systematic), REAL(ll), NULL,
                                REAL(means), REAL(path), NULL, NULL));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3)); SET_VECTOR_ELT(out, 0, ll); SET_VECTOR_ELT(out, 1, means); SET_VECTOR_ELT(out, 2, path);
  UNPROTECT(4); return out;
}
SEXP cpimh_R_coupled_pimh(SEXP cfg_raw, SEXP obs, SEXP R_, SEXP k_, SEXP K_, SEXP maxit) {   /* coupled_pimh_chains + H_bar, R/debiasing_algorithm.R:140-316 */
  const cpimh_config* cfg = (const cpimh_config*)RAW(cfg_raw);
  int R = Rf_asInteger(R_), P = cfg->dim * (cfg->nobs + 1);
  SEXP hbar = PROTECT(Rf_allocMatrix(REALSXP, P, R));
  SEXP tau = PROTECT(Rf_allocVector(INTSXP, R)), it = PROTECT(Rf_allocVector(INTSXP, R));
  cpimh_chain_out o = {REAL(hbar), INTEGER(tau), INTEGER(it), NULL, NULL, NULL, NULL};
  check(cpimh_coupled_pimh(cfg, REAL(obs), R, seed_from_R(), 0, Rf_asInteger(k_), Rf_asInteger(K_), Rf_asInteger(maxit), &o));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3)); SET_VECTOR_ELT(out, 0, hbar); SET_VECTOR_ELT(out, 1, tau); SET_VECTOR_ELT(out, 2, it);
  UNPROTECT(4); return out;
}

/* CPF(ref_trajectory = x, with_as) / CPF_coupled(ref1, ref2, CR_indexmatching, with_as): R/CPF.R:22-24,45-58, R/CPF_coupled.R:14-112 */
SEXP cpimh_R_ccpf(SEXP cfg_raw, SEXP obs, SEXP ref1, SEXP ref2, SEXP with_as) {
  const cpimh_config* cfg = (const cpimh_config*)RAW(cfg_raw);
  int d = cfg->dim, T = cfg->nobs, nsys = Rf_isNull(ref2) ? 1 : 2;
  SEXP t1 = PROTECT(Rf_allocMatrix(REALSXP, d, T + 1)), t2 = PROTECT(Rf_allocMatrix(REALSXP, d, T + 1));
  SEXP lz = PROTECT(Rf_allocVector(REALSXP, 1));
  check(cpimh_ccpf_batch(cfg, REAL(obs), 1, seed_from_R(), 0, nsys, Rf_asLogical(with_as), REAL(ref1), nsys == 2 ? REAL(ref2) : NULL, NULL,
                         REAL(t1), REAL(t2), REAL(lz), NULL, NULL, NULL, NULL));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3)); SET_VECTOR_ELT(out, 0, t1); SET_VECTOR_ELT(out, 1, t2); SET_VECTOR_ELT(out, 2, lz);
  UNPROTECT(4); return out;
}
/* coupled_ccpf_chains + H_bar for R replicates (R/debiasing_algorithm.R:23-111); rb = TRUE also returns the Rao-Blackwellised H_bar */
SEXP cpimh_R_coupled_ccpf(SEXP cfg_raw, SEXP obs, SEXP R_, SEXP k_, SEXP K_, SEXP maxit, SEXP with_as, SEXP rb) {
  const cpimh_config* cfg = (const cpimh_config*)RAW(cfg_raw);
  int R = Rf_asInteger(R_), P = cfg->dim * (cfg->nobs + 1), want_rb = Rf_asLogical(rb);
  SEXP hbar = PROTECT(Rf_allocMatrix(REALSXP, P, R)), hrb = PROTECT(Rf_allocMatrix(REALSXP, P, want_rb ? R : 0));
  SEXP tau = PROTECT(Rf_allocVector(INTSXP, R)), it = PROTECT(Rf_allocVector(INTSXP, R));
  cpimh_chain_out o = {REAL(hbar), INTEGER(tau), INTEGER(it), NULL, NULL, NULL, NULL, want_rb ? REAL(hrb) : NULL, NULL, NULL};
  check(cpimh_coupled_ccpf(cfg, REAL(obs), R, seed_from_R(), 0, Rf_asInteger(k_), Rf_asInteger(K_), Rf_asInteger(maxit), Rf_asLogical(with_as), &o));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, hbar); SET_VECTOR_ELT(out, 1, tau); SET_VECTOR_ELT(out, 2, it); SET_VECTOR_ELT(out, 3, hrb);
  UNPROTECT(5); return out;
}

#define ENTRY(name, n) {#name, (DL_FUNC)&name, n}
static const R_CallMethodDef CallEntries[] = {
  ENTRY(_CoupledPIMH_systematic_resampling_n_, 3), ENTRY(_CoupledPIMH_multinomial_resampling_n_, 2),
  ENTRY(_CoupledPIMH_coupled_multinomial_resampling_n_, 3), ENTRY(_CoupledPIMH_indexmatching_cpp, 3),
  ENTRY(_CoupledPIMH_dmvnorm, 3), ENTRY(_CoupledPIMH_dmvnorm_transpose, 3), ENTRY(_CoupledPIMH_dmvnorm_transpose_cholesky, 3),
  ENTRY(_CoupledPIMH_rmvnorm, 3), ENTRY(_CoupledPIMH_rmvnorm_transpose, 3), ENTRY(_CoupledPIMH_rmvnorm_transpose_cholesky, 3),
  ENTRY(_CoupledPIMH_create_A_, 2), ENTRY(_CoupledPIMH_one_step_lorenz_vector, 5),
  ENTRY(_CoupledPIMH_rinitial_SVLevy_cpp, 5), ENTRY(_CoupledPIMH_rtransition_SVLevy_cpp, 6), ENTRY(_CoupledPIMH_dobs_SVLevy_cpp, 5),
  ENTRY(cpimh_R_tree_new, 3), ENTRY(cpimh_R_tree_init, 2), ENTRY(cpimh_R_tree_update, 3), ENTRY(cpimh_R_tree_get_path, 2),
  ENTRY(cpimh_R_tree_retrieve_xgeneration, 2), ENTRY(cpimh_R_tree_field, 2), ENTRY(cpimh_R_tree_reset, 1),
  ENTRY(cpimh_R_pf_batch, 4), ENTRY(cpimh_R_pf_adaptive, 5), ENTRY(cpimh_R_coupled_pimh, 6),
  ENTRY(cpimh_R_ccpf, 5), ENTRY(cpimh_R_coupled_ccpf, 8),
  {NULL, NULL, 0}};
void R_init_CoupledPIMH(DllInfo* dll) { R_registerRoutines(dll, NULL, CallEntries, NULL, NULL); R_useDynamicSymbols(dll, FALSE); }
