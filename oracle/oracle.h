/*
 * oracle.h -- CPU restatement ("oracle") of the CoupledPIMH bootstrap-particle-filter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (unbiased_pimh_b200/, include/, the CUDA
 * library) may include, link or call anything in this directory.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY UNPINNED: the reference (an R package; /root/reference) ships no tests, golden
 * vectors or known-answer files for this path, and it cannot be built or run here (no R,
 * Rcpp, RcppEigen, BH/Boost, Eigen).  Each function below restates a reference routine and
 * cites the file:line it follows.  What pins it instead (tests/test_oracle_*.py):
 *   - the exact Kalman log-likelihood of the shipped LGSSM data set,
 *   - the authors' stored Var(log Z) tables and meeting-time frequencies (statistical goldens),
 *   - hand-checked small cases and brute-force re-derivations in numpy.
 *
 * Conventions: fp64 everywhere; R's long-double accumulation in sum()/mean()/colSums() is
 * reproduced with `long double`; particle matrices are R's column-major d x N (particle n's
 * d components contiguous); indices are 0-based at this boundary exactly as in the reference's
 * C++ natives (the +1 lives in the R wrappers).  All randomness is INJECTED (arrays) or drawn
 * from the Philox4x32-10 streams specified in DESIGN.md so that the CUDA library can be fed
 * the very same numbers.
 */
#ifndef CPIMH_ORACLE_H
#define CPIMH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- model ids / flags: numerically identical to include/cpimh.h (checked by tests) ---- */
enum { ORC_MODEL_GSS = 0, ORC_MODEL_AR = 1, ORC_MODEL_LEVY_SV = 2, ORC_MODEL_SK = 3, ORC_MODEL_LORENZ = 4 };
enum { ORC_INTEGRATOR_RK4 = 0, ORC_INTEGRATOR_DOPRI5 = 1 };
#define ORC_MAX_ITERATION 0xFFFFF   /* chain iterations use 20 bits of the Philox counter; this value is the stream of the second initial filter of coupled_ccpf_chains */
enum { ORC_P_RESAMPLE = 1, ORC_P_INIT = 2, ORC_P_TRANS = 3, ORC_P_PATH = 4, ORC_P_MH = 5, ORC_P_SHUFFLE = 6, ORC_P_COUPLE = 7, ORC_P_AS = 8 };

#define ORC_MAX_THETA 32
#define ORC_SK_MAX_EVENTS_DEFAULT 64

typedef struct orc_config {
  int model;        /* ORC_MODEL_* */
  int dim;          /* state dimension d */
  int dim_y;        /* observation dimension */
  int nparticles;   /* N */
  int nobs;         /* T */
  int shuffle;      /* 1: apply std::random_shuffle (src/multinomial.cpp:30) ; 0: skip */
  int integrator;   /* lorenz: ORC_INTEGRATOR_* */
  int substeps;     /* lorenz rk4 sub-steps per observation interval */
  int sk_max_events;/* injected-noise SK: events available per particle-step */
  int ntheta;
  double theta[ORC_MAX_THETA];
} orc_config;

/* injected randomness for one filter; any pointer may be NULL => drawn from Philox(seed, stream) */
typedef struct orc_noise {
  const double* init;      /* [n_init][N]              model init noise, component-major         */
  const double* trans;     /* [T][n_trans][N]          model transition noise                    */
  const double* resample;  /* [T][N]  ascending exp(lnMax_i) in (0,1) (post-recurrence)          */
  const double* shuffle;   /* [T][N-1] uniforms for randWrapper                                  */
  const double* path_u;    /* [1]     uniform for the final systematic(1) draw                   */
} orc_noise;

/* ---- Philox4x32-10 streams (DESIGN.md "RNG") ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_philox_uniform1_raw(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double* ua, uint32_t* z, uint32_t* w);
void orc_philox_uniform2(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double* ua, double* ub);

void orc_rng_transform(const double* u, int64_t n, double* lg, double* sn, double* cs);   /* rngmath.h, elementwise */

/* ---- resamplers ---- */
int  orc_systematic_resampling_n(const double* w, int nparticles, int ndraws, double u, int* anc);
void orc_sorted_uniforms(const double* u_raw, int n, double* e_sorted);
int  orc_multinomial_from_sorted(const double* w, int nparticles, int ndraws, const double* e_sorted, int scale_by_sumw, int* anc);
int  orc_multinomial_from_ladder(const double* w, int nparticles, int ndraws, const double* P, double total, int* anc);
void orc_random_shuffle_int(int* a, int n, const double* u_shuffle);
int  orc_multinomial_resampling_n(const double* w, int nparticles, int ndraws, const double* u_raw, const double* u_shuffle, int* anc);
int  orc_coupled_multinomial_resampling_n(const double* w1, const double* w2, int nparticles, int ndraws,
                                          const double* u_raw, const double* u_shuffle, int* anc /* ndraws x 2 col-major */);
int  orc_indexmatching(const double* w1, const double* w2, int nparticles, int ndraws, const double* u_couple,
                       const double* u_mult_raw, const double* u_mult_shuffle,
                       const double* u_cm_raw, const double* u_cm_shuffle, int* anc /* ndraws x 2 */);
void orc_normalize_weight(const double* logw, int n, double* w);
double orc_logz_increment(const double* logw, int n, double* normw);

/* ---- ancestor tree ---- */
typedef struct orc_tree orc_tree;
orc_tree* orc_tree_new(int N, int M, int dimx);
void orc_tree_free(orc_tree*);
void orc_tree_reset(orc_tree*);
void orc_tree_init(orc_tree*, const double* x0 /* d x N */);
void orc_tree_update(orc_tree*, const double* x, const int* a);
void orc_tree_prune(orc_tree*, const int* o);
void orc_tree_insert(orc_tree*, const double* x, const int* a);
void orc_tree_get_path(const orc_tree*, int n, double* path /* d x (nsteps+1) */);
void orc_tree_retrieve_xgeneration(const orc_tree*, int lag, double* xgen /* d x N */);
int  orc_tree_M(const orc_tree*);
int  orc_tree_nsteps(const orc_tree*);
int  orc_tree_live_nodes(const orc_tree*);
const int* orc_tree_a_star(const orc_tree*);
const int* orc_tree_o_star(const orc_tree*);
const int* orc_tree_l_star(const orc_tree*);
const double* orc_tree_x_star(const orc_tree*);

/* ---- Gaussian kernels (src/mvnorm.cpp) ---- */
int  orc_cholesky_lower(const double* cov, int d, double* L);
void orc_dmvnorm_transpose_cholesky(const double* x /* d x n */, int d, int n, const double* mean, const double* L, double* out);
int  orc_dmvnorm_transpose(const double* x, int d, int n, const double* mean, const double* cov, double* out);
int  orc_dmvnorm(const double* x /* n x d */, int n, int d, const double* mean, const double* cov, double* out);
void orc_rmvnorm_transpose_cholesky(int n, int d, const double* mean, const double* L, const double* z /* d x n */, double* out /* d x n */);
int  orc_rmvnorm_transpose(int n, int d, const double* mean, const double* cov, const double* z, double* out);
int  orc_rmvnorm(int n, int d, const double* mean, const double* cov, const double* z /* n x d col-major */, double* out /* n x d */);
void orc_create_A(double alpha, int d, double* A /* d x d col-major */);

/* ---- models ---- */
int  orc_model_n_init(const orc_config*);
int  orc_model_n_trans(const orc_config*);
void orc_model_rinit(const orc_config*, const double* noise /* [n_init][N] */, double* x /* d x N */);
int  orc_model_rtransition(const orc_config*, int time /* 1-based */, const double* noise /* [n_trans][N] */, double* x);
void orc_model_dmeasurement(const orc_config*, const double* x, const double* y /* dim_y */, double* logw);
void orc_fill_init_noise(const orc_config*, uint64_t seed, uint64_t filter_id, double* noise);
void orc_fill_trans_noise(const orc_config*, uint64_t seed, uint64_t filter_id, int time, double* noise);
void orc_lorenz_step(int d, double F, int integrator, int substeps, double t0, double t1, double dt, double* x /* d */);

/* ---- the filter: CPF(ref_trajectory=NULL) and particlefilter() ---- */
typedef struct orc_pf_out {
  double* trajectory;   /* d x (T+1) col-major, required                        */
  double* logz;         /* [1] sum of increments                                */
  double* logz_t;       /* [T] per-step increments (nullable)                   */
  double* exp_path;     /* d x (T+1) (nullable; all_particles=TRUE)             */
  double* all_paths;    /* d x (T+1) x N (nullable; particlefilter())           */
  double* normweights;  /* [N] final normalised weights (nullable)              */
  int*    ancestors_t;  /* [T][N] ancestors fed to Tree$update, 0-based (nullable) */
  int*    path_index;   /* [1] 0-based index drawn for the trajectory (nullable)*/
  int*    tree_M_final; /* [1] final tree capacity (nullable)                   */
  int*    clamp_count;  /* [1] out-of-range ancestors clamped (nullable) */
} orc_pf_out;

int orc_cpf(const orc_config* cfg, const double* obs /* T x dim_y col-major, NaN = NA */,
            const orc_noise* noise /* nullable */, uint64_t seed, uint64_t filter_id, orc_pf_out* out);
int orc_cpf_batch(const orc_config* cfg, const double* obs, uint64_t seed, uint64_t first_filter, int nfilters,
                  int nthreads, double* logz /* [B] */, double* traj /* [B] d x (T+1), nullable */);

/* ---- PIMH / coupled PIMH (R/pf_kernels.R:86-145, R/debiasing_algorithm.R:140-316) ---- */
typedef struct orc_chains {
  int iteration; int meetingtime; /* meetingtime = -1 if never met (R: Inf) */ int finished;
  int p;            /* T+1 */
  int nrows1;       /* iteration+1 */
  double* samples1; /* [nrows1][p] row-major */
  double* samples2; /* [iteration][p] */
  double* log_pdf1; /* [nrows1] */
  double* log_pdf2; /* [iteration] */
  int nfilters;     /* filters actually run */
  double* exp_samples1; /* [nrows1][p]   all_particles only, else NULL (debiasing_algorithm.R:168-171) */
  double* exp_samples2; /* [iteration][p] */
} orc_chains;
int  orc_coupled_pimh_chains(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t replicate,
                             int K, int max_iterations, int run_discarded_init, orc_chains* out);
int  orc_coupled_pimh_chains_ap(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t replicate,
                                int K, int max_iterations, int run_discarded_init, int all_particles, orc_chains* out);
void orc_chains_free(orc_chains*);
int  orc_H_bar(const orc_chains* c, int k, int K, double* hbar /* [p] */);
int  orc_H_bar_exp(const orc_chains* c, int k, int K, double* hbar /* [p] */);
int  orc_coupled_pimh_batch(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t first_replicate, int nrep,
                            int k, int K, int max_iterations, int nthreads,
                            double* hbar /* [nrep][p] */, int* meetingtime /* [nrep] */, int* iterations /* [nrep] */,
                            long long* nfilters_total);

/* ---- conditional filters: CPF(ref_trajectory, with_as) R/CPF.R:22-24,45-58 ; CPF_coupled R/CPF_coupled.R:14-112 ;
 *      get_cpf_kernels R/pf_kernels.R:236-265 ; coupled_ccpf_chains R/debiasing_algorithm.R:23-111 ---- */
typedef struct orc_ccpf_out {
  double* trajectory[2];  /* d x (T+1) each; [1] used when nsys == 2            */
  double* exp_path[2];    /* d x (T+1) weighted average of all N paths (nullable) */
  int*    ancestors[2];   /* [T][N] 0-based (nullable)                          */
  int     path_index[2];
  double  logz;           /* system 1's sum of log Z increments                 */
  int     clamp_count;
} orc_ccpf_out;
int  orc_model_has_dtransition(const orc_config*);
void orc_model_dtransition(const orc_config*, const double* next_x /* d */, const double* x /* d x N */, int time, double* out /* N */);
int  orc_ccpf(const orc_config* cfg, const double* obs, const orc_noise* noise /* init/trans only */, uint64_t seed, uint64_t filter_id,
              int nsys, const double* ref1, const double* ref2 /* d x (T+1) */, int with_as, orc_ccpf_out* out);
int  orc_coupled_ccpf_chains(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t replicate,
                             int K, int max_iterations, int with_as, orc_chains* out);
int  orc_coupled_ccpf_chains_rb(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t replicate,
                                int K, int max_iterations, int with_as, int rb, orc_chains* out);
int  orc_coupled_ccpf_batch(const orc_config* cfg, const double* obs, uint64_t seed, uint32_t first_replicate, int nrep,
                            int k, int K, int max_iterations, int with_as, int nthreads,
                            double* hbar /* [nrep][p] */, double* hbar_rb /* [nrep][p], nullable */, int* meetingtime, int* iterations,
                            long long* nfilters_total);

/* ---- pf(): ESS-adaptive bootstrap filter (R/CPF.R:106-155) ---- */
typedef struct orc_apf_out {
  double* path;        /* d x (T+1), required                          */
  double* pf_means;    /* [T][d] filtering means (nullable)             */
  double* loglik_t;    /* [T] cumulative log-likelihood (nullable)      */
  int*    resampled;   /* [T] 1 if the step resampled (nullable)        */
  int*    ancestors;   /* [T][N] 0-based (nullable)                     */
  double  loglik;
  int     path_index;
  int     clamp_count;
} orc_apf_out;
int orc_pf_adaptive(const orc_config* cfg, const double* obs, uint64_t seed, uint64_t filter_id, double ess_threshold, int systematic,
                    orc_apf_out* out);

const char* orc_version(void);

#ifdef __cplusplus
}
#endif
#endif
