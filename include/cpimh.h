/*
 * cpimh.h -- C ABI of libcpimh.so: the B200-native (sm_100a CUDA) bootstrap-particle-filter hot path
 * of the R package CoupledPIMH (particlemontecarlo/unbiased_pimh).
 *
 * This is the drop-in boundary: plain C, caller-allocated host (or, where noted, device) buffers,
 * int status returns, no R / torch / C++ types.  The R package reaches its natives through
 * `.Call("_CoupledPIMH_*")` thunks (reference src/RcppExports.cpp:9-380, registered at :346-380) and
 * the Rcpp module "module_tree" (src/tree.cpp:144-162); INTEGRATION.md shows the `.Call` shim a
 * maintainer adds so those names dispatch here.  Each entry point below cites the reference
 * interface it replaces.  All paths citations are relative to the reference repository root.
 *
 * Conventions
 *   - fp64 (R double) and int32 (R integer) only.  Ancestor indices are 0-based at this boundary,
 *     exactly like the reference's C++ natives; the +1 lives in the R wrappers
 *     (R/resampling-multinomial.R:4, R/resampling-systematic.R:9).
 *   - Matrices use R's column-major layout: a `d x N` particle matrix stores particle n's d
 *     components contiguously; `observations` is `T x dim_y` column-major; NA is NaN.
 *   - Every function returns 0 on success or a CPIMH_ERR_* code; cpimh_last_error() gives the
 *     message (thread-local).  CUDA errors are surfaced, never swallowed.  There is NO CPU
 *     fallback: without a usable sm_100 GPU the compute calls fail with CPIMH_ERR_CUDA.
 *   - Randomness: counter-based Philox4x32-10 streams keyed by (seed, filter id, time, particle,
 *     purpose) -- see DESIGN.md "RNG" -- or caller-injected arrays (cpimh_noise) for parity runs.
 */
#ifndef CPIMH_H
#define CPIMH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden; only this ABI is exported */
#endif

#define CPIMH_VERSION_MAJOR 0
#define CPIMH_VERSION_MINOR 1

/* ---- status codes ---- */
enum {
  CPIMH_OK = 0,
  CPIMH_ERR_INVALID = 1,        /* bad argument / unsupported configuration            */
  CPIMH_ERR_CUDA = 2,           /* CUDA runtime error (message has the CUDA string)    */
  CPIMH_ERR_TREE_OVERFLOW = 3,  /* ancestor tree capacity exhausted (Tree::double_size) */
  CPIMH_ERR_NOISE_OVERFLOW = 4, /* injected SK noise ran out of events                  */
  CPIMH_ERR_NOMEM = 5,
  CPIMH_ERR_NOT_MET = 6         /* coupled chains: max_iterations reached before the chains met (R: finished = FALSE) */
};

/* ---- models shipped by name (R/model_get_gss.R, model_get_ar.R, model_levy_sv.R,
 *      model_stochastic_kinetic.R, model_get_lorenz.R) ---- */
enum { CPIMH_MODEL_GSS = 0, CPIMH_MODEL_AR = 1, CPIMH_MODEL_LEVY_SV = 2, CPIMH_MODEL_SK = 3, CPIMH_MODEL_LORENZ = 4 };
enum { CPIMH_INTEGRATOR_RK4 = 0, CPIMH_INTEGRATOR_DOPRI5 = 1 };
/* Philox stream purposes (counter word 3, low byte) */
enum { CPIMH_P_RESAMPLE = 1, CPIMH_P_INIT = 2, CPIMH_P_TRANS = 3, CPIMH_P_PATH = 4, CPIMH_P_MH = 5, CPIMH_P_SHUFFLE = 6, CPIMH_P_COUPLE = 7, CPIMH_P_AS = 8 };

#define CPIMH_MAX_THETA 32
/* chain iterations live in 20 bits of the Philox counter: max_iterations < CPIMH_MAX_ITERATION; the value itself is the stream
 * of the second initial bootstrap filter of coupled_ccpf_chains (never a chain iteration) */
#define CPIMH_MAX_ITERATION 0xFFFFF

/* Model + filter configuration.  theta layout per model:
 *   GSS      : (none)
 *   AR       : theta[0]=alpha, theta[1]=sigma_y                       (R/model_get_ar.R:21,37-41)
 *   LEVY_SV  : theta[0..4] = mu, beta, xi, omega^2, lambda            (R/model_levy_sv.R:14)
 *   SK       : theta[0]=s2_obs, theta[1..8]=A_obs (2x4 column-major)  (R/model_stochastic_kinetic.R:147-148)
 *   LORENZ   : theta[0]=F, theta[1]=sigma_x, theta[2]=sigma_y         (R/model_get_lorenz.R:28-29,41) */
typedef struct cpimh_config {
  int32_t model;          /* CPIMH_MODEL_*                                                        */
  int32_t dim;            /* state dimension d (model$dimension)                                  */
  int32_t dim_y;          /* observation dimension                                                */
  int32_t nparticles;     /* N                                                                    */
  int32_t nobs;           /* T = nrow(observations)                                               */
  int32_t shuffle;        /* 1: std::random_shuffle after multinomial (src/multinomial.cpp:30);
                             0: skip (law-invariant for ref_trajectory=NULL; product default)     */
  int32_t integrator;     /* LORENZ: CPIMH_INTEGRATOR_*                                           */
  int32_t substeps;       /* LORENZ rk4 sub-steps per 0.05 interval                               */
  int32_t sk_max_events;  /* SK with injected noise: events available per particle-step           */
  int32_t ntheta;
  double  theta[CPIMH_MAX_THETA];
  /* ---- execution knobs (0 = automatic) ---- */
  int32_t tree_capacity;      /* tree mode: M nodes per filter; 0 => min(max(10*N*d, 3N), 16N)     */
  int32_t threads_per_filter; /* CTA size; 0 => chosen from N                                      */
  int32_t store_paths;        /* path storage: 0 none (log Z only), 1 auto, 2 dense particle history in
                                 HBM, 3 Jacob-Murray-Rubenthaler tree (src/tree.cpp); auto = dense when
                                 the histories of all resident CTAs fit in free HBM, else tree         */
  int32_t reserved;
} cpimh_config;

/* Injected randomness (host pointers; any may be NULL => Philox).  B = number of filters.
 * Layouts are component-major so that they are also the device layout:
 *   init     [B][n_init][N]       n_init : GSS 1, AR d, LORENZ d, LEVY_SV 3 (z0, sum_weighted_e, sum_e), SK 0
 *   trans    [B][T][n_trans][N]   n_trans: GSS 1, AR d, LORENZ d, LEVY_SV 2 (sum_weighted_e, sum_e),
 *                                          SK 2*sk_max_events (E_0..E_{M-1}, U_0..U_{M-1})
 *   resample [B][T][N]            ascending exp(lnMax_i) in (0,1): the uniforms AFTER the recurrence of
 *                                 src/multinomial.cpp:22-24 and BEFORE the multiplication by sumw
 *   shuffle  [B][T][N-1]          uniforms consumed by randWrapper (src/multinomial.cpp:6-10)
 *   path_u   [B]                  uniform of the final systematic_resampling_n(normweights, 1, u) (R/CPF.R:69) */
typedef struct cpimh_noise {
  const double* init;
  const double* trans;
  const double* resample;
  const double* shuffle;
  const double* path_u;
} cpimh_noise;

const char* cpimh_last_error(void);
int  cpimh_version(void);                         /* major*100 + minor */
void cpimh_config_default(cpimh_config* cfg);     /* zero + store_paths=1, substeps=1, sk_max_events=64 */
int  cpimh_model_n_init(const cpimh_config* cfg);
int  cpimh_model_n_trans(const cpimh_config* cfg);
int  cpimh_device_count(int* count);
int  cpimh_set_device(int device);

/* ============================== the filter ==============================
 * Batched CPF(nparticles, model, theta, observations, ref_trajectory=NULL) -- R/CPF.R:14-78 -- and
 * particlefilter() -- R/particlefilter.R:11-44.  One CTA owns one filter; filters form the grid.
 * Filter b uses Philox stream id (first_filter + b).
 *
 * Handle API (device-resident; used for HBM-resident timing and by the chain drivers): */
typedef struct cpimh_pf cpimh_pf;
int cpimh_pf_create(const cpimh_config* cfg, int64_t max_filters, cpimh_pf** out);
int cpimh_pf_destroy(cpimh_pf* h);
int cpimh_pf_set_observations(cpimh_pf* h, const double* obs_host /* T x dim_y col-major */);
int cpimh_pf_set_noise(cpimh_pf* h, int64_t nfilters, const cpimh_noise* noise_host /* NULL clears */);
/* asynchronous on `stream` (a cudaStream_t, NULL = default stream); results stay in HBM */
int cpimh_pf_run(cpimh_pf* h, int64_t nfilters, uint64_t seed, uint64_t first_filter, void* stream);
/* synchronises `stream`, checks per-filter status, copies requested outputs to host (NULL = skip):
 *   logz [B]; logz_t [B][T]; trajectory [B] x (d x (T+1) col-major); path_index [B] */
int cpimh_pf_fetch(cpimh_pf* h, int64_t nfilters, void* stream, double* logz, double* logz_t, double* trajectory, int32_t* path_index);
/* all_particles=TRUE outputs (R/CPF.R:73-76, R/particlefilter.R:39-43) for filter b of the last run:
 *   exp_path d x (T+1); trajectories d x (T+1) x N; normweights [N]   (any may be NULL) */
int cpimh_pf_fetch_all_particles(cpimh_pf* h, int64_t b, void* stream, double* exp_path, double* trajectories, double* normweights);
/* the same exp_path for EVERY filter of a batch, computed inside the filter kernel (what coupled_pimh_chains_all_particles
 * consumes, R/pf_kernels.R:205-230): switch on before cpimh_pf_run, then fetch [B] x (d x (T+1) col-major) */
int cpimh_pf_all_particles(cpimh_pf* h, int on);
int cpimh_pf_fetch_exp_paths(cpimh_pf* h, int64_t nfilters, void* stream, double* exp_path);
/* per-step ancestors [T][N] (0-based, as passed to Tree$update) of filter b; requires cpimh_pf_record_ancestors(h,1) before run */
int cpimh_pf_record_ancestors(cpimh_pf* h, int on);
int cpimh_pf_fetch_ancestors(cpimh_pf* h, int64_t b, void* stream, int32_t* ancestors);
/* device pointers of the resident outputs (valid until destroy): logz [B], trajectory [B][T+1][d] */
int cpimh_pf_device_outputs(cpimh_pf* h, double** d_logz, double** d_trajectory);
int cpimh_pf_device_exp_paths(cpimh_pf* h, double** d_exp_path);   /* [B][T+1][d]; after cpimh_pf_all_particles(h, 1) */
/* number of time steps each filter of the last run repeated with sequential-order sums because a sorted uniform came within
 * rounding distance of a CDF knot (DESIGN.md "Exact ancestors"): repeats [B] */
int cpimh_pf_fetch_exact_repeats(cpimh_pf* h, int64_t nfilters, void* stream, int32_t* repeats);
/* Wide filter (Lorenz 96, d = 32 / 64) only: the fused filter settles a near-knot step inside the kernel, one filter across the
 * GPU cannot (at N = 65536 a uniform comes within the worst-case rounding bound of a knot about once per step, and settling it
 * needs the reference's serial sum).  on = 1: every step uses cumsum(normweights) in the reference's order (src/multinomial.cpp:17),
 * ~12x slower at N = 65536; on = 0 (default): fast pass, cpimh_pf_fetch_exact_repeats then reports per filter whether any draw was that close. */
int cpimh_pf_exact_mode(cpimh_pf* h, int on);
/* Conditional filters (cpimh_ccpf_batch, cpimh_coupled_ccpf, cpimh_rheeglynn) and pf() (cpimh_pf_adaptive_batch) repeat a step in
 * the reference's operation order when a uniform comes within rounding distance of a CDF knot (or of the coupling probability);
 * this returns how many steps the last such call repeated (diagnostic). */
int64_t cpimh_last_exact_repeats(void);
int cpimh_pf_device_status(cpimh_pf* h, int32_t** d_status);   /* [B] per-filter status of the last run (0 = ok) */
/* kernels launched by this handle so far (bench.py's gpu_launches) and bytes of HBM it holds */
int64_t cpimh_pf_launch_count(const cpimh_pf* h);
int64_t cpimh_pf_device_bytes(const cpimh_pf* h);
int cpimh_pf_effective_config(const cpimh_pf* h, cpimh_config* out);

/* One-shot host API: H2D observations (+ injected noise), run, D2H results.  This is the call the
 * R shim's CPF()/particlefilter()/batched twins make (r-pkg/src/shim.c cpimh_R_pf_batch), once per PIMH sweep: the HBM
 * workspaces stay cached in the library between calls with the same configuration (they are 24 GB for BASELINE configs[1]);
 * cpimh_release_cached() frees them, and so does any allocation that would otherwise run out of memory. */
int cpimh_pf_batch(const cpimh_config* cfg, const double* obs_host, int64_t nfilters, uint64_t seed, uint64_t first_filter,
                   const cpimh_noise* noise_host, double* logz, double* logz_t, double* trajectory, int32_t* path_index);
/* CPF(..., all_particles = TRUE) for B filters (R/CPF.R:73-76, the *_all_particles PIMH kernels R/pf_kernels.R:147-223):
 * exp_path [B] x (d x (T+1)) is the weighted average of all N paths of each filter */
int cpimh_pf_batch_all_particles(const cpimh_config* cfg, const double* obs_host, int64_t nfilters, uint64_t seed, uint64_t first_filter,
                                 double* logz, double* trajectory, double* exp_path);
int cpimh_release_cached(void);

/* ============================== coupled PIMH ==============================
 * Batched coupled_pimh_chains(single_kernel, coupled_kernel, pimh_init, N, K) + H_bar(c_chains, h=identity on
 * state component 1, k, K) -- R/pf_kernels.R:86-145, R/debiasing_algorithm.R:140-265,284-316.
 * One persistent kernel runs all replicates (owner CTAs replay each chain in order, idle CTAs run proposals ahead for the
 * replicates still unfinished: csrc/chains.cuh), until iter >= max(tau, K) for every replicate.
 * Replicate r uses Philox stream id r for its filters (iteration in counter word 1) and its MH uniforms. */
typedef struct cpimh_chain_out {
  double*  hbar;         /* [R] x (d x (T+1) col-major)  H_bar_{k:K} per replicate, all state components (nullable) */
  int32_t* meetingtime;  /* [R]       tau, or -1 if max_iterations was hit first (R: Inf)          */
  int32_t* iteration;    /* [R]       iterations run = max(tau, K)                                 */
  double*  sum_hbar;     /* [d x (T+1)] sum over replicates (nullable)                             */
  double*  sum_hbar_sq;  /* [d x (T+1)] sum of squares (nullable)                                  */
  int64_t* nfilters;     /* [1]       particle filters actually run (nullable)                     */
  double*  bc;           /* [R] x (d x (T+1))  bias-correction term BC(c_chains, k, K) alone,
                            R/debiasing_algorithm.R:330-363 (nullable)                              */
  /* all_particles = TRUE chains (cpimh_chains_set_all_particles): H_bar over exp_samples1/2, the weighted average of
   * all N paths instead of the one drawn path (R/debiasing_algorithm.R:164-171,239-242; R/pf_kernels.R:147-215) */
  double*  hbar_rb;        /* [R] x (d x (T+1))  (nullable) */
  double*  sum_hbar_rb;    /* [d x (T+1)]        (nullable) */
  double*  sum_hbar_rb_sq; /* [d x (T+1)]        (nullable) */
  /* `finished` of the reference's c_chains list (R/debiasing_algorithm.R:262): 0 for a replicate that reached max_iterations
   * before its chains met.  Such replicates are left OUT of sum_hbar / sum_hbar_sq (H_bar refuses them in the reference);
   * nfinished is the number of replicates the sums hold. */
  int32_t* finished;       /* [R]                (nullable) */
  int64_t* nfinished;      /* [1]                (nullable) */
} cpimh_chain_out;
int cpimh_coupled_pimh(const cpimh_config* cfg, const double* obs_host, int64_t nreplicates, uint64_t seed, uint32_t first_replicate,
                       int k, int K, int max_iterations, cpimh_chain_out* out_host);
/* device-resident variant: `d_sums` (device, 2*d*(T+1)+2 doubles: sum_hbar, sum_hbar_sq, nfilters, nreplicates) is
 * what torch.distributed all-reduces across GPUs; per-replicate arrays stay in the handle. */
typedef struct cpimh_chains cpimh_chains;
int cpimh_chains_create(const cpimh_config* cfg, int64_t max_replicates, int record_samples, cpimh_chains** out);
int cpimh_chains_destroy(cpimh_chains* h);
int cpimh_chains_set_observations(cpimh_chains* h, const double* obs_host);
int cpimh_chains_run(cpimh_chains* h, int64_t nreplicates, uint64_t seed, uint32_t first_replicate, int k, int K, int max_iterations,
                     double* d_sums /* device, nullable */, void* stream);
int cpimh_chains_fetch(cpimh_chains* h, int64_t nreplicates, void* stream, cpimh_chain_out* out_host);
/* recorded chains of replicate r (requires record_samples): samples1 [(iter+1)][T+1], samples2 [iter][T+1],
 * log_pdf1 [iter+1], log_pdf2 [iter]  (R/debiasing_algorithm.R:254-263); rows beyond max_record are dropped */
int cpimh_chains_fetch_samples(cpimh_chains* h, int64_t r, void* stream, int max_rows, double* samples1, double* samples2,
                               double* log_pdf1, double* log_pdf2);
int64_t cpimh_chains_launch_count(const cpimh_chains* h);
/* execution strategy: 0 auto, 1 one persistent CTA per replicate, 2 iteration-synchronous rounds of batched (and, once few
 * replicates remain, speculative) proposals -- PIMH proposals do not depend on the chain state, so a long replicate's
 * proposals run in parallel.  Results are identical; auto = rounds unless samples are recorded. */
int cpimh_chains_set_mode(cpimh_chains* h, int mode);
/* get_pimh_kernels(with_as, all_particles = TRUE): every proposal also carries exp_path; fills the *_rb outputs.
 * Round-based driver only (no recorded samples). */
int cpimh_chains_set_all_particles(cpimh_chains* h, int on);
int64_t cpimh_chains_filters_launched(const cpimh_chains* h);   /* including discarded speculative proposals */

/* ============================== conditional particle filters ==============================
 * nsys = 1: CPF(nparticles, model, theta, observations, ref_trajectory = ref1, with_as)   -- R/CPF.R:14-78 (:22-24, :45-58)
 * nsys = 2: CPF_coupled(nparticles, model, theta, observations, ref1, ref2, CR_indexmatching, with_as) -- R/CPF_coupled.R:14-112
 * One CTA per filter; filter b uses Philox stream id first_filter + b for BOTH systems (common random numbers).  The last
 * particle follows the reference path, the other N-1 get sorted multinomial (nsys = 1) or index-matching (nsys = 2,
 * src/coupledresamplingindexmatching.cpp:8-57) ancestors; cfg.shuffle must be 0.  with_as (ancestor sampling) needs
 * model$dtransition: get_gss and get_ar.  ref*, traj*: [B] x (d x (T+1) col-major); logz [B] (system 1, R/CPF.R:71);
 * path_index [B][2]; ancestors [B][2][T][N] 0-based (nullable).  noise: init / trans arrays only (nullable). */
int cpimh_ccpf_batch(const cpimh_config* cfg, const double* obs_host, int64_t nfilters, uint64_t seed, uint64_t first_filter,
                     int nsys, int with_as, const double* ref1_host, const double* ref2_host, const cpimh_noise* noise,
                     double* traj1, double* traj2, double* logz, int32_t* path_index, int32_t* ancestors,
                     double* exp_path1, double* exp_path2 /* CPF_RB / CPF_RB_coupled estimates with test_function = identity
                                                             (R/rheeglynn_estimator_RB.R:2-167): [B] x (d x (T+1)), nullable */);
/* coupled_ccpf_chains(single_kernel, coupled_kernel, pimh_init, N, K, max_iterations) with get_cpf_kernels(with_as) and
 * pimh_init of get_pimh_kernels, + H_bar(c_chains, k, K) and BC -- R/debiasing_algorithm.R:23-111, R/pf_kernels.R:236-265,134-145.
 * R replicates advance together, one launch of conditional filters per iteration; replicate r uses stream id
 * first_replicate + r with the iteration index in the counter (0 and CPIMH_MAX_ITERATION: the two initial bootstrap filters).
 * Requesting hbar_rb / sum_hbar_rb* also carries every filter's weighted average of all N paths through the chains:
 * with k = K = 0 that is rheeglynn_estimator_RB, with k = K = m rheeglynn_estimator_RB2 (R/rheeglynn_estimator_RB.R:170-265);
 * k = K = 0 without it is rheeglynn_estimator with lambda = 0 (R/rheeglynn_estimator.R:14-66). */
int cpimh_coupled_ccpf(const cpimh_config* cfg, const double* obs_host, int64_t nreplicates, uint64_t seed, uint32_t first_replicate,
                       int k, int K, int max_iterations, int with_as, cpimh_chain_out* out_host);

/* rheeglynn_estimator(observations, model, theta, algoparameters) -- R/rheeglynn_estimator.R:14-66 -- for R replicates: X_0 +
 * sum_n (X_n - X~_{n-1}) / (1 - lambda)^n over coupled conditional filters until the references meet or, with lambda > 0, until
 * the replicate's geometric truncation variable truncation[r] (host, R draws of rgeom(1, lambda); 0 => the estimate is X_0).
 * lambda = 0: no truncation (truncation may be NULL).  out->hbar holds the estimates, out->iteration the iterations run. */
int cpimh_rheeglynn(const cpimh_config* cfg, const double* obs_host, int64_t nreplicates, uint64_t seed, uint32_t first_replicate,
                    double lambda, const int32_t* truncation, int max_iterations, int with_as, cpimh_chain_out* out_host);

/* ============================== pf(): ESS-adaptive bootstrap filter ==============================
 * pf(y, theta, model, nparticles, ess_threshold = 1, systematic = FALSE) -- R/CPF.R:106-155, B filters at once (the PMMH
 * likelihood evaluator): propagate, weigh with the carried log-weights, resample iff (1 / sum(w^2)) / N < ess_threshold by
 * multinomial or systematic resampling.  loglik [B] = sum of the increments; loglik_t [B][T] cumulative (nullable);
 * pf_means [B][T][d] filtering means (nullable); path [B] x (d x (T+1) col-major); resampled [B][T] flags (nullable);
 * ancestors [B][T][N] 0-based (nullable). */
int cpimh_pf_adaptive_batch(const cpimh_config* cfg, const double* obs_host, int64_t nfilters, uint64_t seed, uint64_t first_filter,
                            double ess_threshold, int systematic, double* loglik, double* loglik_t, double* pf_means, double* path,
                            int32_t* resampled, int32_t* ancestors);

/* the RNG specification's transcendental maps (DESIGN.md "RNG"; csrc/common.cuh rng_log / rng_sincos2pi), elementwise on
 * the device: lg = log(u), sn/cs = sin/cos(2 pi u) for u in (0,1).  Bit-identical to oracle/rngmath.h (tests). */
int cpimh_rng_transform(const double* u, int64_t n, double* lg, double* sn, double* cs);

/* ============================== primitives (batched over B independent problems) ==============================
 * Bit-exact with the reference natives given identical weights and uniforms.  All pointers are HOST pointers. */
/* systematic_resampling_n_(weights, ndraws, u)  src/systematic.cpp:7-32 ; w [B][N], u [B], anc [B][ndraws] */
int cpimh_systematic_resampling_n(const double* w, int64_t B, int N, int ndraws, const double* u, int32_t* anc);
/* multinomial_resampling_n_(weights, ndraws)  src/multinomial.cpp:13-32 with the sorted uniforms e [B][ndraws]
 * (ascending exp(lnMax_i)) and optional shuffle uniforms us [B][ndraws-1] injected; anc [B][ndraws] */
int cpimh_multinomial_resampling_n(const double* w, int64_t B, int N, int ndraws, const double* e_sorted, const double* u_shuffle, int32_t* anc);
/* coupled_multinomial_resampling_n_(w1, w2, ndraws)  src/multinomial.cpp:35-71 ; anc [B][2][ndraws] (ndraws x 2 col-major) */
int cpimh_coupled_multinomial_resampling_n(const double* w1, const double* w2, int64_t B, int N, int ndraws,
                                           const double* e_sorted, const double* u_shuffle, int32_t* anc);
/* indexmatching_cpp(w1, w2, ndraws)  src/coupledresamplingindexmatching.cpp:8-57 ; u_couple [B][N];
 * e_mult/us_mult sized [B][ndraws]/[B][ndraws-1] (first ncoupled used), e_cm/us_cm likewise ; anc [B][2][ndraws] */
int cpimh_indexmatching(const double* w1, const double* w2, int64_t B, int N, int ndraws, const double* u_couple,
                        const double* e_mult, const double* us_mult, const double* e_cm, const double* us_cm, int32_t* anc);
/* sorted uniforms from raw uniforms by the recurrence of src/multinomial.cpp:20-24 (parallel suffix scan) */
int cpimh_sorted_uniforms(const double* u_raw, int64_t B, int n, double* e_sorted);
/* normalize_weight(logw)  R/util_normalize_weight.R:5-9, plus log(mean(exp(logw-max)))+max (R/CPF.R:61-66) ; logz nullable */
int cpimh_normalize_weight(const double* logw, int64_t B, int N, double* w, double* logz);
/* dmvnorm / dmvnorm_transpose / dmvnorm_transpose_cholesky  src/mvnorm.cpp:66-132 */
int cpimh_dmvnorm_transpose_cholesky(const double* x /* d x n */, int d, int64_t n, const double* mean, const double* chol, double* out);
int cpimh_dmvnorm_transpose(const double* x /* d x n */, int d, int64_t n, const double* mean, const double* cov, double* out);
int cpimh_dmvnorm(const double* x /* n x d */, int64_t n, int d, const double* mean, const double* cov, double* out);
/* rmvnorm / rmvnorm_transpose / rmvnorm_transpose_cholesky  src/mvnorm.cpp:6-62 ; normals from Philox(seed) or injected z (nullable) */
int cpimh_rmvnorm_transpose_cholesky(int64_t n, int d, const double* mean, const double* chol, uint64_t seed, const double* z, double* out /* d x n */);
int cpimh_rmvnorm_transpose(int64_t n, int d, const double* mean, const double* cov, uint64_t seed, const double* z, double* out /* d x n */);
int cpimh_rmvnorm(int64_t n, int d, const double* mean, const double* cov, uint64_t seed, const double* z, double* out /* n x d */);
/* create_A_(alpha, d)  src/create_A_.cpp:6-13 (host-side precompute) */
int cpimh_create_A(double alpha, int d, double* A /* d x d */);
/* model pieces, one call each (R/model_*.R closures): x is d x N column-major, noise component-major [n][N] */
int cpimh_model_rinit(const cpimh_config* cfg, const double* noise, uint64_t seed, uint64_t filter_id, double* x);
int cpimh_model_rtransition(const cpimh_config* cfg, int time, const double* noise, uint64_t seed, uint64_t filter_id, double* x);
int cpimh_model_dmeasurement(const cpimh_config* cfg, const double* x, const double* y, double* logw);
/* one_step_lorenz_vector(xparticles, tstart, tend, h, parameters)  src/lorenz96.cpp:65-93 (any d) */
int cpimh_lorenz_transition(const double* x_in /* d x n */, int d, int64_t n, double tstart, double tend, double h, double F,
                            int integrator, int substeps, double* x_out);

/* ---- ancestor tree as a standalone object: class Tree, src/tree.h:6-39, src/tree.cpp:6-162 ---- */
typedef struct cpimh_tree cpimh_tree;
int cpimh_tree_create(int N, int M, int dimx, cpimh_tree** out);                 /* Tree(N, M, dimx)    :6-16   */
int cpimh_tree_destroy(cpimh_tree* t);
int cpimh_tree_reset(cpimh_tree* t);                                             /* reset()             :22-26  */
int cpimh_tree_init(cpimh_tree* t, const double* x0 /* d x N */);                /* init(x_0)           :28-34  */
int cpimh_tree_update(cpimh_tree* t, const double* x, const int32_t* a);         /* update(x, a)        :88-99  */
int cpimh_tree_get_path(cpimh_tree* t, int n, double* path /* d x (nsteps+1) */);/* get_path(n)         :118-129*/
int cpimh_tree_retrieve_xgeneration(cpimh_tree* t, int lag, double* xgen);       /* retrieve_xgeneration:131-141*/
int cpimh_tree_fields(cpimh_tree* t, int* N, int* M, int* dimx, int* nsteps);    /* fields              :146-150*/
int cpimh_tree_arrays(cpimh_tree* t, int32_t* a_star, int32_t* o_star, int32_t* l_star, double* x_star); /* :151-154 (nullable) */

/* FP64 FMA peak of the current device, measured with a register-resident DFMA kernel (TFLOP/s);
 * used by bench.py as the secondary (compute) bound for Lorenz-96. */
int cpimh_measure_fp64_peak(double* tflops);
/* measured peak warp-instruction issue rate of the device (independent integer / fp32 chains, 32 warps per SM), in 1e9 warp
 * instructions per second: the denominator of bench.py's roofline_secondary */
int cpimh_measure_issue_peak(double* gwarp_instr_per_s);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CPIMH_H */
