# cpimh_dispatch.R -- R-side changes that route the bootstrap-particle-filter hot path of CoupledPIMH to libcpimh.so through
# src/shim.c.  NOT executable in the build image (no R); see INTEGRATION.md.  Collate this file LAST (DESCRIPTION `Collate:`):
# it keeps the reference's functions under `*_reference` names and re-points the public names, so the package's own callers
# (get_pimh_kernels, coupled_pimh_chains, rheeglynn_estimator*, the scripts under inst/) reach the GPU without being edited.
# Every function keeps the reference's name and arguments (R/CPF.R, R/particlefilter.R, R/pf_kernels.R,
# R/debiasing_algorithm.R, R/resampling-*.R, R/coupledresampling-*.R).

cpimh_model_ids <- c(gss = 0L, ar = 1L, levy_sv = 2L, stochastic_kinetic = 3L, lorenz = 4L)

# ---- model lists: the reference's carry no name (R/model_get_gss.R:43-46); the constructors are re-pointed to add one
.with_name <- function(model, name, ...) { model$name <- name; model$cpimh <- list(...); model }
get_gss_reference <- get_gss;                               get_gss <- function() .with_name(get_gss_reference(), "gss")
get_ar_reference <- get_ar;                                 get_ar <- function(dimension) .with_name(get_ar_reference(dimension), "ar")
get_levy_sv_reference <- get_levy_sv;                       get_levy_sv <- function() .with_name(get_levy_sv_reference(), "levy_sv")
get_stochastic_kinetic_reference <- get_stochastic_kinetic; get_stochastic_kinetic <- function() .with_name(get_stochastic_kinetic_reference(), "stochastic_kinetic")
get_lorenz_reference <- get_lorenz;                         get_lorenz <- function(sigma_y) .with_name(get_lorenz_reference(sigma_y), "lorenz", sigma_y = sigma_y)

# struct cpimh_config (include/cpimh.h) as a raw vector: 10 int32, 32 doubles, 4 int32
cpimh_config <- function(model, theta, nparticles, nobs, shuffle = 0L, store_paths = 1L) {
  th <- switch(model$name,
               gss = numeric(0), ar = as.numeric(theta[1:2]), levy_sv = as.numeric(theta[1:5]),
               stochastic_kinetic = c(theta$s2_obs, as.numeric(theta$A_obs)),
               lorenz = c(as.numeric(theta[1:2]), model$cpimh$sigma_y))
  dy <- if (!is.null(model$dimension_y)) model$dimension_y else if (model$name %in% c("gss", "levy_sv")) 1L else model$dimension
  ints1 <- as.integer(c(cpimh_model_ids[[model$name]], model$dimension, dy, nparticles, nobs, shuffle,
                        1L, 1L, 64L, length(th)))   # integrator = dopri5 (reference), substeps, sk_max_events, ntheta
  con <- rawConnection(raw(0), "wb"); on.exit(close(con))
  writeBin(ints1, con, size = 4L); writeBin(c(th, numeric(32 - length(th))), con, size = 8L)
  writeBin(as.integer(c(0L, 0L, store_paths, 0L)), con, size = 4L)
  rawConnectionValue(con)
}
.on_gpu <- function(model) !is.null(model$name) && model$name %in% names(cpimh_model_ids)

# ---- CPF(nparticles, model, theta, observations, ref_trajectory = NULL, with_as = FALSE, all_particles = FALSE): R/CPF.R:14-78
CPF_reference <- CPF
CPF <- function(nparticles, model, theta, observations, ref_trajectory = NULL, with_as = FALSE, all_particles = FALSE) {
  if (!.on_gpu(model)) return(CPF_reference(nparticles, model, theta, observations, ref_trajectory, with_as, all_particles))
  cfg <- cpimh_config(model, theta, nparticles, nrow(observations))
  if (!is.null(ref_trajectory)) {                                                                 # conditional filter (R/CPF.R:22-24,45-58)
    if (all_particles) return(CPF_reference(nparticles, model, theta, observations, ref_trajectory, with_as, all_particles))
    res <- .Call("cpimh_R_ccpf", cfg, observations, ref_trajectory, NULL, with_as)
    return(list(trajectory = res[[1]], logz = res[[3]][1]))
  }
  res <- .Call("cpimh_R_pf_batch", cfg, observations, 1L, all_particles)
  out <- list(trajectory = matrix(res[[2]][, , 1], nrow = model$dimension), logz = res[[1]][1])
  if (all_particles) out$exp_path <- matrix(res[[3]][, , 1], nrow = model$dimension)              # R/CPF.R:73-76
  out
}

# ---- pf(y, theta, model, nparticles, ess_threshold = 1, systematic = FALSE): R/CPF.R:106-155 (particles are N x d there)
pf_reference <- pf
pf <- function(y, theta, model, nparticles, ess_threshold = 1, systematic = FALSE) {
  if (!.on_gpu(model)) return(pf_reference(y, theta, model, nparticles, ess_threshold, systematic))
  res <- .Call("cpimh_R_pf_adaptive", cpimh_config(model, theta, nparticles, nrow(y)), y, 1L, as.numeric(ess_threshold), systematic)
  list(pf_means = t(matrix(res[[2]][, , 1], nrow = model$dimension)), loglik = res[[1]][1], path = t(matrix(res[[3]][, , 1], nrow = model$dimension)))
}

# B proposals at once (what a batch of PIMH chains needs per sweep)
cpf_batch <- function(nparticles, model, theta, observations, nfilters, all_particles = FALSE) {
  res <- .Call("cpimh_R_pf_batch", cpimh_config(model, theta, nparticles, nrow(observations)), observations, as.integer(nfilters), all_particles)
  out <- list(logz = res[[1]], trajectories = res[[2]])
  if (all_particles) out$exp_paths <- res[[3]]
  out
}

# foreach(r = 1:R) %dorng% { c_chains <- coupled_pimh_chains(...); H_bar(c_chains, k = k, K = K) }  in ONE call
coupled_pimh_chains_batch <- function(model, theta, observations, nparticles, R, k = 0, K = 1, max_iterations = -1L) {
  res <- .Call("cpimh_R_coupled_pimh", cpimh_config(model, theta, nparticles, nrow(observations)), observations,
               as.integer(R), as.integer(k), as.integer(K), as.integer(max_iterations))
  P <- nrow(observations) + 1
  list(hbar = array(res[[1]], dim = c(model$dimension, P, R)), meetingtime = ifelse(res[[2]] < 0, Inf, res[[2]]), iteration = res[[3]],
       finished = res[[2]] >= 0)
}

# ---- CPF_coupled(nparticles, model, theta, observations, ref_trajectory1, ref_trajectory2, coupled_resampling, with_as): index matching is fused
CPF_coupled_reference <- CPF_coupled
CPF_coupled <- function(nparticles, model, theta, observations, ref_trajectory1, ref_trajectory2, coupled_resampling = NULL, with_as = FALSE) {
  if (!.on_gpu(model)) return(CPF_coupled_reference(nparticles, model, theta, observations, ref_trajectory1, ref_trajectory2, coupled_resampling, with_as))
  res <- .Call("cpimh_R_ccpf", cpimh_config(model, theta, nparticles, nrow(observations)), observations, ref_trajectory1, ref_trajectory2, with_as)
  list(new_trajectory1 = res[[1]], new_trajectory2 = res[[2]])
}
# foreach(r = 1:R) { coupled_ccpf_chains(get_cpf_kernels(with_as), pimh_init, N, K); H_bar(., k, K) }  in ONE call;
# k = K = 0 is rheeglynn_estimator (lambda = 0), and with rb = TRUE rheeglynn_estimator_RB
coupled_ccpf_chains_batch <- function(model, theta, observations, nparticles, R, k = 0, K = 1, max_iterations = -1L, with_as = FALSE, rb = FALSE) {
  res <- .Call("cpimh_R_coupled_ccpf", cpimh_config(model, theta, nparticles, nrow(observations)), observations,
               as.integer(R), as.integer(k), as.integer(K), as.integer(max_iterations), with_as, rb)
  P <- nrow(observations) + 1
  out <- list(hbar = array(res[[1]], dim = c(model$dimension, P, R)), meetingtime = ifelse(res[[2]] < 0, Inf, res[[2]]), iteration = res[[3]])
  if (rb) out$hbar_rb <- array(res[[4]], dim = c(model$dimension, P, R))
  out
}

# ---- class Tree (Rcpp module "module_tree", src/tree.cpp:144-162): the same 8 fields and methods as a reference class whose
# generator `new()` accepts, so `Tree <- new(TreeClass, nparticles, 10*nparticles*d, d)` (R/CPF.R:17, R/particlefilter.R:14,
# R/CPF_coupled.R:19-20, R/rheeglynn_estimator_RB.R:6,68-69) and `Tree$update(x, a - 1)`, `Tree$get_path(n)`, `Tree$a_star` ...
# keep working unchanged.  Fields are read-only views of the device arrays.
.tree_field <- function(which) eval(substitute(function(value) {
  if (!missing(value)) stop("Tree fields are read-only")
  .Call("cpimh_R_tree_field", ptr, W)
}, list(W = which)))
methods::setRefClass("CpimhTree",
  fields = list(ptr = "externalptr",
                N = .tree_field(0L), M = .tree_field(1L), dimx = .tree_field(2L), nsteps = .tree_field(3L),
                a_star = .tree_field(4L), o_star = .tree_field(5L), x_star = .tree_field(6L), l_star = .tree_field(7L)),
  methods = list(
    initialize = function(N, M, dimx, ...) {
      if (!missing(N)) ptr <<- .Call("cpimh_R_tree_new", as.integer(N), as.integer(M), as.integer(dimx))
      invisible(.self)
    },
    reset = function() invisible(.Call("cpimh_R_tree_reset", ptr)),
    init = function(x_0) invisible(.Call("cpimh_R_tree_init", ptr, x_0)),
    update = function(x, a) invisible(.Call("cpimh_R_tree_update", ptr, x, as.integer(a))),        # a is 0-based (R/CPF.R:65)
    get_path = function(n) .Call("cpimh_R_tree_get_path", ptr, as.integer(n)),
    retrieve_xgeneration = function(lag) .Call("cpimh_R_tree_retrieve_xgeneration", ptr, as.integer(lag))))
# `insert` and `prune` of the module are the two halves of update(); no R code of the package calls them separately
TreeClass <- methods::getClass("CpimhTree")      # new(TreeClass, N, M, dimx) dispatches to initialize() above

# referenced at R/pf_kernels.R:238 and R/CPF_coupled.R:80,102 but never defined in the package
CR_indexmatching <- function(xparticles1, xparticles2, normweights1, normweights2, ntrials = ncol(xparticles1), ...) {
  indexmatching_cpp(normweights1, normweights2, as.integer(ntrials)) + 1L
}
# the reference returns 0-based indices here (R/coupledresampling-multinomial.R:8-9)
CR_multinomial_reference <- CR_multinomial
CR_multinomial <- function(xparticles1, xparticles2, normweights1, normweights2, ...) {
  coupled_multinomial_resampling_n_(normweights1, normweights2, ncol(xparticles1)) + 1L
}
