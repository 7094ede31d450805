# cpimh_dispatch.R -- additive R-side changes that route the bootstrap-particle-filter hot path of CoupledPIMH to
# libcpimh.so through src/shim.c.  Not executable in the build image (no R); see INTEGRATION.md.
# Every function keeps the reference's name and arguments (R/CPF.R, R/particlefilter.R, R/pf_kernels.R,
# R/debiasing_algorithm.R, R/resampling-*.R, R/coupledresampling-*.R).

cpimh_model_ids <- c(gss = 0L, ar = 1L, levy_sv = 2L, stochastic_kinetic = 3L, lorenz = 4L)

# the reference's model lists carry no name (R/model_get_gss.R:43-46): wrap the constructors
.with_name <- function(model, name, ...) { model$name <- name; model$cpimh <- list(...); model }
get_gss_gpu <- function() .with_name(get_gss(), "gss")
get_ar_gpu <- function(dimension) .with_name(get_ar(dimension), "ar")
get_levy_sv_gpu <- function() .with_name(get_levy_sv(), "levy_sv")
get_stochastic_kinetic_gpu <- function() .with_name(get_stochastic_kinetic(), "stochastic_kinetic")
get_lorenz_gpu <- function(sigma_y) .with_name(get_lorenz(sigma_y), "lorenz", sigma_y = sigma_y)

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

# CPF(nparticles, model, theta, observations, ref_trajectory = NULL, with_as = FALSE, all_particles = FALSE)
CPF_gpu <- function(nparticles, model, theta, observations, ref_trajectory = NULL, with_as = FALSE, all_particles = FALSE) {
  if (all_particles || is.null(model$name))
    return(CPF(nparticles, model, theta, observations, ref_trajectory, with_as, all_particles))   # outside the GPU path
  if (!is.null(ref_trajectory)) {                                                                 # conditional filter (R/CPF.R:22-24,45-58)
    res <- .Call("cpimh_R_ccpf", cpimh_config(model, theta, nparticles, nrow(observations)), observations, ref_trajectory, NULL, with_as)
    return(list(trajectory = res[[1]], logz = res[[3]][1]))
  }
  res <- .Call("cpimh_R_pf_batch", cpimh_config(model, theta, nparticles, nrow(observations)), observations, 1L)
  list(trajectory = matrix(res[[2]][, , 1], nrow = model$dimension), logz = res[[1]][1])
}

# B proposals at once (what the PIMH kernels need)
cpf_batch <- function(nparticles, model, theta, observations, nfilters) {
  res <- .Call("cpimh_R_pf_batch", cpimh_config(model, theta, nparticles, nrow(observations)), observations, as.integer(nfilters))
  list(logz = res[[1]], trajectories = res[[2]])
}

# foreach(r = 1:R) %dorng% { c_chains <- coupled_pimh_chains(...); H_bar(c_chains, k = k, K = K) }  in ONE call
coupled_pimh_chains_batch <- function(model, theta, observations, nparticles, R, k = 0, K = 1, max_iterations = -1L) {
  res <- .Call("cpimh_R_coupled_pimh", cpimh_config(model, theta, nparticles, nrow(observations)), observations,
               as.integer(R), as.integer(k), as.integer(K), as.integer(max_iterations))
  P <- nrow(observations) + 1
  list(hbar = array(res[[1]], dim = c(model$dimension, P, R)), meetingtime = ifelse(res[[2]] < 0, Inf, res[[2]]), iteration = res[[3]])
}

# CPF_coupled(nparticles, model, theta, observations, ref_trajectory1, ref_trajectory2, coupled_resampling, with_as): index matching is fused
CPF_coupled_gpu <- function(nparticles, model, theta, observations, ref_trajectory1, ref_trajectory2, coupled_resampling = NULL, with_as = FALSE) {
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

# TreeClass <- module_tree$Tree  becomes:
TreeClass_gpu <- function(N, M, dimx) {
  p <- .Call("cpimh_R_tree_new", as.integer(N), as.integer(M), as.integer(dimx))
  list(init = function(x0) invisible(.Call("cpimh_R_tree_init", p, x0)),
       update = function(x, a) invisible(.Call("cpimh_R_tree_update", p, x, as.integer(a))),     # a is 0-based (R/CPF.R:65)
       get_path = function(n) .Call("cpimh_R_tree_get_path", p, as.integer(n)),
       retrieve_xgeneration = function(lag) .Call("cpimh_R_tree_retrieve_xgeneration", p, as.integer(lag)))
}

# referenced at R/pf_kernels.R:238 and R/CPF_coupled.R:80,102 but never defined in the package
CR_indexmatching <- function(xparticles1, xparticles2, normweights1, normweights2, ntrials = ncol(xparticles1), ...) {
  indexmatching_cpp(normweights1, normweights2, as.integer(ntrials)) + 1L
}
# the reference returns 0-based indices here (R/coupledresampling-multinomial.R:8-9)
CR_multinomial_fixed <- function(xparticles1, xparticles2, normweights1, normweights2, ...) {
  coupled_multinomial_resampling_n_(normweights1, normweights2, ncol(xparticles1)) + 1L
}
