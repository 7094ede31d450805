"""PIMH kernels, coupled PIMH chains and the unbiased estimator H_bar -- the callers of the hot path.

Reference: get_pimh_kernels (R/pf_kernels.R:86-226), coupled_pimh_chains (R/debiasing_algorithm.R:140-265),
H_bar (R/debiasing_algorithm.R:284-316), rheeglynn_estimator (R/rheeglynn_estimator.R:14-66).

Two layers:
  * the R-shaped functions (closures, one replicate per call, lists of matrices) -- same names, arguments and return
    keys as the reference; every proposal is one CUDA filter;
  * the batched twins (`coupled_pimh_chains_batch`, `unbiased_estimator_sharded`) that run R replicates with one
    persistent CTA per replicate and, across GPUs, shard replicates over ranks with a single all-reduce of the
    estimator sums (this replaces the reference's `foreach(...) %dorng%` fork workers, which cannot share a CUDA context).
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L
from . import rng
from .filters import CPF, CPF_coupled


# ---------------------------------------------------------------------------------------------- R-shaped API
def get_pimh_kernels(with_as=False, all_particles=False, model=None, theta=None, observations=None):
    """get_pimh_kernels(with_as, all_particles = F).  In R the closures read `model`, `theta`, `observations` from the
    caller's global environment (R/pf_kernels.R:90); Python has no such lookup, so they are keyword arguments here.
    Returns {"single_kernel", "coupled_kernel", "pimh_init"}."""
    if model is None or observations is None:
        raise ValueError("pass model=, theta=, observations= (globals in the R version)")

    def _cpf(nparticles, all_p=False):
        return CPF(nparticles, model, theta, observations, with_as=False, all_particles=all_p)

    def single_kernel(chain_state, log_pdf_state, iter, nparticles):
        prop = _cpf(nparticles)
        u = float(rng.runif(1)[0])
        if math.log(u) < (prop["logz"] - log_pdf_state):
            chain_state, log_pdf_state = prop["trajectory"], prop["logz"]
        return {"chain_state": chain_state, "log_pdf_state": log_pdf_state}

    def coupled_kernel(chain_state1, chain_state2, log_pdf_state1, log_pdf_state2, iter, nparticles):
        prop = _cpf(nparticles)
        u = float(rng.runif(1)[0])
        if math.log(u) < (prop["logz"] - log_pdf_state1):
            chain_state1, log_pdf_state1 = prop["trajectory"], prop["logz"]
        if math.log(u) < (prop["logz"] - log_pdf_state2):
            chain_state2, log_pdf_state2 = prop["trajectory"], prop["logz"]
        return {"chain_state1": chain_state1, "chain_state2": chain_state2, "log_pdf_state1": log_pdf_state1, "log_pdf_state2": log_pdf_state2}

    def pimh_init(nparticles):
        r1 = _cpf(nparticles)
        r2 = _cpf(nparticles)     # run and discarded by coupled_pimh_chains, as in the reference (R/pf_kernels.R:140-142)
        return {"chain_state1": r1["trajectory"], "chain_state2": r2["trajectory"], "log_pdf_state1": r1["logz"], "log_pdf_state2": r2["logz"]}

    def single_kernel_all_particles(chain_state, log_pdf_state, exp_path, iter, nparticles):
        prop = _cpf(nparticles, True)
        u = float(rng.runif(1)[0])
        if math.log(u) < (prop["logz"] - log_pdf_state):
            chain_state, log_pdf_state, exp_path = prop["trajectory"], prop["logz"], prop["exp_path"]
        return {"chain_state": chain_state, "log_pdf_state": log_pdf_state, "exp_path": exp_path}

    def coupled_kernel_all_particles(chain_state1, chain_state2, log_pdf_state1, log_pdf_state2, exp_path1, exp_path2, iter, nparticles):
        prop = _cpf(nparticles, True)
        u = float(rng.runif(1)[0])
        if math.log(u) < (prop["logz"] - log_pdf_state1):
            chain_state1, log_pdf_state1, exp_path1 = prop["trajectory"], prop["logz"], prop["exp_path"]
        if math.log(u) < (prop["logz"] - log_pdf_state2):
            chain_state2, log_pdf_state2, exp_path2 = prop["trajectory"], prop["logz"], prop["exp_path"]
        return {"chain_state1": chain_state1, "chain_state2": chain_state2, "log_pdf_state1": log_pdf_state1, "log_pdf_state2": log_pdf_state2,
                "exp_path1": exp_path1, "exp_path2": exp_path2}

    def pimh_init_all_particles(nparticles):
        r1, r2 = _cpf(nparticles, True), _cpf(nparticles, True)
        return {"chain_state1": r1["trajectory"], "chain_state2": r2["trajectory"], "log_pdf_state1": r1["logz"], "log_pdf_state2": r2["logz"],
                "exp_path1": r1["exp_path"], "exp_path2": r2["exp_path"]}

    if all_particles:
        return {"single_kernel": single_kernel_all_particles, "coupled_kernel": coupled_kernel_all_particles, "pimh_init": pimh_init_all_particles}
    return {"single_kernel": single_kernel, "coupled_kernel": coupled_kernel, "pimh_init": pimh_init}


def coupled_pimh_chains(single_kernel, coupled_kernel, pimh_init, N, K=1, max_iterations=math.inf, preallocate=10, verbose=False):
    """coupled_pimh_chains(single_kernel, coupled_kernel, pimh_init, N, K = 1, max_iterations = Inf, preallocate = 10)
    -- R/debiasing_algorithm.R:140-265, statement for statement (the growing preallocated matrices become lists)."""
    init = pimh_init(N)
    chain_state1, log_pdf_state1 = init["chain_state1"], init["log_pdf_state1"]
    all_particles = init.get("exp_path1") is not None
    samples1, samples2, log_pdf1, log_pdf2 = [np.array(chain_state1[0, :])], [], [log_pdf_state1], []
    chain_state2, log_pdf_state2 = None, -math.inf                      # :160-161
    exp_path1 = init.get("exp_path1")
    exp_path2 = None
    exp_samples1, exp_samples2 = ([np.array(exp_path1[0, :])], []) if all_particles else (None, None)
    it, meet, finished, meetingtime = 0, False, False, math.inf
    while not finished and it < max_iterations:
        if verbose:
            print(it)
        it += 1
        if meet:
            mh = single_kernel(chain_state1, log_pdf_state1, exp_path1, it, N) if all_particles else single_kernel(chain_state1, log_pdf_state1, it, N)
            chain_state1 = chain_state2 = mh["chain_state"]
            log_pdf_state1 = log_pdf_state2 = mh["log_pdf_state"]
            if all_particles:
                exp_path1 = exp_path2 = mh["exp_path"]
        else:
            if all_particles:
                mh = coupled_kernel(chain_state1, chain_state2, log_pdf_state1, log_pdf_state2, exp_path1, exp_path2, it, N)
                exp_path1, exp_path2 = mh["exp_path1"], mh["exp_path2"]
            else:
                mh = coupled_kernel(chain_state1, chain_state2, log_pdf_state1, log_pdf_state2, it, N)
            chain_state1, chain_state2 = mh["chain_state1"], mh["chain_state2"]
            log_pdf_state1, log_pdf_state2 = mh["log_pdf_state1"], mh["log_pdf_state2"]
            if log_pdf_state1 == log_pdf_state2 and chain_state2 is not None and np.array_equal(chain_state1, chain_state2) and not meet:
                meet, meetingtime = True, it                                # :216-220
        samples1.append(np.array(chain_state1[0, :]))                       # :235-236 first state component only
        samples2.append(np.array(chain_state2[0, :]))
        log_pdf1.append(log_pdf_state1)
        log_pdf2.append(log_pdf_state2)
        if all_particles:
            exp_samples1.append(np.array(exp_path1[0, :]))
            exp_samples2.append(np.array(exp_path2[0, :]))
        if it >= max(meetingtime, K):                                       # :245
            finished = True
    out = {"samples1": np.array(samples1), "samples2": np.array(samples2).reshape(it, -1), "log_pdf1": np.array(log_pdf1).reshape(-1, 1),
           "log_pdf2": np.array(log_pdf2).reshape(-1, 1), "meetingtime": meetingtime, "iteration": it, "finished": finished}
    if all_particles:
        out["exp_samples1"], out["exp_samples2"] = np.array(exp_samples1), np.array(exp_samples2).reshape(it, -1)
    return out


def H_bar(c_chains, h=lambda x: x, k=0, K=1):
    """H_bar(c_chains, h = function(x) x, k = 0, K = 1) -- R/debiasing_algorithm.R:284-316."""
    maxiter = c_chains["iteration"]
    if k > maxiter:
        print("error: k has to be less than the horizon of the coupled chains")
        return None
    if K > maxiter:
        print("error: K has to be less than the horizon of the coupled chains")
        return None
    s1, s2 = c_chains["samples1"], c_chains["samples2"]
    hb = np.sum(np.array([np.atleast_1d(h(s1[t])) for t in range(k, K + 1)], dtype=np.float64), axis=0)
    tau = c_chains["meetingtime"]
    if not (tau <= k + 1):
        dterm = np.zeros_like(hb)
        for t in range(k, int(min(maxiter - 1, tau - 1)) + 1):
            coef = min(t - k + 1, K - k + 1)
            dterm = dterm + coef * (np.atleast_1d(h(s1[t + 1])) - np.atleast_1d(h(s2[t])))
        hb = hb + dterm
    return hb / (K - k + 1)


def BC(c_chains, h=lambda x: x, k=0, K=1):
    """BC(c_chains, h = function(x) x, k = 0, K = 1) -- the bias-correction term of H_bar alone
    (R/debiasing_algorithm.R:330-363); K = Inf gives the un-normalised almost-sure limit."""
    maxiter = c_chains["iteration"]
    if k > maxiter:
        print("error: k has to be less than the horizon of the coupled chains")
        return None
    if K > maxiter and math.isfinite(K):
        print("error: K has to be less than the horizon of the coupled chains")
        return None
    s1, s2 = c_chains["samples1"], c_chains["samples2"]
    dterm = np.zeros_like(np.atleast_1d(h(s1[0])), dtype=np.float64)
    tau = c_chains["meetingtime"]
    if not (tau <= k + 1):
        for t in range(k, int(min(maxiter - 1, tau - 1)) + 1):
            coef = min(t - k + 1, K - k + 1)
            dterm = dterm + coef * (np.atleast_1d(h(s1[t + 1])) - np.atleast_1d(h(s2[t])))
    return dterm if not math.isfinite(K) else dterm / (K - k + 1)


# ---------------------------------------------------------------------------------------------- conditional-filter chains
def get_cpf_kernels(with_as=False, model=None, theta=None, observations=None):
    """get_cpf_kernels(with_as) -- R/pf_kernels.R:236-265.  As for get_pimh_kernels, the R closures capture `model`, `theta`
    and `observations` from the calling environment; here they are passed.  Returns single_kernel / coupled_kernel."""
    def single_kernel(chain_state, log_pdf_state, iter, nparticles):
        r = CPF(nparticles, model, theta, observations, ref_trajectory=chain_state, with_as=with_as)
        return {"chain_state": r["trajectory"], "log_pdf_state": r["logz"]}

    def coupled_kernel(chain_state1, chain_state2, log_pdf_state1, log_pdf_state2, iter, nparticles):
        r = CPF_coupled(nparticles, model, theta, observations, chain_state1, chain_state2, with_as=with_as)
        return {"chain_state1": r["new_trajectory1"], "chain_state2": r["new_trajectory2"],
                "log_pdf_state1": log_pdf_state1, "log_pdf_state2": log_pdf_state2}

    return {"single_kernel": single_kernel, "coupled_kernel": coupled_kernel}


def coupled_ccpf_chains(single_kernel, coupled_kernel, pimh_init, N, K=1, max_iterations=math.inf, preallocate=10, verbose=False):
    """coupled_ccpf_chains(single_kernel, coupled_kernel, pimh_init, N, K, max_iterations) -- R/debiasing_algorithm.R:23-111.
    Chain 1 is moved once by the single kernel, then both chains by the coupled kernel until they meet (X_t == Y_{t-1})."""
    init = pimh_init(N)
    chain_state1, chain_state2 = init["chain_state1"], init["chain_state2"]
    log_pdf_state1, log_pdf_state2 = init["log_pdf_state1"], init["log_pdf_state2"]
    samples1, samples2 = [np.array(chain_state1[0, :])], [np.array(chain_state2[0, :])]
    log_pdf1, log_pdf2 = [log_pdf_state1], [log_pdf_state2]
    it = 1
    mh = single_kernel(chain_state1, log_pdf_state1, it, N)
    chain_state1, log_pdf_state1 = mh["chain_state"], mh["log_pdf_state"]
    samples1.append(np.array(chain_state1[0, :])); log_pdf1.append(log_pdf_state1)
    meet, finished, meetingtime = False, False, math.inf
    while not finished and it < max_iterations:
        it += 1
        if meet:
            mh = single_kernel(chain_state1, log_pdf_state1, it, N)
            chain_state1 = chain_state2 = mh["chain_state"]
            log_pdf_state1 = log_pdf_state2 = mh["log_pdf_state"]
        else:
            mh = coupled_kernel(chain_state1, chain_state2, log_pdf_state1, log_pdf_state2, it, N)
            chain_state1, chain_state2 = mh["chain_state1"], mh["chain_state2"]
            log_pdf_state1, log_pdf_state2 = mh["log_pdf_state1"], mh["log_pdf_state2"]
            if np.array_equal(chain_state1, chain_state2) and not meet:
                meet, meetingtime = True, it
        samples1.append(np.array(chain_state1[0, :])); samples2.append(np.array(chain_state2[0, :]))
        log_pdf1.append(log_pdf_state1); log_pdf2.append(log_pdf_state2)
        if it >= max(meetingtime, K):
            finished = True
    # samples2 holds iteration rows: the initial state and one row per loop iteration but the last is aligned as in R
    return {"samples1": np.array(samples1), "samples2": np.array(samples2[:it]), "log_pdf1": np.array(log_pdf1).reshape(-1, 1),
            "log_pdf2": np.array(log_pdf2[:it]).reshape(-1, 1), "meetingtime": meetingtime, "iteration": it, "finished": finished}


def coupled_ccpf_chains_batch(model, theta, observations, nparticles, nreplicates, K=1, k=0, max_iterations=math.inf, with_as=False,
                              seed=None, first_replicate=0, rb=False, **kw):
    """R replicates of coupled_ccpf_chains(get_cpf_kernels(with_as), pimh_init, N = nparticles, K) + H_bar(., k, K) + BC on the
    GPU (cpimh_coupled_ccpf): all replicates advance together, one launch of conditional filters per iteration.
    Returns hbar / bc (R, d, T+1), meetingtime (R,) [-1: not met within max_iterations], iteration (R,), sums, nfilters.
    rb=True adds hbar_rb: the same estimator over every filter's weighted average of all N paths (Rao-Blackwellised,
    R/rheeglynn_estimator_RB.R:170-265: k = K = 0 is rheeglynn_estimator_RB, k = K = m is rheeglynn_estimator_RB2)."""
    obs = np.asarray(observations, dtype=np.float64)
    T = obs.shape[0]
    cfg = model.config(nparticles, theta, nobs=T, **kw)
    ob = L.obs_matrix(obs, T)
    R, d, P = int(nreplicates), model.dimension, T + 1
    seed = rng.next_seed() if seed is None else int(seed)
    mi = -1 if (max_iterations is None or max_iterations == math.inf or max_iterations < 0) else int(max_iterations)
    hb, bcv = np.empty((R, P, d)), np.empty((R, P, d))
    mt, it = np.empty(R, dtype=np.int32), np.empty(R, dtype=np.int32)
    s, s2, nf = np.empty((P, d)), np.empty((P, d)), np.zeros(1, dtype=np.int64)
    o = L.ChainOut()
    o.hbar, o.bc, o.meetingtime, o.iteration = hb.ctypes.data, bcv.ctypes.data, mt.ctypes.data, it.ctypes.data
    o.sum_hbar, o.sum_hbar_sq, o.nfilters = s.ctypes.data, s2.ctypes.data, nf.ctypes.data
    hrb, srb, srb2 = (np.empty((R, P, d)), np.empty((P, d)), np.empty((P, d))) if rb else (None, None, None)
    if rb:
        o.hbar_rb, o.sum_hbar_rb, o.sum_hbar_rb_sq = hrb.ctypes.data, srb.ctypes.data, srb2.ctypes.data
    L.check(L.lib().cpimh_coupled_ccpf(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_int64(R), C.c_uint64(seed), C.c_uint32(int(first_replicate)),
                                        C.c_int(int(k)), C.c_int(int(K)), C.c_int(mi), C.c_int(int(bool(with_as))), C.byref(o)))
    out = {"hbar": hb.transpose(0, 2, 1), "bc": bcv.transpose(0, 2, 1), "meetingtime": mt, "iteration": it,
           "sum_hbar": s.T.copy(), "sum_hbar_sq": s2.T.copy(), "nfilters": int(nf[0])}
    if rb:
        out.update({"hbar_rb": hrb.transpose(0, 2, 1), "sum_hbar_rb": srb.T.copy(), "sum_hbar_rb_sq": srb2.T.copy()})
    return out


# ---------------------------------------------------------------------------------------------- batched GPU twins
class ChainBatch:
    """Device-resident coupled-PIMH driver (cpimh_chains handle): one persistent CTA per replicate."""

    def __init__(self, model, theta, observations, nparticles, max_replicates, record_samples=0, tree_capacity=0,
                 threads_per_filter=0, sk_max_events=64, device=None, mode=0, store_paths=1, all_particles=False):
        obs = np.asarray(observations, dtype=np.float64)
        self.nobs = obs.shape[0]
        self.model = model
        self.all_particles = bool(all_particles)
        self.cfg = model.config(nparticles, theta, nobs=self.nobs, tree_capacity=tree_capacity, threads_per_filter=threads_per_filter,
                                sk_max_events=sk_max_events, store_paths=store_paths)
        if device is not None:
            L.check(L.lib().cpimh_set_device(C.c_int(int(device))))
        self.h = C.c_void_p()
        self.record = int(record_samples)
        L.check(L.lib().cpimh_chains_create(C.byref(self.cfg), C.c_int64(int(max_replicates)), C.c_int(self.record), C.byref(self.h)))
        ob = L.obs_matrix(obs, self.nobs)
        L.check(L.lib().cpimh_chains_set_observations(self.h, ob.ctypes.data_as(C.c_void_p)))
        if mode:
            L.check(L.lib().cpimh_chains_set_mode(self.h, C.c_int(int(mode))))
        if self.all_particles:
            L.check(L.lib().cpimh_chains_set_all_particles(self.h, C.c_int(1)))
        self.last = 0

    def close(self):
        if getattr(self, "h", None):
            L.lib().cpimh_chains_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def nsums(self):
        return 2 * self.cfg.dim * (self.cfg.nobs + 1) + 2

    def run(self, nreplicates, seed, first_replicate=0, k=0, K=1, max_iterations=-1, d_sums=None, stream=None):
        mi = -1 if (max_iterations is None or max_iterations == math.inf or max_iterations < 0) else int(max_iterations)
        L.check(L.lib().cpimh_chains_run(self.h, C.c_int64(int(nreplicates)), C.c_uint64(int(seed)), C.c_uint32(int(first_replicate)),
                                         C.c_int(int(k)), C.c_int(int(K)), C.c_int(mi), C.c_void_p(d_sums or 0), C.c_void_p(stream or 0)))
        self.last = int(nreplicates)

    def fetch(self, nreplicates=None, stream=None, hbar=True, bc=False):
        R = self.last if nreplicates is None else int(nreplicates)
        d, P = self.cfg.dim, self.cfg.nobs + 1
        hb = np.empty((R, P, d)) if hbar else None
        mt = np.empty(R, dtype=np.int32)
        it = np.empty(R, dtype=np.int32)
        s = np.empty((P, d))
        s2 = np.empty((P, d))
        nf = np.zeros(1, dtype=np.int64)
        o = L.ChainOut()
        o.hbar = hb.ctypes.data if hbar else None
        bcv = np.empty((R, P, d)) if bc else None
        o.bc = bcv.ctypes.data if bc else None
        o.meetingtime, o.iteration = mt.ctypes.data, it.ctypes.data
        o.sum_hbar, o.sum_hbar_sq, o.nfilters = s.ctypes.data, s2.ctypes.data, nf.ctypes.data
        fin = np.empty(R, dtype=np.int32)
        nfin = np.zeros(1, dtype=np.int64)
        o.finished, o.nfinished = fin.ctypes.data, nfin.ctypes.data
        rb = self.all_particles
        hrb = np.empty((R, P, d)) if (rb and hbar) else None
        srb, srb2 = (np.empty((P, d)), np.empty((P, d))) if rb else (None, None)
        if rb:
            o.hbar_rb = hrb.ctypes.data if hbar else None
            o.sum_hbar_rb, o.sum_hbar_rb_sq = srb.ctypes.data, srb2.ctypes.data
        L.check(L.lib().cpimh_chains_fetch(self.h, C.c_int64(R), C.c_void_p(stream or 0), C.byref(o)))
        # finished[r] is False when max_iterations was reached before the chains met (c_chains$finished, R/debiasing_algorithm.R:262);
        # the sums hold the finished replicates only
        out = {"meetingtime": mt, "iteration": it, "sum_hbar": s.T.copy(), "sum_hbar_sq": s2.T.copy(), "nfilters": int(nf[0]),
               "finished": fin.astype(bool), "nfinished": int(nfin[0])}
        if hbar:
            out["hbar"] = hb.transpose(0, 2, 1)       # (R, d, T+1)
        if bc:
            out["bc"] = bcv.transpose(0, 2, 1)
        if rb:
            out["sum_hbar_rb"], out["sum_hbar_rb_sq"] = srb.T.copy(), srb2.T.copy()
            if hbar:
                out["hbar_rb"] = hrb.transpose(0, 2, 1)
        return out

    def fetch_chains(self, r, stream=None):
        """c_chains list of replicate r in the reference's format (R/debiasing_algorithm.R:254-263)."""
        P, rows = self.cfg.nobs + 1, self.record
        s1 = np.empty((rows + 1, P)); s2 = np.empty((rows, P)); l1 = np.empty(rows + 1); l2 = np.empty(rows)
        L.check(L.lib().cpimh_chains_fetch_samples(self.h, C.c_int64(int(r)), C.c_void_p(stream or 0), C.c_int(rows), L.ptr(s1), L.ptr(s2), L.ptr(l1), L.ptr(l2)))
        return s1, s2, l1, l2

    @property
    def launch_count(self):
        return int(L.lib().cpimh_chains_launch_count(self.h))

    @property
    def filters_launched(self):
        """filters launched by the last run of the round-based driver, including discarded speculative proposals (-1: persistent kernel)"""
        return int(L.lib().cpimh_chains_filters_launched(self.h))


def coupled_pimh_chains_batch(model, theta, observations, nparticles, nreplicates, K=1, k=0, max_iterations=math.inf, seed=None,
                              first_replicate=0, record_samples=0, mode=0, **kw):
    """R replicates of coupled_pimh_chains(..., N = nparticles, K) + H_bar(., k = k, K = K) on the GPU.
    Returns hbar (R, d, T+1), meetingtime (R,) [-1 = not met within max_iterations], iteration (R,), the sums that are
    all-reduced across GPUs, the number of filters run, and (record_samples > 0) the c_chains lists.
    all_particles=True (get_pimh_kernels(with_as, all_particles = TRUE), R/pf_kernels.R:147-222) adds hbar_rb: H_bar over
    exp_samples1/2, the weighted average of all N paths of each accepted filter."""
    seed = rng.next_seed() if seed is None else int(seed)
    cb = ChainBatch(model, theta, observations, nparticles, nreplicates, record_samples=record_samples, mode=mode, **kw)
    try:
        cb.run(nreplicates, seed, first_replicate, k, K, max_iterations)
        out = cb.fetch(bc=True)
        if record_samples:
            chains = []
            for r in range(int(nreplicates)):
                s1, s2, l1, l2 = cb.fetch_chains(r)
                it = int(out["iteration"][r])
                n = min(it, record_samples)
                tau = int(out["meetingtime"][r])
                chains.append({"samples1": s1[:n + 1], "samples2": s2[:n], "log_pdf1": l1[:n + 1].reshape(-1, 1), "log_pdf2": l2[:n].reshape(-1, 1),
                               "meetingtime": (math.inf if tau < 0 else tau), "iteration": it, "finished": tau > 0 and it >= max(tau, K)})
            out["c_chains"] = chains
    finally:
        cb.close()
    return out


def rheeglynn_estimator(observations, model, theta, algoparameters, force_niterations=0, max_niterations=0, seed=None, replicate=0):
    """rheeglynn_estimator(observations, model, theta, algoparameters, force_niterations = 0, max_niterations = 0)
    -- R/rheeglynn_estimator.R:14-66: X_0 + sum_{n >= 1} (X_n - X~_{n-1}) over coupled conditional particle filters until the two
    reference trajectories meet, i.e. H_bar with k = K = 0 over coupled_ccpf_chains.  algoparameters: nparticles, with_as,
    coupled_resampling (index matching is what is fused), lambda (geometric truncation).  The shipped R function
    adds the lists returned by CPF(); this is the estimator it intends, for all state components."""
    lam = float(algoparameters.get("lambda", 0) or 0)
    if lam > 0:
        # truncation <- rgeom(1, lambda) (:27): failures before the first success; difference n is divided by (1 - lambda)^n (:58)
        seed = rng.next_seed() if seed is None else int(seed)
        trunc = int(np.random.default_rng([seed, int(replicate), 0x7267]).geometric(lam)) - 1
        if force_niterations:
            raise NotImplementedError("force_niterations with lambda > 0")
        obs = np.asarray(observations, dtype=np.float64)
        T, d = obs.shape[0], model.dimension
        cfg = model.config(int(algoparameters["nparticles"]), theta, nobs=T)
        hb, it, mt = np.empty((1, T + 1, d)), np.empty(1, dtype=np.int32), np.empty(1, dtype=np.int32)
        o = L.ChainOut()
        o.hbar, o.iteration, o.meetingtime = hb.ctypes.data, it.ctypes.data, mt.ctypes.data
        tr = np.array([trunc], dtype=np.int32)
        L.check(L.lib().cpimh_rheeglynn(C.byref(cfg), L.obs_matrix(obs, T).ctypes.data_as(C.c_void_p), C.c_int64(1), C.c_uint64(seed), C.c_uint32(int(replicate)),
                                        C.c_double(lam), L.ptr(tr), C.c_int(int(max_niterations) if max_niterations else -1),
                                        C.c_int(int(bool(algoparameters.get("with_as", False)))), C.byref(o)))
        return {"estimate": hb[0].T.copy(), "iteration": int(it[0]) if trunc > 0 else 0, "truncation": trunc}
    mi = int(force_niterations) if force_niterations else (int(max_niterations) if max_niterations else math.inf)
    r = coupled_ccpf_chains_batch(model, theta, observations, int(algoparameters["nparticles"]), 1, K=0, k=0, max_iterations=mi,
                                  with_as=bool(algoparameters.get("with_as", False)), seed=seed, first_replicate=replicate)
    it = int(force_niterations) if force_niterations else int(r["iteration"][0])
    return {"estimate": r["hbar"][0], "iteration": it, "truncation": math.inf}


def rheeglynn_estimator_RB(observations, model, theta, algoparameters, test_function=None, seed=None, replicate=0):
    """rheeglynn_estimator_RB / rheeglynn_estimator_RB2 (algoparameters["m"]) -- R/rheeglynn_estimator_RB.R:170-265, identity
    test function: the Rhee-Glynn sum over the filters' weighted averages of all N paths instead of the one drawn path."""
    if test_function is not None:
        raise NotImplementedError("the fused Rao-Blackwellised estimate is the identity test function (the reference's default)")
    m = int(algoparameters.get("m", 0) or 0)
    r = coupled_ccpf_chains_batch(model, theta, observations, int(algoparameters["nparticles"]), 1, K=m, k=m,
                                  with_as=bool(algoparameters.get("with_as", False)), seed=seed, first_replicate=replicate, rb=True)
    tau = int(r["meetingtime"][0])
    return {"estimate": r["hbar_rb"][0], "iteration": int(r["iteration"][0]), "meetingtime": math.inf if tau < 0 else tau}


def unbiased_ccpf_estimator_sharded(model, theta, observations, nparticles, nreplicates, K=1, k=0, max_iterations=math.inf, with_as=False,
                                    seed=0, rb=False, **kw):
    """The coupled-CCPF estimator over all ranks of torch.distributed (or one process): rank g runs the replicates
    shard_range(nreplicates, g, G) of coupled_ccpf_chains_batch and ONE all-reduce combines [sum H_bar, sum H_bar^2, filters,
    replicates] -- the multi-GPU form of foreach(r) coupled_ccpf_chains + H_bar.  Returns mean / standard error per component."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_range(nreplicates, rank, world)
    g = coupled_ccpf_chains_batch(model, theta, observations, nparticles, hi - lo, K=K, k=k, max_iterations=max_iterations, with_as=with_as,
                                  seed=seed, first_replicate=lo, rb=rb, **kw)
    key, key2 = ("sum_hbar_rb", "sum_hbar_rb_sq") if rb else ("sum_hbar", "sum_hbar_sq")
    buf = np.concatenate([g[key].T.ravel(), g[key2].T.ravel(), [g["nfilters"], hi - lo]])     # [T+1][d] order of combine_sums
    t = torch.tensor(buf, dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
    if world > 1:
        dist.all_reduce(t)
    out = combine_sums(t.cpu().numpy(), model.dimension, np.asarray(observations).shape[0])
    out["meetingtime_local"] = g["meetingtime"]
    return out


def shard_range(nreplicates, rank, world_size):
    """Replicates [lo, hi) owned by `rank`: contiguous blocks, remainder spread over the first ranks."""
    base, rem = divmod(int(nreplicates), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_sums(sums, dim, nobs):
    """Turn the all-reduced buffer [sum_hbar, sum_hbar_sq, nfilters, nreplicates] into mean / standard error."""
    P = dim * (nobs + 1)
    n = float(sums[2 * P + 1])
    mean = np.asarray(sums[:P]) / n
    var = np.maximum(np.asarray(sums[P:2 * P]) / n - mean ** 2, 0.0) * (n / max(n - 1.0, 1.0))
    shape = (nobs + 1, dim)
    return {"mean": mean.reshape(shape).T.copy(), "se": np.sqrt(var / n).reshape(shape).T.copy(), "nfilters": int(round(float(sums[2 * P]))), "nreplicates": int(round(n))}


def unbiased_estimator_sharded(model, theta, observations, nparticles, nreplicates, K=1, k=0, max_iterations=math.inf, seed=17,
                               stream=None, process_group=None, chain_batch=None):
    """Coupled-PIMH unbiased smoothing estimator with replicates sharded over the ranks of a torch.distributed job
    (one process per GPU).  Each rank runs its contiguous block of replicate ids -- Philox streams are keyed by the global
    replicate id, so the per-replicate results do not depend on the number of GPUs -- and ONE all-reduce (NCCL on GPUs)
    combines [sum H_bar, sum H_bar^2, filters run, replicates]; nothing is exchanged inside the filter loop."""
    import torch
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(process_group) if distributed else 0
    world = dist.get_world_size(process_group) if distributed else 1
    lo, hi = shard_range(nreplicates, rank, world)
    cb = chain_batch if chain_batch is not None else ChainBatch(model, theta, observations, nparticles, max(hi - lo, 1))
    try:
        sums = torch.zeros(cb.nsums, dtype=torch.float64, device="cuda")
        if hi > lo:
            cb.run(hi - lo, seed, lo, k, K, max_iterations, d_sums=sums.data_ptr(), stream=stream or torch.cuda.current_stream().cuda_stream)
            cb.fetch(hbar=False)          # surfaces per-replicate errors; synchronises the stream
        if distributed and world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=process_group)
        res = combine_sums(sums.cpu().numpy(), model.dimension, cb.nobs)
    finally:
        if chain_batch is None:
            cb.close()
    return res
