"""The particle filters: `CPF` (ref_trajectory=NULL path) and `particlefilter`, plus their batched twins.

Reference: R/CPF.R:14-78 and R/particlefilter.R:11-44.  The R functions run one filter; the CUDA library runs
B filters per launch (one CTA each), so `cpf_batch` / `PfBatch` are the calls that use the GPU well and the
single-filter functions are thin B = 1 wrappers with the reference's return values.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import rng


class PfBatch:
    """Device-resident batch of bootstrap filters (cpimh_pf handle).  Holds observations, workspaces and outputs in
    HBM across runs; `run` is asynchronous on the given CUDA stream."""

    def __init__(self, model, theta, observations, nparticles, max_filters, shuffle=0, tree_capacity=0, threads_per_filter=0,
                 store_paths=1, sk_max_events=64, device=None):
        obs = np.asarray(observations, dtype=np.float64)
        self.nobs = obs.shape[0]
        self.model = model
        self.cfg = model.config(nparticles, theta, nobs=self.nobs, shuffle=shuffle, tree_capacity=tree_capacity,
                                threads_per_filter=threads_per_filter, store_paths=store_paths, sk_max_events=sk_max_events)
        if device is not None:
            L.check(L.lib().cpimh_set_device(C.c_int(int(device))))
        self.h = C.c_void_p()
        L.check(L.lib().cpimh_pf_create(C.byref(self.cfg), C.c_int64(int(max_filters)), C.byref(self.h)))
        self.max_filters = int(max_filters)
        self.set_observations(obs)
        self.last = 0

    def close(self):
        if getattr(self, "h", None):
            L.lib().cpimh_pf_destroy(self.h)
            self.h = None

    __del__ = close

    def set_observations(self, observations):
        ob = L.obs_matrix(observations, self.nobs)
        L.check(L.lib().cpimh_pf_set_observations(self.h, ob.ctypes.data_as(C.c_void_p)))

    def set_noise(self, nfilters, noise):
        if noise is None:
            L.check(L.lib().cpimh_pf_set_noise(self.h, C.c_int64(0), None))
            return
        nz, keep = L.make_noise(noise)
        L.check(L.lib().cpimh_pf_set_noise(self.h, C.c_int64(int(nfilters)), C.byref(nz)))

    def record_ancestors(self, on=True):
        L.check(L.lib().cpimh_pf_record_ancestors(self.h, C.c_int(int(on))))

    def all_particles(self, on=True):
        """CPF(..., all_particles = TRUE) for the whole batch: the kernel also reduces the weighted average of the N paths."""
        L.check(L.lib().cpimh_pf_all_particles(self.h, C.c_int(int(on))))

    def fetch_exp_paths(self, nfilters=None, stream=None):
        B = self.last if nfilters is None else int(nfilters)
        ep = np.empty((B, self.cfg.nobs + 1, self.cfg.dim))
        L.check(L.lib().cpimh_pf_fetch_exp_paths(self.h, C.c_int64(B), C.c_void_p(stream or 0), L.ptr(ep)))
        return ep.transpose(0, 2, 1)                         # (B, d, T+1)

    def run(self, nfilters, seed, first_filter=0, stream=None):
        L.check(L.lib().cpimh_pf_run(self.h, C.c_int64(int(nfilters)), C.c_uint64(int(seed)), C.c_uint64(int(first_filter)), C.c_void_p(stream or 0)))
        self.last = int(nfilters)

    def fetch(self, nfilters=None, stream=None, logz_t=False, trajectory=True, path_index=False):
        B = self.last if nfilters is None else int(nfilters)
        d, T = self.cfg.dim, self.cfg.nobs
        out = {"logz": np.empty(B)}
        lzt = np.empty((B, T)) if logz_t else None
        traj = np.empty((B, T + 1, d)) if (trajectory and self.cfg.store_paths) else None
        pidx = np.empty(B, dtype=np.int32) if path_index else None
        L.check(L.lib().cpimh_pf_fetch(self.h, C.c_int64(B), C.c_void_p(stream or 0), L.ptr(out["logz"]), L.ptr(lzt), L.ptr(traj), L.ptr(pidx)))
        if lzt is not None:
            out["logz_t"] = lzt
        if traj is not None:
            out["trajectory"] = traj.transpose(0, 2, 1)      # (B, d, T+1) like R's d x (T+1)
        if pidx is not None:
            out["path_index"] = pidx
        return out

    def fetch_all_particles(self, b=0, stream=None, exp_path=True, trajectories=False, normweights=True):
        d, T, N = self.cfg.dim, self.cfg.nobs, self.cfg.nparticles
        ep = np.empty((T + 1, d)) if exp_path else None
        tr = np.empty((N, T + 1, d)) if trajectories else None
        nw = np.empty(N) if normweights else None
        L.check(L.lib().cpimh_pf_fetch_all_particles(self.h, C.c_int64(int(b)), C.c_void_p(stream or 0), L.ptr(ep), L.ptr(tr), L.ptr(nw)))
        out = {}
        if ep is not None:
            out["exp_path"] = np.ascontiguousarray(ep.T)                      # d x (T+1)
        if tr is not None:
            out["trajectories"] = np.ascontiguousarray(tr.transpose(2, 1, 0))  # d x (T+1) x N
        if nw is not None:
            out["weights"] = nw
        return out

    def fetch_ancestors(self, b=0, stream=None):
        anc = np.empty((self.cfg.nobs, self.cfg.nparticles), dtype=np.int32)
        L.check(L.lib().cpimh_pf_fetch_ancestors(self.h, C.c_int64(int(b)), C.c_void_p(stream or 0), L.ptr(anc)))
        return anc

    def exact_mode(self, on=True):
        """Wide filter only (cpimh_pf_exact_mode): every step with the reference-order CDF; the other filters are always exact."""
        L.check(L.lib().cpimh_pf_exact_mode(self.h, C.c_int(1 if on else 0)))

    def fetch_exact_repeats(self, nfilters=None, stream=None):
        """Steps each filter of the last run repeated with sequential-order sums (a sorted uniform within rounding distance
        of a CDF knot; DESIGN.md "Exact ancestors")."""
        B = int(nfilters if nfilters is not None else self.max_filters)
        rep = np.empty(B, dtype=np.int32)
        L.check(L.lib().cpimh_pf_fetch_exact_repeats(self.h, C.c_int64(B), C.c_void_p(stream or 0), L.ptr(rep)))
        return rep

    def device_outputs(self):
        a, b = C.c_void_p(), C.c_void_p()
        L.check(L.lib().cpimh_pf_device_outputs(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def launch_count(self):
        return int(L.lib().cpimh_pf_launch_count(self.h))

    @property
    def device_bytes(self):
        return int(L.lib().cpimh_pf_device_bytes(self.h))

    def effective_config(self):
        c = L.Config()
        L.check(L.lib().cpimh_pf_effective_config(self.h, C.byref(c)))
        return {"tree_capacity": c.tree_capacity, "threads_per_filter": c.threads_per_filter, "ctas_per_sm": c.reserved,
                "path_storage": {0: "none", 2: "dense", 3: "tree"}.get(c.store_paths, str(c.store_paths))}


def cpf_batch(nparticles, model, theta, observations, nfilters, seed=None, first_filter=0, noise=None, logz_t=False, all_particles=False, **kw):
    """B independent CPF(nparticles, model, theta, observations[, all_particles = TRUE]) runs through the one-shot host API
    (cpimh_pf_batch / cpimh_pf_batch_all_particles: the entries the R shim calls; their HBM workspaces stay cached in the library)."""
    if all_particles:
        obs = np.asarray(observations, dtype=np.float64)
        T = obs.shape[0]
        cfg = model.config(nparticles, theta, nobs=T, **kw)
        ob = L.obs_matrix(obs, T)
        B, d = int(nfilters), model.dimension
        logz, traj, ep = np.empty(B), np.empty((B, T + 1, d)), np.empty((B, T + 1, d))
        seed = rng.next_seed() if seed is None else int(seed)
        L.check(L.lib().cpimh_pf_batch_all_particles(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_int64(B), C.c_uint64(seed), C.c_uint64(int(first_filter)),
                                                      L.ptr(logz), L.ptr(traj), L.ptr(ep)))
        return {"logz": logz, "trajectory": traj.transpose(0, 2, 1), "exp_path": ep.transpose(0, 2, 1)}
    obs = np.asarray(observations, dtype=np.float64)
    T = obs.shape[0]
    cfg = model.config(nparticles, theta, nobs=T, **kw)
    ob = L.obs_matrix(obs, T)
    B, d = int(nfilters), model.dimension
    logz = np.empty(B)
    lzt = np.empty((B, T)) if logz_t else None
    traj = np.empty((B, T + 1, d)) if cfg.store_paths else None
    pidx = np.empty(B, dtype=np.int32)
    nz, keep = L.make_noise(noise)
    seed = rng.next_seed() if seed is None else int(seed)
    L.check(L.lib().cpimh_pf_batch(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_int64(B), C.c_uint64(seed), C.c_uint64(int(first_filter)),
                                    C.byref(nz) if noise else None, L.ptr(logz), L.ptr(lzt), L.ptr(traj), L.ptr(pidx)))
    out = {"logz": logz, "path_index": pidx}
    if traj is not None:
        out["trajectory"] = traj.transpose(0, 2, 1)
    if lzt is not None:
        out["logz_t"] = lzt
    return out


def CPF(nparticles, model, theta, observations, ref_trajectory=None, with_as=False, all_particles=False, seed=None, filter_id=0, noise=None, **kw):
    """CPF(nparticles, model, theta, observations, ref_trajectory = NULL, with_as = FALSE, all_particles = FALSE)
    -- R/CPF.R:14-78.  Returns {"trajectory": d x (T+1), "logz": float [, "exp_path": d x (T+1)]}."""
    seed = rng.next_seed() if seed is None else int(seed)
    if ref_trajectory is not None:
        if all_particles:
            raise NotImplementedError("all_particles = TRUE is wired for the bootstrap filter (ref_trajectory = NULL) only")
        r = ccpf_batch(nparticles, model, theta, observations, [ref_trajectory], with_as=with_as, seed=seed, first_filter=filter_id, noise=noise, **kw)
        return {"trajectory": np.ascontiguousarray(r["trajectory1"][0]), "logz": float(r["logz"][0])}
    # with_as without a reference trajectory has no effect in the reference either (R/CPF.R:45)
    if not all_particles:
        r = cpf_batch(nparticles, model, theta, observations, 1, seed=seed, first_filter=filter_id, noise=noise, **kw)
        return {"trajectory": np.ascontiguousarray(r["trajectory"][0]), "logz": float(r["logz"][0])}
    if model.dimension >= 32:     # wide filter (one filter across the GPU): the one-shot entry computes the all-particle average
        r = cpf_batch(nparticles, model, theta, observations, 1, seed=seed, first_filter=filter_id, all_particles=True, **kw)
        return {"trajectory": np.ascontiguousarray(r["trajectory"][0]), "exp_path": np.ascontiguousarray(r["exp_path"][0]), "logz": float(r["logz"][0])}
    pb = PfBatch(model, theta, observations, nparticles, 1, **kw)
    try:
        if noise:
            pb.set_noise(1, noise)
        pb.run(1, seed, filter_id)
        r = pb.fetch()
        ap = pb.fetch_all_particles(0)
    finally:
        pb.close()
    return {"trajectory": np.ascontiguousarray(r["trajectory"][0]), "exp_path": ap["exp_path"], "logz": float(r["logz"][0])}


def ccpf_batch(nparticles, model, theta, observations, ref1, ref2=None, with_as=False, seed=None, first_filter=0, noise=None,
               ancestors=False, exp_path=False, **kw):
    """B conditional particle filters (cpimh_ccpf_batch).  ref1 / ref2: (B, d, T+1) reference trajectories; ref2 = None runs
    CPF(ref_trajectory = ref1[b], with_as) (R/CPF.R:14-78), otherwise CPF_coupled(ref1[b], ref2[b], CR_indexmatching, with_as)
    (R/CPF_coupled.R:14-112).  Filter b uses Philox stream first_filter + b for both systems."""
    obs = np.asarray(observations, dtype=np.float64)
    T = obs.shape[0]
    cfg = model.config(nparticles, theta, nobs=T, **kw)
    ob = L.obs_matrix(obs, T)
    d, N = model.dimension, int(nparticles)
    r1 = np.ascontiguousarray(np.asarray(ref1, dtype=np.float64).reshape(-1, d, T + 1).transpose(0, 2, 1))     # (B, T+1, d) = d x (T+1) col-major
    B = r1.shape[0]
    nsys = 1 if ref2 is None else 2
    r2 = np.ascontiguousarray(np.asarray(ref2, dtype=np.float64).reshape(-1, d, T + 1).transpose(0, 2, 1)) if nsys == 2 else None
    if nsys == 2 and r2.shape != r1.shape:
        raise ValueError("ref1 and ref2 must have the same shape")
    t1 = np.empty((B, T + 1, d)); t2 = np.empty((B, T + 1, d)) if nsys == 2 else None
    logz = np.empty(B); pidx = np.empty((B, 2), dtype=np.int32)
    anc = np.empty((B, 2, T, N), dtype=np.int32) if ancestors else None
    e1 = np.empty((B, T + 1, d)) if exp_path else None
    e2 = np.empty((B, T + 1, d)) if (exp_path and nsys == 2) else None
    nz, keep = L.make_noise(noise)
    seed = rng.next_seed() if seed is None else int(seed)
    L.check(L.lib().cpimh_ccpf_batch(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_int64(B), C.c_uint64(seed), C.c_uint64(int(first_filter)),
                                      C.c_int(nsys), C.c_int(int(bool(with_as))), L.ptr(r1), L.ptr(r2), C.byref(nz) if noise else None,
                                      L.ptr(t1), L.ptr(t2), L.ptr(logz), L.ptr(pidx), L.ptr(anc), L.ptr(e1), L.ptr(e2)))
    out = {"trajectory1": t1.transpose(0, 2, 1), "logz": logz, "path_index": pidx}
    if exp_path:
        out["exp_path1"] = e1.transpose(0, 2, 1)
        if nsys == 2:
            out["exp_path2"] = e2.transpose(0, 2, 1)
    if nsys == 2:
        out["trajectory2"] = t2.transpose(0, 2, 1)
    if ancestors:
        out["ancestors"] = anc
    return out


def CPF_coupled(nparticles, model, theta, observations, ref_trajectory1, ref_trajectory2, coupled_resampling=None, with_as=False,
                seed=None, filter_id=0, noise=None, **kw):
    """CPF_coupled(nparticles, model, theta, observations, ref_trajectory1, ref_trajectory2, coupled_resampling, with_as)
    -- R/CPF_coupled.R:14-112.  The coupled resampling scheme is index matching (what get_cpf_kernels passes,
    R/pf_kernels.R:238), fused into the kernel; `coupled_resampling` is accepted for signature compatibility.
    Returns {"new_trajectory1", "new_trajectory2"}: d x (T+1) each."""
    r = ccpf_batch(nparticles, model, theta, observations, [ref_trajectory1], [ref_trajectory2], with_as=with_as, seed=seed,
                   first_filter=filter_id, noise=noise, **kw)
    return {"new_trajectory1": np.ascontiguousarray(r["trajectory1"][0]), "new_trajectory2": np.ascontiguousarray(r["trajectory2"][0])}


def CPF_RB(nparticles, model, theta, observations, ref_trajectory=None, with_as=False, test_function=None, seed=None, filter_id=0, **kw):
    """CPF_RB(nparticles, model, theta, observations, ref_trajectory, with_as, test_function) -- R/rheeglynn_estimator_RB.R:2-60.
    `estimate` is the weighted average of test_function over all N paths, reduced inside the filter kernel; only the default
    test_function (identity) is fused.  Returns {"new_trajectory", "estimate"}: d x (T+1) each."""
    if test_function is not None:
        raise NotImplementedError("the fused Rao-Blackwellised estimate is the identity test function (the reference's default)")
    if ref_trajectory is None:
        r = CPF(nparticles, model, theta, observations, all_particles=True, seed=seed, filter_id=filter_id, **kw)
        return {"new_trajectory": r["trajectory"], "estimate": r["exp_path"]}
    r = ccpf_batch(nparticles, model, theta, observations, [ref_trajectory], with_as=with_as, seed=seed, first_filter=filter_id, exp_path=True, **kw)
    return {"new_trajectory": np.ascontiguousarray(r["trajectory1"][0]), "estimate": np.ascontiguousarray(r["exp_path1"][0])}


def CPF_RB_coupled(nparticles, model, theta, observations, ref_trajectory1, ref_trajectory2, coupled_resampling=None, with_as=False,
                   test_function=None, seed=None, filter_id=0, **kw):
    """CPF_RB_coupled(...) -- R/rheeglynn_estimator_RB.R:62-167 (identity test function).
    Returns {"new_trajectory1", "new_trajectory2", "estimate1", "estimate2"}."""
    if test_function is not None:
        raise NotImplementedError("the fused Rao-Blackwellised estimate is the identity test function (the reference's default)")
    r = ccpf_batch(nparticles, model, theta, observations, [ref_trajectory1], [ref_trajectory2], with_as=with_as, seed=seed,
                   first_filter=filter_id, exp_path=True, **kw)
    return {"new_trajectory1": np.ascontiguousarray(r["trajectory1"][0]), "new_trajectory2": np.ascontiguousarray(r["trajectory2"][0]),
            "estimate1": np.ascontiguousarray(r["exp_path1"][0]), "estimate2": np.ascontiguousarray(r["exp_path2"][0])}


def pf_batch(y, theta, model, nparticles, nfilters, ess_threshold=1, systematic=False, seed=None, first_filter=0, ancestors=False, **kw):
    """B runs of pf() (cpimh_pf_adaptive_batch): loglik (B,), loglik_t (B, T) cumulative, pf_means (B, T, d), path (B, T+1, d),
    resampled (B, T) flags [, ancestors (B, T, N)]."""
    obs = np.asarray(y, dtype=np.float64)
    T = obs.shape[0]
    cfg = model.config(nparticles, theta, nobs=T, **kw)
    ob = L.obs_matrix(obs, T)
    B, d, N = int(nfilters), model.dimension, int(nparticles)
    ll, llt, pm, pa = np.empty(B), np.empty((B, T)), np.empty((B, T, d)), np.empty((B, T + 1, d))
    rs = np.empty((B, T), dtype=np.int32)
    anc = np.empty((B, T, N), dtype=np.int32) if ancestors else None
    seed = rng.next_seed() if seed is None else int(seed)
    L.check(L.lib().cpimh_pf_adaptive_batch(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_int64(B), C.c_uint64(seed), C.c_uint64(int(first_filter)),
                                             C.c_double(float(ess_threshold)), C.c_int(int(bool(systematic))), L.ptr(ll), L.ptr(llt), L.ptr(pm), L.ptr(pa),
                                             L.ptr(rs), L.ptr(anc)))
    out = {"loglik": ll, "loglik_t": llt, "pf_means": pm, "path": pa, "resampled": rs}
    if ancestors:
        out["ancestors"] = anc
    return out


def pf(y, theta, model, nparticles, ess_threshold=1, systematic=False, seed=None, filter_id=0, **kw):
    """pf(y, theta, model, nparticles, ess_threshold = 1, systematic = FALSE) -- R/CPF.R:106-155: bootstrap filter with
    ESS-triggered resampling.  Returns {"pf_means": T x d, "loglik": float, "path": (T+1) x d, "loglik_t": cumulative (T,)}."""
    r = pf_batch(y, theta, model, nparticles, 1, ess_threshold=ess_threshold, systematic=systematic, seed=seed, first_filter=filter_id, **kw)
    return {"pf_means": r["pf_means"][0], "loglik": float(r["loglik"][0]), "path": r["path"][0], "loglik_t": r["loglik_t"][0]}


def particlefilter(nparticles, model, theta, observations, seed=None, filter_id=0, noise=None, **kw):
    """particlefilter(nparticles, model, theta, observations) -- R/particlefilter.R:11-44.
    Returns {"trajectories": d x (T+1) x N, "weights": N}."""
    seed = rng.next_seed() if seed is None else int(seed)
    pb = PfBatch(model, theta, observations, nparticles, 1, **kw)
    try:
        if noise:
            pb.set_noise(1, noise)
        pb.run(1, seed, filter_id)
        pb.fetch(trajectory=False)
        ap = pb.fetch_all_particles(0, exp_path=False, trajectories=True)
    finally:
        pb.close()
    return {"trajectories": ap["trajectories"], "weights": ap["weights"]}


def pf_get_weighted_average(arr_test, w):
    """R/CPF.R:84-92: sum_i arr[:, :, i] * w[i].  Device version lives in cpimh_pf_fetch_all_particles; this host helper
    only mirrors the exported R utility for arrays the user already holds."""
    arr = np.asarray(arr_test, dtype=np.float64)
    return np.tensordot(arr, np.asarray(w, dtype=np.float64), axes=([2], [0]))
