"""ctypes binding of the CPU oracle (oracle/build/liboracle.so).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "build", "liboracle.so")
SOURCES = ["philox.c", "resampling.c", "tree.c", "mvnorm.c", "models.c", "cpf.c", "ccpf.c", "apf.c"]

MODEL_GSS, MODEL_AR, MODEL_LEVY_SV, MODEL_SK, MODEL_LORENZ = range(5)
INTEGRATOR_RK4, INTEGRATOR_DOPRI5 = 0, 1
MAX_THETA = 32


def build(force=False):
    """Compile the oracle with gcc (plain `gcc` from PATH; $CC in this image points at a wrapper without OpenMP specs)."""
    srcs = [os.path.join(HERE, s) for s in SOURCES] + [os.path.join(HERE, "oracle.h")]
    deps = srcs + [os.path.join(HERE, "rngmath.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in deps):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    base = ["gcc", "-O2", "-std=gnu11", "-fPIC", "-ffp-contract=off", "-shared", "-o", LIB_PATH] + srcs[:-1] + ["-lm"]
    # -mfma only makes the explicit fma() calls of rngmath.h inline (libm's fma gives the same correctly rounded result)
    for extra in (["-mfma", "-fopenmp"], ["-fopenmp"], []):
        try:
            subprocess.run(base[:5] + extra + base[5:], check=True, capture_output=True, cwd=HERE)
            return LIB_PATH
        except subprocess.CalledProcessError:
            continue
    subprocess.run(base, check=True, cwd=HERE)
    return LIB_PATH


class Config(C.Structure):
    _fields_ = [("model", C.c_int), ("dim", C.c_int), ("dim_y", C.c_int), ("nparticles", C.c_int), ("nobs", C.c_int),
                ("shuffle", C.c_int), ("integrator", C.c_int), ("substeps", C.c_int), ("sk_max_events", C.c_int),
                ("ntheta", C.c_int), ("theta", C.c_double * MAX_THETA)]


class Noise(C.Structure):
    _fields_ = [("init", C.c_void_p), ("trans", C.c_void_p), ("resample", C.c_void_p), ("shuffle", C.c_void_p), ("path_u", C.c_void_p)]


class PfOut(C.Structure):
    _fields_ = [("trajectory", C.c_void_p), ("logz", C.c_void_p), ("logz_t", C.c_void_p), ("exp_path", C.c_void_p),
                ("all_paths", C.c_void_p), ("normweights", C.c_void_p), ("ancestors_t", C.c_void_p), ("path_index", C.c_void_p),
                ("tree_M_final", C.c_void_p), ("clamp_count", C.c_void_p)]


class Chains(C.Structure):
    _fields_ = [("iteration", C.c_int), ("meetingtime", C.c_int), ("finished", C.c_int), ("p", C.c_int), ("nrows1", C.c_int),
                ("samples1", C.POINTER(C.c_double)), ("samples2", C.POINTER(C.c_double)),
                ("log_pdf1", C.POINTER(C.c_double)), ("log_pdf2", C.POINTER(C.c_double)), ("nfilters", C.c_int),
                ("exp_samples1", C.POINTER(C.c_double)), ("exp_samples2", C.POINTER(C.c_double))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_version.restype = C.c_char_p
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_config(model, dim, dim_y, nparticles, nobs, theta=(), shuffle=0, integrator=0, substeps=1, sk_max_events=64):
    c = Config()
    c.model, c.dim, c.dim_y, c.nparticles, c.nobs = model, dim, dim_y, nparticles, nobs
    c.shuffle, c.integrator, c.substeps, c.sk_max_events = shuffle, integrator, substeps, sk_max_events
    th = list(np.ravel(np.asarray(theta, dtype=np.float64)))
    c.ntheta = len(th)
    for i, v in enumerate(th):
        c.theta[i] = v
    return c


# ---------------------------------------------------------------- RNG
def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def philox_uniform2(seed, c0, c1, c2, c3):
    ua, ub = C.c_double(), C.c_double()
    lib().orc_philox_uniform2(C.c_uint64(seed), C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), C.byref(ua), C.byref(ub))
    return ua.value, ub.value


# ---------------------------------------------------------------- resamplers (0-based, like the reference's natives)
def systematic_resampling_n_(w, ndraws, u):
    w = _f64(w); anc = np.empty(ndraws, dtype=np.int32)
    lib().orc_systematic_resampling_n(_p(w), C.c_int(len(w)), C.c_int(ndraws), C.c_double(u), _p(anc))
    return anc


def sorted_uniforms(u_raw):
    u = _f64(u_raw); e = np.empty_like(u)
    lib().orc_sorted_uniforms(_p(u), C.c_int(len(u)), _p(e))
    return e


def multinomial_from_sorted(w, ndraws, e_sorted, scale_by_sumw=True):
    w = _f64(w); e = _f64(e_sorted); anc = np.empty(ndraws, dtype=np.int32)
    lib().orc_multinomial_from_sorted(_p(w), C.c_int(len(w)), C.c_int(ndraws), _p(e), C.c_int(int(scale_by_sumw)), _p(anc))
    return anc


def random_shuffle_int(a, u_shuffle):
    a = np.ascontiguousarray(a, dtype=np.int32).copy(); u = _f64(u_shuffle)
    lib().orc_random_shuffle_int(_p(a), C.c_int(len(a)), _p(u))
    return a


def multinomial_resampling_n_(w, ndraws, u_raw, u_shuffle=None):
    w = _f64(w); u = _f64(u_raw); us = None if u_shuffle is None else _f64(u_shuffle); anc = np.empty(ndraws, dtype=np.int32)
    lib().orc_multinomial_resampling_n(_p(w), C.c_int(len(w)), C.c_int(ndraws), _p(u), _p(us), _p(anc))
    return anc


def coupled_multinomial_resampling_n_(w1, w2, ndraws, u_raw, u_shuffle=None):
    w1 = _f64(w1); w2 = _f64(w2); u = _f64(u_raw); us = None if u_shuffle is None else _f64(u_shuffle)
    anc = np.empty(2 * ndraws, dtype=np.int32)
    lib().orc_coupled_multinomial_resampling_n(_p(w1), _p(w2), C.c_int(len(w1)), C.c_int(ndraws), _p(u), _p(us), _p(anc))
    return anc.reshape(2, ndraws).T.copy()


def indexmatching(w1, w2, ndraws, u_couple, u_mult_raw, u_mult_shuffle, u_cm_raw, u_cm_shuffle):
    w1 = _f64(w1); w2 = _f64(w2)
    arrs = [None if a is None else _f64(a) for a in (u_couple, u_mult_raw, u_mult_shuffle, u_cm_raw, u_cm_shuffle)]
    anc = np.empty(2 * ndraws, dtype=np.int32)
    lib().orc_indexmatching(_p(w1), _p(w2), C.c_int(len(w1)), C.c_int(ndraws), *[_p(a) for a in arrs], _p(anc))
    return anc.reshape(2, ndraws).T.copy()


def normalize_weight(logw):
    lw = _f64(logw); w = np.empty_like(lw)
    lib().orc_normalize_weight(_p(lw), C.c_int(len(lw)), _p(w))
    return w


def logz_increment(logw):
    lw = _f64(logw); w = np.empty_like(lw)
    lib().orc_logz_increment.restype = C.c_double
    lz = lib().orc_logz_increment(_p(lw), C.c_int(len(lw)), _p(w))
    return lz, w


# ---------------------------------------------------------------- tree
class Tree:
    """Mirror of the reference's Rcpp module class (src/tree.cpp:144-162); x matrices are d x N (numpy shape (d, N))."""

    def __init__(self, N, M, dimx):
        L = lib()
        L.orc_tree_new.restype = C.c_void_p
        self.h = C.c_void_p(L.orc_tree_new(C.c_int(N), C.c_int(M), C.c_int(dimx)))
        self.N, self.dimx = N, dimx

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tree_free(self.h); self.h = None

    @staticmethod
    def _colmajor(x):
        return np.asfortranarray(np.asarray(x, dtype=np.float64))

    def init(self, x0):
        x = self._colmajor(x0); lib().orc_tree_init(self.h, x.ctypes.data_as(C.c_void_p))

    def update(self, x, a):
        x = self._colmajor(x); a = np.ascontiguousarray(a, dtype=np.int32)
        lib().orc_tree_update(self.h, x.ctypes.data_as(C.c_void_p), _p(a))

    @property
    def M(self):
        return lib().orc_tree_M(self.h)

    @property
    def nsteps(self):
        return lib().orc_tree_nsteps(self.h)

    def live_nodes(self):
        return lib().orc_tree_live_nodes(self.h)

    def get_path(self, n):
        path = np.empty((self.dimx, self.nsteps + 1), dtype=np.float64, order="F")
        lib().orc_tree_get_path(self.h, C.c_int(n), path.ctypes.data_as(C.c_void_p))
        return np.ascontiguousarray(path)

    def retrieve_xgeneration(self, lag):
        xg = np.empty((self.dimx, self.N), dtype=np.float64, order="F")
        lib().orc_tree_retrieve_xgeneration(self.h, C.c_int(lag), xg.ctypes.data_as(C.c_void_p))
        return np.ascontiguousarray(xg)

    def _arr(self, fn, n, ctype, dtype):
        f = getattr(lib(), fn); f.restype = C.POINTER(ctype)
        return np.ctypeslib.as_array(f(self.h), shape=(n,)).astype(dtype).copy()

    @property
    def a_star(self):
        return self._arr("orc_tree_a_star", self.M, C.c_int, np.int32)

    @property
    def o_star(self):
        return self._arr("orc_tree_o_star", self.M, C.c_int, np.int32)

    @property
    def l_star(self):
        return self._arr("orc_tree_l_star", self.N, C.c_int, np.int32)


# ---------------------------------------------------------------- mvnorm
def dmvnorm_transpose_cholesky(x, mean, L):
    x = np.asfortranarray(x, dtype=np.float64); d, n = x.shape
    Lm = np.asfortranarray(L, dtype=np.float64); m = _f64(mean); out = np.empty(n)
    lib().orc_dmvnorm_transpose_cholesky(x.ctypes.data_as(C.c_void_p), C.c_int(d), C.c_int(n), _p(m), Lm.ctypes.data_as(C.c_void_p), _p(out))
    return out


def dmvnorm_transpose(x, mean, cov):
    x = np.asfortranarray(x, dtype=np.float64); d, n = x.shape
    Cm = np.asfortranarray(cov, dtype=np.float64); m = _f64(mean); out = np.empty(n)
    rc = lib().orc_dmvnorm_transpose(x.ctypes.data_as(C.c_void_p), C.c_int(d), C.c_int(n), _p(m), Cm.ctypes.data_as(C.c_void_p), _p(out))
    if rc:
        raise ValueError("covariance is not positive definite")
    return out


def dmvnorm(x, mean, cov):
    x = np.asfortranarray(x, dtype=np.float64); n, d = x.shape
    Cm = np.asfortranarray(cov, dtype=np.float64); m = _f64(mean); out = np.empty(n)
    rc = lib().orc_dmvnorm(x.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(d), _p(m), Cm.ctypes.data_as(C.c_void_p), _p(out))
    if rc:
        raise ValueError("covariance is not positive definite")
    return out


def rmvnorm_transpose_cholesky(n, mean, L, z):
    Lm = np.asfortranarray(L, dtype=np.float64); d = Lm.shape[0]; m = _f64(mean)
    zz = np.asfortranarray(z, dtype=np.float64); out = np.empty((d, n), order="F")
    lib().orc_rmvnorm_transpose_cholesky(C.c_int(n), C.c_int(d), _p(m), Lm.ctypes.data_as(C.c_void_p), zz.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return np.ascontiguousarray(out)


def rmvnorm_transpose(n, mean, cov, z):
    Cm = np.asfortranarray(cov, dtype=np.float64); d = Cm.shape[0]; m = _f64(mean)
    zz = np.asfortranarray(z, dtype=np.float64); out = np.empty((d, n), order="F")
    lib().orc_rmvnorm_transpose(C.c_int(n), C.c_int(d), _p(m), Cm.ctypes.data_as(C.c_void_p), zz.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return np.ascontiguousarray(out)


def rmvnorm(n, mean, cov, z):
    Cm = np.asfortranarray(cov, dtype=np.float64); d = Cm.shape[0]; m = _f64(mean)
    zz = np.asfortranarray(z, dtype=np.float64); out = np.empty((n, d), order="F")
    lib().orc_rmvnorm(C.c_int(n), C.c_int(d), _p(m), Cm.ctypes.data_as(C.c_void_p), zz.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return np.ascontiguousarray(out)


def create_A(alpha, d):
    A = np.empty((d, d), order="F")
    lib().orc_create_A(C.c_double(alpha), C.c_int(d), A.ctypes.data_as(C.c_void_p))
    return np.ascontiguousarray(A)


def lorenz_step(x, F, integrator=0, substeps=1, t0=0.0, t1=0.05, dt=0.05):
    x = _f64(x).copy()
    lib().orc_lorenz_step(C.c_int(len(x)), C.c_double(F), C.c_int(integrator), C.c_int(substeps), C.c_double(t0), C.c_double(t1), C.c_double(dt), _p(x))
    return x


# ---------------------------------------------------------------- models / noise
def n_init(cfg):
    return lib().orc_model_n_init(C.byref(cfg))


def n_trans(cfg):
    return lib().orc_model_n_trans(C.byref(cfg))


def fill_init_noise(cfg, seed, filter_id):
    out = np.empty((max(n_init(cfg), 1), cfg.nparticles))
    lib().orc_fill_init_noise(C.byref(cfg), C.c_uint64(seed), C.c_uint64(filter_id), _p(out))
    return out[:n_init(cfg)]


def fill_trans_noise(cfg, seed, filter_id, time):
    out = np.empty((max(n_trans(cfg), 1), cfg.nparticles))
    lib().orc_fill_trans_noise(C.byref(cfg), C.c_uint64(seed), C.c_uint64(filter_id), C.c_int(time), _p(out))
    return out[:n_trans(cfg)]


def model_rinit(cfg, noise):
    x = np.empty((cfg.dim, cfg.nparticles), order="F"); nz = _f64(noise)
    lib().orc_model_rinit(C.byref(cfg), _p(nz), x.ctypes.data_as(C.c_void_p))
    return np.ascontiguousarray(x)


def model_rtransition(cfg, x, time, noise):
    xx = np.asfortranarray(x, dtype=np.float64).copy(order="F"); nz = _f64(noise)
    rc = lib().orc_model_rtransition(C.byref(cfg), C.c_int(time), _p(nz), xx.ctypes.data_as(C.c_void_p))
    if rc:
        raise RuntimeError("oracle rtransition failed (%d)" % rc)
    return np.ascontiguousarray(xx)


def model_dmeasurement(cfg, x, y):
    xx = np.asfortranarray(x, dtype=np.float64); yy = _f64(y); out = np.empty(cfg.nparticles)
    lib().orc_model_dmeasurement(C.byref(cfg), xx.ctypes.data_as(C.c_void_p), _p(yy), _p(out))
    return out


# ---------------------------------------------------------------- filters
def cpf(cfg, obs, seed=0, filter_id=0, noise=None, all_particles=False, all_paths=False, want_ancestors=False):
    """CPF(nparticles, model, theta, observations) with ref_trajectory=NULL.  obs: (T, dim_y).  noise: dict of injected arrays."""
    T, d, N = cfg.nobs, cfg.dim, cfg.nparticles
    ob = np.asfortranarray(np.asarray(obs, dtype=np.float64).reshape(T, -1))
    traj = np.empty((d, T + 1), order="F"); logz = np.empty(1); logz_t = np.empty(T); normw = np.empty(N)
    exp_path = np.empty((d, T + 1), order="F") if all_particles else None
    paths = np.empty((N, T + 1, d)) if all_paths else None  # C-order (N, T+1, d) == R's d x (T+1) x N column-major
    anc = np.empty((T, N), dtype=np.int32) if want_ancestors else None
    pidx = np.empty(1, dtype=np.int32); Mfin = np.empty(1, dtype=np.int32); clamp = np.empty(1, dtype=np.int32)
    nz = Noise(); keep = []
    if noise:
        for k in ("init", "trans", "resample", "shuffle", "path_u"):
            if noise.get(k) is not None:
                a = _f64(noise[k]); keep.append(a); setattr(nz, k, a.ctypes.data)
    o = PfOut()
    o.trajectory = traj.ctypes.data; o.logz = logz.ctypes.data; o.logz_t = logz_t.ctypes.data; o.normweights = normw.ctypes.data
    o.exp_path = exp_path.ctypes.data if all_particles else None
    o.all_paths = paths.ctypes.data if all_paths else None
    o.ancestors_t = anc.ctypes.data if want_ancestors else None
    o.path_index = pidx.ctypes.data; o.tree_M_final = Mfin.ctypes.data; o.clamp_count = clamp.ctypes.data
    rc = lib().orc_cpf(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.byref(nz) if noise else None, C.c_uint64(seed), C.c_uint64(filter_id), C.byref(o))
    if rc:
        raise RuntimeError("oracle cpf failed (%d)" % rc)
    res = dict(trajectory=np.ascontiguousarray(traj), logz=float(logz[0]), logz_t=logz_t, normweights=normw,
               path_index=int(pidx[0]), tree_M=int(Mfin[0]), clamp_count=int(clamp[0]))
    if all_particles:
        res["exp_path"] = np.ascontiguousarray(exp_path)
    if all_paths:
        res["trajectories"] = np.ascontiguousarray(paths.transpose(2, 1, 0))  # (d, T+1, N) like R
    if want_ancestors:
        res["ancestors"] = anc
    return res


def cpf_batch(cfg, obs, seed, first_filter, nfilters, nthreads=0, want_traj=True):
    T, d = cfg.nobs, cfg.dim
    ob = np.asfortranarray(np.asarray(obs, dtype=np.float64).reshape(T, -1))
    logz = np.empty(nfilters); traj = np.empty((nfilters, T + 1, d)) if want_traj else None
    rc = lib().orc_cpf_batch(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_uint64(seed), C.c_uint64(first_filter), C.c_int(nfilters),
                             C.c_int(nthreads), _p(logz), _p(traj))
    if rc:
        raise RuntimeError("oracle cpf_batch failed (%d)" % rc)
    return logz, (traj.transpose(0, 2, 1) if want_traj else None)   # (B, d, T+1)


def coupled_pimh_chains(cfg, obs, seed, replicate, K=1, max_iterations=-1, run_discarded_init=False, all_particles=False, hbar_k=None):
    T = cfg.nobs
    ob = np.asfortranarray(np.asarray(obs, dtype=np.float64).reshape(T, -1))
    c = Chains()
    rc = lib().orc_coupled_pimh_chains_ap(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_uint64(seed), C.c_uint32(replicate), C.c_int(K),
                                          C.c_int(max_iterations), C.c_int(int(run_discarded_init)), C.c_int(int(all_particles)), C.byref(c))
    if rc:
        raise RuntimeError("oracle coupled_pimh_chains failed (%d)" % rc)
    p, it = c.p, c.iteration
    out = dict(samples1=np.ctypeslib.as_array(c.samples1, shape=(it + 1, p)).copy(),
               samples2=np.ctypeslib.as_array(c.samples2, shape=(max(it, 1), p))[:it].copy(),
               log_pdf1=np.ctypeslib.as_array(c.log_pdf1, shape=(it + 1,)).copy(),
               log_pdf2=np.ctypeslib.as_array(c.log_pdf2, shape=(max(it, 1),))[:it].copy(),
               meetingtime=(np.inf if c.meetingtime < 0 else c.meetingtime), iteration=it, finished=bool(c.finished), nfilters=c.nfilters)
    if all_particles:
        out["exp_samples1"] = np.ctypeslib.as_array(c.exp_samples1, shape=(it + 1, p)).copy()
        out["exp_samples2"] = np.ctypeslib.as_array(c.exp_samples2, shape=(max(it, 1), p))[:it].copy()
    if hbar_k is not None:                                 # H_bar / H_bar over exp_samples by the C restatement
        hb = np.empty(p)
        if lib().orc_H_bar(C.byref(c), C.c_int(int(hbar_k)), C.c_int(int(K)), _p(hb)) == 0:
            out["hbar"] = hb.copy()
        if all_particles and lib().orc_H_bar_exp(C.byref(c), C.c_int(int(hbar_k)), C.c_int(int(K)), _p(hb)) == 0:
            out["hbar_exp"] = hb.copy()
    lib().orc_chains_free(C.byref(c))
    return out


def coupled_pimh_batch(cfg, obs, seed, first_replicate, nrep, k=0, K=1, max_iterations=-1, nthreads=0):
    T = cfg.nobs; p = T + 1
    ob = np.asfortranarray(np.asarray(obs, dtype=np.float64).reshape(T, -1))
    hbar = np.empty((nrep, p)); mt = np.empty(nrep, dtype=np.int32); it = np.empty(nrep, dtype=np.int32); nf = C.c_longlong(0)
    rc = lib().orc_coupled_pimh_batch(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_uint64(seed), C.c_uint32(first_replicate), C.c_int(nrep),
                                      C.c_int(k), C.c_int(K), C.c_int(max_iterations), C.c_int(nthreads), _p(hbar), _p(mt), _p(it), C.byref(nf))
    if rc:
        raise RuntimeError("oracle coupled_pimh_batch failed (%d)" % rc)
    return dict(hbar=hbar, meetingtime=mt, iteration=it, nfilters=nf.value)


# ---------------------------------------------------------------- conditional filters
class CcpfOut(C.Structure):
    _fields_ = [("trajectory", C.c_void_p * 2), ("exp_path", C.c_void_p * 2), ("ancestors", C.c_void_p * 2), ("path_index", C.c_int * 2), ("logz", C.c_double),
                ("clamp_count", C.c_int)]


def ccpf(cfg, obs, ref1, ref2=None, with_as=False, seed=0, filter_id=0, noise=None, want_ancestors=False, exp_path=False):
    """CPF(..., ref_trajectory = ref1, with_as) when ref2 is None, else CPF_coupled(..., ref1, ref2, CR_indexmatching, with_as).
    refs and returned trajectories are d x (T+1)."""
    T, d, N = cfg.nobs, cfg.dim, cfg.nparticles
    ob = np.asfortranarray(np.asarray(obs, dtype=np.float64).reshape(T, -1))
    nsys = 1 if ref2 is None else 2
    refs = [np.asfortranarray(np.asarray(r, dtype=np.float64).reshape(d, T + 1)) for r in ([ref1] if nsys == 1 else [ref1, ref2])]
    tr = [np.empty((d, T + 1), order="F") for _ in range(2)]
    ep = [np.empty((d, T + 1), order="F") for _ in range(2)]
    anc = [np.empty((T, N), dtype=np.int32) for _ in range(2)] if want_ancestors else None
    nz = Noise(); keep = []
    if noise:
        for k in ("init", "trans"):
            if noise.get(k) is not None:
                a = _f64(noise[k]); keep.append(a); setattr(nz, k, a.ctypes.data)
    o = CcpfOut()
    for s_ in range(2):
        o.trajectory[s_] = tr[s_].ctypes.data
        o.exp_path[s_] = ep[s_].ctypes.data if exp_path else None
        o.ancestors[s_] = anc[s_].ctypes.data if want_ancestors else None
    lib().orc_ccpf.restype = C.c_int
    rc = lib().orc_ccpf(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.byref(nz) if noise else None, C.c_uint64(seed), C.c_uint64(filter_id),
                        C.c_int(nsys), refs[0].ctypes.data_as(C.c_void_p), refs[1].ctypes.data_as(C.c_void_p) if nsys == 2 else None,
                        C.c_int(int(with_as)), C.byref(o))
    if rc:
        raise RuntimeError("oracle ccpf failed (%d)" % rc)
    res = dict(trajectory1=np.ascontiguousarray(tr[0]), logz=float(o.logz), path_index=[int(o.path_index[0]), int(o.path_index[1])],
               clamp_count=int(o.clamp_count))
    if nsys == 2:
        res["trajectory2"] = np.ascontiguousarray(tr[1])
    if exp_path:
        res["exp_path1"] = np.ascontiguousarray(ep[0])
        if nsys == 2:
            res["exp_path2"] = np.ascontiguousarray(ep[1])
    if want_ancestors:
        res["ancestors1"] = anc[0]
        if nsys == 2:
            res["ancestors2"] = anc[1]
    return res


def model_dtransition(cfg, next_x, x, time):
    """model$dtransition(next_x, xparticles (d x N), theta, time): gss and ar only."""
    xx = np.asfortranarray(np.asarray(x, dtype=np.float64).reshape(cfg.dim, cfg.nparticles))
    out = np.empty(cfg.nparticles)
    lib().orc_model_dtransition(C.byref(cfg), _p(_f64(next_x)), xx.ctypes.data_as(C.c_void_p), C.c_int(int(time)), _p(out))
    return out


def coupled_ccpf_chains(cfg, obs, seed, replicate, K=1, max_iterations=-1, with_as=False, hbar_k=None):
    T = cfg.nobs
    ob = np.asfortranarray(np.asarray(obs, dtype=np.float64).reshape(T, -1))
    c = Chains()
    rc = lib().orc_coupled_ccpf_chains(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_uint64(seed), C.c_uint32(replicate), C.c_int(K),
                                       C.c_int(max_iterations), C.c_int(int(with_as)), C.byref(c))
    if rc:
        raise RuntimeError("oracle coupled_ccpf_chains failed (%d)" % rc)
    p, it = c.p, c.iteration
    out = dict(samples1=np.ctypeslib.as_array(c.samples1, shape=(it + 1, p)).copy(),
               samples2=np.ctypeslib.as_array(c.samples2, shape=(it, p)).copy(),
               log_pdf1=np.ctypeslib.as_array(c.log_pdf1, shape=(it + 1,)).copy(),
               log_pdf2=np.ctypeslib.as_array(c.log_pdf2, shape=(it,)).copy(),
               meetingtime=(np.inf if c.meetingtime < 0 else c.meetingtime), iteration=it, finished=bool(c.finished), nfilters=c.nfilters)
    if hbar_k is not None:
        hb = np.empty(p)
        if lib().orc_H_bar(C.byref(c), C.c_int(int(hbar_k)), C.c_int(int(K)), _p(hb)) == 0:
            out["hbar"] = hb.copy()
    lib().orc_chains_free(C.byref(c))
    return out


def coupled_ccpf_batch(cfg, obs, seed, first_replicate, nrep, k=0, K=1, max_iterations=-1, with_as=False, nthreads=0, rb=False):
    T = cfg.nobs; p = T + 1
    ob = np.asfortranarray(np.asarray(obs, dtype=np.float64).reshape(T, -1))
    hbar = np.empty((nrep, p)); mt = np.empty(nrep, dtype=np.int32); it = np.empty(nrep, dtype=np.int32); nf = C.c_longlong(0)
    hrb = np.empty((nrep, p)) if rb else None
    rc = lib().orc_coupled_ccpf_batch(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_uint64(seed), C.c_uint32(first_replicate), C.c_int(nrep),
                                      C.c_int(k), C.c_int(K), C.c_int(max_iterations), C.c_int(int(with_as)), C.c_int(nthreads),
                                      _p(hbar), _p(hrb), _p(mt), _p(it), C.byref(nf))
    if rc:
        raise RuntimeError("oracle coupled_ccpf_batch failed (%d)" % rc)
    out = dict(hbar=hbar, meetingtime=mt, iteration=it, nfilters=nf.value)
    if rb:
        out["hbar_rb"] = hrb
    return out


def rng_transform(u):
    """The RNG specification's log and sin/cos(2 pi u) (oracle/rngmath.h) applied elementwise."""
    u = _f64(np.ravel(u))
    lg, sn, cs = np.empty_like(u), np.empty_like(u), np.empty_like(u)
    lib().orc_rng_transform(_p(u), C.c_int64(u.size), _p(lg), _p(sn), _p(cs))
    return lg, sn, cs


class ApfOut(C.Structure):
    _fields_ = [("path", C.c_void_p), ("pf_means", C.c_void_p), ("loglik_t", C.c_void_p), ("resampled", C.c_void_p), ("ancestors", C.c_void_p),
                ("loglik", C.c_double), ("path_index", C.c_int), ("clamp_count", C.c_int)]


def pf_adaptive(cfg, obs, seed=0, filter_id=0, ess_threshold=1.0, systematic=False, want_ancestors=False):
    """pf(y, theta, model, nparticles, ess_threshold, systematic) (R/CPF.R:106-155) for the hot-path models."""
    T, d, N = cfg.nobs, cfg.dim, cfg.nparticles
    ob = np.asfortranarray(np.asarray(obs, dtype=np.float64).reshape(T, -1))
    path = np.empty((d, T + 1), order="F"); means = np.empty((T, d)); llt = np.empty(T); res = np.empty(T, dtype=np.int32)
    anc = np.empty((T, N), dtype=np.int32) if want_ancestors else None
    o = ApfOut()
    o.path, o.pf_means, o.loglik_t, o.resampled = path.ctypes.data, means.ctypes.data, llt.ctypes.data, res.ctypes.data
    o.ancestors = anc.ctypes.data if want_ancestors else None
    rc = lib().orc_pf_adaptive(C.byref(cfg), ob.ctypes.data_as(C.c_void_p), C.c_uint64(seed), C.c_uint64(filter_id), C.c_double(ess_threshold),
                               C.c_int(int(systematic)), C.byref(o))
    if rc:
        raise RuntimeError("oracle pf_adaptive failed (%d)" % rc)
    out = dict(path=np.ascontiguousarray(path.T), pf_means=means, loglik=float(o.loglik), loglik_t=llt, resampled=res, path_index=int(o.path_index),
               clamp_count=int(o.clamp_count))
    if want_ancestors:
        out["ancestors"] = anc
    return out
