"""Model lists of the reference (`get_gss`, `get_ar`, `get_levy_sv`, `get_stochastic_kinetic`, `get_lorenz`).

In the reference each `get_*()` returns an R list of closures (rinit, rinit_rand, rtransition, rtransition_rand,
dmeasurement, precompute, dimension) -- R/model_get_gss.R:43-46, R/model_get_ar.R:46-50, R/model_levy_sv.R:80-83,
R/model_stochastic_kinetic.R:159-163, R/model_get_lorenz.R:47-51.  Here the list becomes a `Model` object that also
carries a `name` (the reference lists have none, SURVEY D-table/8b) so the filter can dispatch to the CUDA model of
that name; the closures are kept and run the same CUDA device code one call at a time (cpimh_model_*).
Particle matrices are numpy arrays of shape (d, N) like R's d x N matrices.
"""
import ctypes as C

import numpy as np

from . import _lib as L


class Model:
    """A model list.  Fields mirror the reference's list entries; `name` and `pack_theta` are additive."""

    def __init__(self, name, model_id, dimension, dimension_y, pack_theta, extra=None):
        self.name = name
        self.model_id = model_id
        self.dimension = dimension
        self.dimension_y = dimension_y
        self._pack = pack_theta
        self.extra = extra or {}

    def pack_theta(self, theta):
        return self._pack(theta)

    def config(self, nparticles, theta, nobs=1, **kw):
        return L.make_config(self.model_id, self.dimension, self.dimension_y, nparticles, nobs, self.pack_theta(theta),
                             integrator=self.extra.get("integrator", 0), substeps=self.extra.get("substeps", 1), **kw)

    # ---- the reference's closures, served by the CUDA model code
    def precompute(self, theta):
        if self.name == "ar":
            A = np.empty((self.dimension, self.dimension), order="F")
            L.check(L.lib().cpimh_create_A(C.c_double(float(theta[0])), C.c_int(self.dimension), A.ctypes.data_as(C.c_void_p)))
            return {"A": np.ascontiguousarray(A), "di": np.eye(self.dimension)}
        return {}

    def rinit_rand(self, nparticles, theta):
        """Randomness is generated inside the CUDA kernels (Philox); kept for signature compatibility."""
        return None

    def rtransition_rand(self, nparticles, theta):
        return None

    def rinit(self, nparticles, theta, rand=None, precomputed=None, seed=0, filter_id=0):
        cfg = self.config(nparticles, theta)
        x = np.empty((self.dimension, nparticles), order="F")
        nz = None if rand is None else L.f64(rand)
        L.check(L.lib().cpimh_model_rinit(C.byref(cfg), L.ptr(nz), C.c_uint64(seed), C.c_uint64(filter_id), x.ctypes.data_as(C.c_void_p)))
        return np.ascontiguousarray(x)

    def rtransition(self, xparticles, theta, time, rand=None, precomputed=None, seed=0, filter_id=0, sk_max_events=64):
        x = np.asfortranarray(np.asarray(xparticles, dtype=np.float64)).copy(order="F")
        cfg = self.config(x.shape[1], theta, nobs=max(int(time), 1), sk_max_events=sk_max_events)
        nz = None if rand is None else L.f64(rand)
        L.check(L.lib().cpimh_model_rtransition(C.byref(cfg), C.c_int(int(time)), L.ptr(nz), C.c_uint64(seed), C.c_uint64(filter_id),
                                                 x.ctypes.data_as(C.c_void_p)))
        return np.ascontiguousarray(x)

    def dmeasurement(self, xparticles, theta, observation, precomputed=None):
        x = np.asfortranarray(np.asarray(xparticles, dtype=np.float64))
        cfg = self.config(x.shape[1], theta)
        y = L.f64(np.ravel(observation))
        out = np.empty(x.shape[1])
        L.check(L.lib().cpimh_model_dmeasurement(C.byref(cfg), x.ctypes.data_as(C.c_void_p), L.ptr(y), L.ptr(out)))
        return out

    def dtransition(self, *a, **k):
        raise NotImplementedError("dtransition is only used by ancestor sampling (with_as=TRUE), which is outside the hot path")

    def __getitem__(self, key):      # model$dimension style access
        return getattr(self, key)


def get_gss():
    """Gordon-Salmond-Smith nonlinear growth model, d = 1 (R/model_get_gss.R:11-48); theta is unused."""
    return Model("gss", L.MODEL_GSS, 1, 1, lambda theta: [])


def get_ar(dimension):
    """Hidden auto-regressive (linear-Gaussian) model, theta = (alpha, sigma_y) (R/model_get_ar.R:11-51)."""
    return Model("ar", L.MODEL_AR, int(dimension), int(dimension), lambda theta: [float(theta[0]), float(theta[1])])


def get_levy_sv():
    """Levy-driven stochastic volatility, theta = (mu, beta, xi, omega^2, lambda), x = (v, z) (R/model_levy_sv.R:14-84)."""
    return Model("levy_sv", L.MODEL_LEVY_SV, 2, 1, lambda theta: [float(v) for v in theta[:5]])


def _pack_sk(theta):
    A = np.asarray(theta["A_obs"], dtype=np.float64).reshape(2, 4)
    return [float(theta["s2_obs"])] + [float(v) for v in np.ravel(A, order="F")]


def get_stochastic_kinetic():
    """Prokaryotic autoregulatory gene network, exact Gillespie, theta = {A_obs (2x4), s2_obs}
    (R/model_stochastic_kinetic.R:13-165)."""
    return Model("stochastic_kinetic", L.MODEL_SK, 4, 2, _pack_sk)


def get_lorenz(sigma_y, dimension=8, integrator="dopri5", substeps=1):
    """Lorenz 96, theta = (F, sigma_x) (R/model_get_lorenz.R:17-52).  The reference hard-codes d = 8 and integrates
    with Boost.odeint's adaptive Dormand-Prince (`integrator="dopri5"`); `integrator="rk4"` with `substeps` fixed steps
    and other dimensions are extensions (BASELINE config 5: d = 64, rk4)."""
    integ = {"rk4": L.INTEGRATOR_RK4, "dopri5": L.INTEGRATOR_DOPRI5}[integrator]
    return Model("lorenz", L.MODEL_LORENZ, int(dimension), int(dimension),
                 lambda theta: [float(theta[0]), float(theta[1]), float(sigma_y)], extra={"integrator": integ, "substeps": int(substeps)})


def create_A(theta, d):
    """R/util_create_A_for_ARmodel.R:2-4 -> create_A_ (src/create_A_.cpp:6-13)."""
    A = np.empty((d, d), order="F")
    L.check(L.lib().cpimh_create_A(C.c_double(float(theta)), C.c_int(int(d)), A.ctypes.data_as(C.c_void_p)))
    return np.ascontiguousarray(A)
