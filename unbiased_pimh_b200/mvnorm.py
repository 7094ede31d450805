"""fast_dmvnorm* / fast_rmvnorm* (R/util_dmvnorm.R, R/util_dmvnorm_transpose.R, R/util_rmvnorm.R, R/util_rmvnorm_transpose.R
over src/mvnorm.cpp:6-132) and lorenz_transition (R/util_lorenz_transition.R over src/lorenz96.cpp:65-93)."""
import ctypes as C

import numpy as np

from . import _lib as L
from . import rng


def fast_dmvnorm(x, mean, covariance):
    """x: n x d."""
    xx = np.asfortranarray(np.asarray(x, dtype=np.float64)); n, d = xx.shape
    cv = np.asfortranarray(np.asarray(covariance, dtype=np.float64)); m = L.f64(mean); out = np.empty(n)
    L.check(L.lib().cpimh_dmvnorm(xx.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int(d), L.ptr(m), cv.ctypes.data_as(C.c_void_p), L.ptr(out)))
    return out


def fast_dmvnorm_transpose(x, mean, covariance):
    """x: d x n."""
    xx = np.asfortranarray(np.asarray(x, dtype=np.float64)); d, n = xx.shape
    cv = np.asfortranarray(np.asarray(covariance, dtype=np.float64)); m = L.f64(mean); out = np.empty(n)
    L.check(L.lib().cpimh_dmvnorm_transpose(xx.ctypes.data_as(C.c_void_p), C.c_int(d), C.c_int64(n), L.ptr(m), cv.ctypes.data_as(C.c_void_p), L.ptr(out)))
    return out


def fast_dmvnorm_transpose_cholesky(x, mean, cholcovariance):
    """x: d x n; cholcovariance: lower Cholesky factor (t(chol(V)) in R)."""
    xx = np.asfortranarray(np.asarray(x, dtype=np.float64)); d, n = xx.shape
    ch = np.asfortranarray(np.asarray(cholcovariance, dtype=np.float64)); m = L.f64(mean); out = np.empty(n)
    L.check(L.lib().cpimh_dmvnorm_transpose_cholesky(xx.ctypes.data_as(C.c_void_p), C.c_int(d), C.c_int64(n), L.ptr(m), ch.ctypes.data_as(C.c_void_p), L.ptr(out)))
    return out


def _z(z, order_shape):
    return None if z is None else np.asfortranarray(np.asarray(z, dtype=np.float64).reshape(order_shape))


def fast_rmvnorm(nparticles, mean, covariance, z=None, seed=None):
    """Returns n x d samples.  z (n x d standard normals) may be injected."""
    cv = np.asfortranarray(np.asarray(covariance, dtype=np.float64)); d = cv.shape[0]; m = L.f64(mean)
    zz = _z(z, (nparticles, d)); out = np.empty((nparticles, d), order="F")
    L.check(L.lib().cpimh_rmvnorm(C.c_int64(nparticles), C.c_int(d), L.ptr(m), cv.ctypes.data_as(C.c_void_p), C.c_uint64(rng.next_seed() if seed is None else seed),
                                  None if zz is None else zz.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    return np.ascontiguousarray(out)


def fast_rmvnorm_transpose(nparticles, mean, covariance, z=None, seed=None):
    """Returns d x n samples."""
    cv = np.asfortranarray(np.asarray(covariance, dtype=np.float64)); d = cv.shape[0]; m = L.f64(mean)
    zz = _z(z, (d, nparticles)); out = np.empty((d, nparticles), order="F")
    L.check(L.lib().cpimh_rmvnorm_transpose(C.c_int64(nparticles), C.c_int(d), L.ptr(m), cv.ctypes.data_as(C.c_void_p), C.c_uint64(rng.next_seed() if seed is None else seed),
                                            None if zz is None else zz.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    return np.ascontiguousarray(out)


def fast_rmvnorm_transpose_cholesky(nparticles, mean, cholcovariance, z=None, seed=None):
    ch = np.asfortranarray(np.asarray(cholcovariance, dtype=np.float64)); d = ch.shape[0]; m = L.f64(mean)
    zz = _z(z, (d, nparticles)); out = np.empty((d, nparticles), order="F")
    L.check(L.lib().cpimh_rmvnorm_transpose_cholesky(C.c_int64(nparticles), C.c_int(d), L.ptr(m), ch.ctypes.data_as(C.c_void_p),
                                                     C.c_uint64(rng.next_seed() if seed is None else seed),
                                                     None if zz is None else zz.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    return np.ascontiguousarray(out)


def lorenz_transition(xparticles, time_start, time_end, dt, parameters, integrator="dopri5", substeps=1):
    """lorenz_transition(xparticles, time_start, time_end, dt, parameters) -> one_step_lorenz_vector."""
    x = np.asfortranarray(np.asarray(xparticles, dtype=np.float64)); d, n = x.shape
    F = float(np.ravel(parameters)[0]); out = np.empty((d, n), order="F")
    integ = {"rk4": L.INTEGRATOR_RK4, "dopri5": L.INTEGRATOR_DOPRI5}[integrator]
    L.check(L.lib().cpimh_lorenz_transition(x.ctypes.data_as(C.c_void_p), C.c_int(d), C.c_int64(n), C.c_double(time_start), C.c_double(time_end),
                                            C.c_double(dt), C.c_double(F), C.c_int(integ), C.c_int(int(substeps)), out.ctypes.data_as(C.c_void_p)))
    return np.ascontiguousarray(out)
