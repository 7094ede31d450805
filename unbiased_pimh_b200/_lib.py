"""ctypes binding of libcpimh.so (the C ABI declared in include/cpimh.h).

There is no CPU fallback: if the CUDA library is missing this module raises at import of the first symbol,
and every compute call raises CpimhError when no usable GPU is present.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPIMH_LIB") or os.path.join(HERE, "libcpimh.so")   # CPIMH_LIB: A/B builds of the same ABI (tools/)
MAX_THETA = 32

MODEL_GSS, MODEL_AR, MODEL_LEVY_SV, MODEL_SK, MODEL_LORENZ = range(5)
INTEGRATOR_RK4, INTEGRATOR_DOPRI5 = 0, 1
OK, ERR_INVALID, ERR_CUDA, ERR_TREE_OVERFLOW, ERR_NOISE_OVERFLOW, ERR_NOMEM, ERR_NOT_MET = range(7)


class CpimhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libcpimh error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("model", C.c_int32), ("dim", C.c_int32), ("dim_y", C.c_int32), ("nparticles", C.c_int32), ("nobs", C.c_int32),
                ("shuffle", C.c_int32), ("integrator", C.c_int32), ("substeps", C.c_int32), ("sk_max_events", C.c_int32),
                ("ntheta", C.c_int32), ("theta", C.c_double * MAX_THETA),
                ("tree_capacity", C.c_int32), ("threads_per_filter", C.c_int32), ("store_paths", C.c_int32), ("reserved", C.c_int32)]


class Noise(C.Structure):
    _fields_ = [("init", C.c_void_p), ("trans", C.c_void_p), ("resample", C.c_void_p), ("shuffle", C.c_void_p), ("path_u", C.c_void_p)]


class ChainOut(C.Structure):
    _fields_ = [("hbar", C.c_void_p), ("meetingtime", C.c_void_p), ("iteration", C.c_void_p), ("sum_hbar", C.c_void_p),
                ("sum_hbar_sq", C.c_void_p), ("nfilters", C.c_void_p), ("bc", C.c_void_p),
                ("hbar_rb", C.c_void_p), ("sum_hbar_rb", C.c_void_p), ("sum_hbar_rb_sq", C.c_void_p),
                ("finished", C.c_void_p), ("nfinished", C.c_void_p)]


_lib = None


def lib():
    """Load libcpimh.so; fail loudly when it has not been built (python -m unbiased_pimh_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libcpimh.so not found at %s: build it with `python -m unbiased_pimh_b200.build` "
                              "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.cpimh_last_error.restype = C.c_char_p
        L.cpimh_pf_launch_count.restype = C.c_int64
        L.cpimh_last_exact_repeats.restype = C.c_int64
        L.cpimh_pf_device_bytes.restype = C.c_int64
        L.cpimh_chains_launch_count.restype = C.c_int64
        L.cpimh_chains_filters_launched.restype = C.c_int64
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise CpimhError(rc, lib().cpimh_last_error().decode("utf-8", "replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def make_config(model, dim, dim_y, nparticles, nobs, theta=(), shuffle=0, integrator=0, substeps=1, sk_max_events=64,
                tree_capacity=0, threads_per_filter=0, store_paths=1):
    c = Config()
    lib().cpimh_config_default(C.byref(c))
    c.model, c.dim, c.dim_y, c.nparticles, c.nobs = int(model), int(dim), int(dim_y), int(nparticles), int(nobs)
    c.shuffle, c.integrator, c.substeps, c.sk_max_events = int(shuffle), int(integrator), int(substeps), int(sk_max_events)
    c.tree_capacity, c.threads_per_filter, c.store_paths = int(tree_capacity), int(threads_per_filter), int(store_paths)
    th = [float(v) for v in np.ravel(np.asarray(theta, dtype=np.float64))]
    if len(th) > MAX_THETA:
        raise ValueError("theta has %d entries (max %d)" % (len(th), MAX_THETA))
    c.ntheta = len(th)
    for i, v in enumerate(th):
        c.theta[i] = v
    return c


def obs_matrix(observations, nobs):
    """R `observations` (T x dim_y) as a column-major float64 buffer."""
    ob = np.asarray(observations, dtype=np.float64)
    if ob.ndim == 1:
        ob = ob.reshape(nobs, -1)
    if ob.shape[0] != nobs:
        raise ValueError("observations has %d rows, expected %d" % (ob.shape[0], nobs))
    return np.asfortranarray(ob)


def make_noise(noise):
    """dict of injected arrays -> (Noise struct, keep-alive list)"""
    nz = Noise()
    keep = []
    if noise:
        for k in ("init", "trans", "resample", "shuffle", "path_u"):
            if noise.get(k) is not None:
                a = f64(noise[k])
                keep.append(a)
                setattr(nz, k, a.ctypes.data)
    return nz, keep
