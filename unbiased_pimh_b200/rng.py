"""Host-side seed source.  The reference draws from R's global RNG (set.seed); the CUDA kernels use Philox streams,
so every call that needs randomness takes one 64-bit seed from here (the R shim takes it from
unif_rand(), so set.seed() still makes runs repeatable)."""
import numpy as np

_rng = np.random.default_rng(17)


def set_seed(seed):
    global _rng
    _rng = np.random.default_rng(int(seed))


def next_seed():
    return int(_rng.integers(0, 2 ** 63 - 1))


def runif(n=1):
    u = _rng.random(n)
    return np.where(u <= 0.0, 2.0 ** -53, u)


def rng_transform(u):
    """log(u), sin(2 pi u), cos(2 pi u) by the RNG specification's device functions (cpimh_rng_transform): the maps that turn
    Philox uniforms into exponential variates and Box-Muller normals inside the kernels."""
    import ctypes as C
    import numpy as np
    from . import _lib as L
    u = L.f64(np.ravel(u))
    lg, sn, cs = np.empty_like(u), np.empty_like(u), np.empty_like(u)
    L.check(L.lib().cpimh_rng_transform(L.ptr(u), C.c_int64(u.size), L.ptr(lg), L.ptr(sn), L.ptr(cs)))
    return lg, sn, cs
