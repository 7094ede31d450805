"""Resamplers with the reference's names and index conventions.

`*_` functions are the natives (0-based, like src/*.cpp); the wrappers add 1 exactly where the R wrappers do
(R/resampling-multinomial.R:3-5, R/resampling-systematic.R:8-10, R/coupledresampling-systematic.R:6-12).
Reference quirks handled explicitly: `CR_multinomial` returns 0-based indices in the reference
(R/coupledresampling-multinomial.R:8-9); here it returns 1-based like every other R-level resampler, and
`CR_indexmatching` (referenced at R/pf_kernels.R:238 but never defined in the package) is provided.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import rng


def _batch(w):
    w = L.f64(w)
    single = w.ndim == 1
    return (w.reshape(1, -1) if single else w), single


def sorted_uniforms(u_raw):
    """Ascending uniforms by the recurrence of src/multinomial.cpp:20-24 (GPU suffix scan)."""
    u, single = _batch(u_raw)
    e = np.empty_like(u)
    L.check(L.lib().cpimh_sorted_uniforms(L.ptr(u), C.c_int64(u.shape[0]), C.c_int(u.shape[1]), L.ptr(e)))
    return e[0] if single else e


def systematic_resampling_n_(weights, ndraws, u):
    """src/systematic.cpp:7-32.  weights (N,) or (B, N); u scalar or (B,).  0-based."""
    w, single = _batch(weights)
    uu = L.f64(np.broadcast_to(np.asarray(u, dtype=np.float64), (w.shape[0],)))
    anc = np.empty((w.shape[0], ndraws), dtype=np.int32)
    L.check(L.lib().cpimh_systematic_resampling_n(L.ptr(w), C.c_int64(w.shape[0]), C.c_int(w.shape[1]), C.c_int(ndraws), L.ptr(uu), L.ptr(anc)))
    return anc[0] if single else anc


def multinomial_resampling_n_(weights, ndraws, e_sorted=None, u_shuffle=None, shuffle=True):
    """src/multinomial.cpp:13-32.  Uniforms are injected (`e_sorted` ascending in (0,1), `u_shuffle` for the
    random_shuffle) or drawn from the host seed source.  0-based."""
    w, single = _batch(weights)
    B = w.shape[0]
    if e_sorted is None:
        e_sorted = sorted_uniforms(rng.runif(B * ndraws).reshape(B, ndraws))
        if shuffle and u_shuffle is None and ndraws > 1:
            u_shuffle = rng.runif(B * (ndraws - 1)).reshape(B, ndraws - 1)
    e = L.f64(e_sorted).reshape(B, ndraws)
    us = None if u_shuffle is None else L.f64(u_shuffle).reshape(B, max(ndraws - 1, 0))
    anc = np.empty((B, ndraws), dtype=np.int32)
    L.check(L.lib().cpimh_multinomial_resampling_n(L.ptr(w), C.c_int64(B), C.c_int(w.shape[1]), C.c_int(ndraws), L.ptr(e), L.ptr(us), L.ptr(anc)))
    return anc[0] if single else anc


def coupled_multinomial_resampling_n_(weights1, weights2, ndraws, e_sorted=None, u_shuffle=None, shuffle=True):
    """src/multinomial.cpp:35-71.  Returns (ndraws, 2) (or (B, ndraws, 2)), 0-based."""
    w1, single = _batch(weights1)
    w2, _ = _batch(weights2)
    B = w1.shape[0]
    if e_sorted is None:
        e_sorted = sorted_uniforms(rng.runif(B * ndraws).reshape(B, ndraws))
        if shuffle and u_shuffle is None and ndraws > 1:
            u_shuffle = rng.runif(B * (ndraws - 1)).reshape(B, ndraws - 1)
    e = L.f64(e_sorted).reshape(B, ndraws)
    us = None if u_shuffle is None else L.f64(u_shuffle).reshape(B, max(ndraws - 1, 0))
    anc = np.empty((B, 2, ndraws), dtype=np.int32)
    L.check(L.lib().cpimh_coupled_multinomial_resampling_n(L.ptr(w1), L.ptr(w2), C.c_int64(B), C.c_int(w1.shape[1]), C.c_int(ndraws),
                                                           L.ptr(e), L.ptr(us), L.ptr(anc)))
    out = anc.transpose(0, 2, 1)
    return np.ascontiguousarray(out[0] if single else out)


def indexmatching_cpp(w1, w2, ndraws, u_couple=None, e_mult=None, us_mult=None, e_cm=None, us_cm=None, shuffle=True):
    """src/coupledresamplingindexmatching.cpp:8-57.  Returns (ndraws, 2), 0-based."""
    a1, single = _batch(w1)
    a2, _ = _batch(w2)
    B, N = a1.shape
    if u_couple is None:
        u_couple = rng.runif(B * N).reshape(B, N)
    if e_mult is None or e_cm is None:
        # the nested resamplers draw runif(ncoupled) / runif(ndraws - ncoupled) (:36-42): size the sorted-uniform
        # ladders for the actual counts.  alpha is the same left-to-right fp64 sum the kernel forms (:13-16).
        alpha = np.cumsum(np.minimum(a1, a2), axis=1)[:, -1]
        nc = (np.reshape(u_couple, (B, N))[:, :ndraws] < alpha[:, None]).sum(axis=1)
        e_mult = np.full((B, ndraws), 0.5)
        e_cm = np.full((B, ndraws), 0.5)
        for b in range(B):
            if nc[b] > 0:
                e_mult[b, :nc[b]] = sorted_uniforms(rng.runif(int(nc[b])))
            if ndraws - nc[b] > 0:
                e_cm[b, :ndraws - nc[b]] = sorted_uniforms(rng.runif(int(ndraws - nc[b])))
        if shuffle and ndraws > 1:
            if us_mult is None:
                us_mult = rng.runif(B * (ndraws - 1)).reshape(B, ndraws - 1)
            if us_cm is None:
                us_cm = rng.runif(B * (ndraws - 1)).reshape(B, ndraws - 1)
    arrs = [L.f64(np.reshape(u_couple, (B, N))), L.f64(np.reshape(e_mult, (B, ndraws))),
            None if us_mult is None else L.f64(np.reshape(us_mult, (B, ndraws - 1))),
            L.f64(np.reshape(e_cm, (B, ndraws))), None if us_cm is None else L.f64(np.reshape(us_cm, (B, ndraws - 1)))]
    anc = np.empty((B, 2, ndraws), dtype=np.int32)
    L.check(L.lib().cpimh_indexmatching(L.ptr(a1), L.ptr(a2), C.c_int64(B), C.c_int(N), C.c_int(ndraws), *[L.ptr(a) for a in arrs], L.ptr(anc)))
    out = anc.transpose(0, 2, 1)
    return np.ascontiguousarray(out[0] if single else out)


# ---------------------------------------------------------------- R-level wrappers (1-based)
def multinomial_resampling_n(normalized_weights, N):
    return multinomial_resampling_n_(normalized_weights, N) + 1


def systematic_resampling_n(normalized_weights, N, u):
    return systematic_resampling_n_(normalized_weights, N, u) + 1


def CR_systematic(xparticles1, xparticles2, normweights1, normweights2, *args):
    u = float(rng.runif(1)[0])
    n = np.asarray(xparticles1).shape[1]
    return np.stack([systematic_resampling_n(normweights1, n, u), systematic_resampling_n(normweights2, n, u)], axis=1)


def CR_multinomial(xparticles1, xparticles2, normweights1, normweights2, *args):
    n = np.asarray(xparticles1).shape[1]
    return coupled_multinomial_resampling_n_(normweights1, normweights2, n) + 1


def CR_indexmatching(xparticles1, xparticles2, normweights1, normweights2, ntrials=None, *args):
    n = np.asarray(xparticles1).shape[1] if ntrials is None else int(ntrials)
    return indexmatching_cpp(normweights1, normweights2, n) + 1


def normalize_weight(logw, return_logz=False):
    """R/util_normalize_weight.R:5-9 (and the log Z increment of R/CPF.R:66 when return_logz)."""
    lw, single = _batch(logw)
    w = np.empty_like(lw)
    lz = np.empty(lw.shape[0])
    L.check(L.lib().cpimh_normalize_weight(L.ptr(lw), C.c_int64(lw.shape[0]), C.c_int(lw.shape[1]), L.ptr(w), L.ptr(lz)))
    if return_logz:
        return (w[0], float(lz[0])) if single else (w, lz)
    return w[0] if single else w
