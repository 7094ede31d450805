"""ctypes bindings of oracle/_ref/libref.so: the reference's OWN natives (src/systematic.cpp, multinomial.cpp,
coupledresamplingindexmatching.cpp, tree.cpp, create_A_.cpp, SVLevy_functions.cpp), compiled unmodified against the stand-in
Rcpp.h of oracle/ref_build/rcpp_stub (recipe: oracle/ref_build/Makefile).  TEST INFRASTRUCTURE: used to pin oracle/*.c, and by
`bench.py --impl reference`; the product never loads it.  R's random numbers are replaced by queues the caller fills, in the
order the reference consumes them."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref.so")
REFERENCE = os.environ.get("CPIMH_REFERENCE", "/root/reference")
_lib = None


def available():
    return os.path.exists(LIB_PATH) or os.path.isdir(os.path.join(REFERENCE, "src"))


def build(force=False):
    """Compile libref.so when the reference sources are present (the build container); elsewhere the prebuilt file is used."""
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        args = ["make", "-C", os.path.join(HERE, "ref_build"), "REF=" + REFERENCE] + (["-B"] if force else [])
        r = subprocess.run(args, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + r.stdout + r.stderr)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("oracle/_ref/libref.so is missing and %s is not available to build it" % REFERENCE)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        for f in ("ref_unif_left", "ref_underflows", "ref_oob"):
            getattr(_lib, f).restype = C.c_long
        _lib.ref_tree_new.restype = C.c_void_p
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _queue(unif=(), exp=(), pois=(), gamma=()):
    L = lib()
    L.ref_clear_queues()
    for fn, arr in (("ref_push_unif", unif), ("ref_push_exp", exp), ("ref_push_pois", pois), ("ref_push_gamma", gamma)):
        a = _f64(np.concatenate([np.ravel(_f64(x)) for x in arr]) if isinstance(arr, (list, tuple)) and len(arr) else (arr if len(arr) else []))
        if a.size:
            getattr(L, fn)(_p(a), C.c_long(a.size))


def _check(consumed_all=True):
    L = lib()
    if L.ref_underflows():
        raise RuntimeError("the reference asked for more random numbers than were queued")
    if consumed_all and L.ref_unif_left():
        raise RuntimeError("the reference left %d queued uniforms unread" % L.ref_unif_left())
    return int(L.ref_oob())


def systematic_resampling_n_(w, ndraws, u):
    """src/systematic.cpp:7-32.  Returns (ancestors 0-based, number of out-of-range reads the reference made)."""
    w = _f64(w); anc = np.empty(ndraws, dtype=np.int32)
    _queue()
    lib().ref_systematic_resampling_n(_p(w), C.c_int(len(w)), C.c_int(ndraws), C.c_double(u), _p(anc))
    return anc, _check()


def multinomial_resampling_n_(w, ndraws, u_raw, u_shuffle):
    """src/multinomial.cpp:13-32: runif(ndraws), then ndraws-1 runif(1) inside std::random_shuffle."""
    w = _f64(w); anc = np.empty(ndraws, dtype=np.int32)
    _queue(unif=[u_raw, u_shuffle])
    lib().ref_multinomial_resampling_n(_p(w), C.c_int(len(w)), C.c_int(ndraws), _p(anc))
    return anc, _check()


def coupled_multinomial_resampling_n_(w1, w2, ndraws, u_raw, u_shuffle):
    """src/multinomial.cpp:35-71 -> (ndraws, 2) 0-based."""
    w1 = _f64(w1); w2 = _f64(w2); anc = np.empty(2 * ndraws, dtype=np.int32)
    _queue(unif=[u_raw, u_shuffle])
    lib().ref_coupled_multinomial_resampling_n(_p(w1), _p(w2), C.c_int(len(w1)), C.c_int(ndraws), _p(anc))
    return anc.reshape(2, ndraws).T.copy(), _check()


def indexmatching(w1, w2, ndraws, uniforms):
    """src/coupledresamplingindexmatching.cpp:8-57; `uniforms` is the whole stream in consumption order: runif(N), then the
    nested multinomial's (ncoupled + ncoupled - 1), then the coupled multinomial's (nres + nres - 1)."""
    w1 = _f64(w1); w2 = _f64(w2); anc = np.empty(2 * ndraws, dtype=np.int32)
    _queue(unif=[uniforms])
    lib().ref_indexmatching(_p(w1), _p(w2), C.c_int(len(w1)), C.c_int(ndraws), _p(anc))
    return anc.reshape(2, ndraws).T.copy(), _check()


class Tree:
    """The reference's Tree (src/tree.cpp:6-141) behind a handle; x matrices are (d, N) numpy arrays."""

    def __init__(self, N, M, dimx):
        self.h = C.c_void_p(lib().ref_tree_new(C.c_int(N), C.c_int(M), C.c_int(dimx)))
        self.N, self.dimx = N, dimx

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_tree_free(self.h); self.h = None

    def init(self, x0):
        x = np.asfortranarray(_f64(x0)); lib().ref_tree_init(self.h, _p(x))

    def update(self, x, a):
        x = np.asfortranarray(_f64(x)); a = np.ascontiguousarray(a, dtype=np.int32)
        lib().ref_tree_update(self.h, _p(x), _p(a))

    @property
    def M(self):
        return lib().ref_tree_M(self.h)

    @property
    def nsteps(self):
        return lib().ref_tree_nsteps(self.h)

    def arrays(self):
        M = self.M
        a = np.empty(M, dtype=np.int32); o = np.empty(M, dtype=np.int32); l = np.empty(self.N, dtype=np.int32)
        x = np.empty((self.dimx, M), dtype=np.float64, order="F")
        lib().ref_tree_arrays(self.h, _p(a), _p(o), _p(l), _p(x))
        return a, o, l, np.ascontiguousarray(x)

    def get_path(self, n):
        path = np.empty((self.dimx, self.nsteps + 1), dtype=np.float64, order="F")
        lib().ref_tree_get_path(self.h, C.c_int(n), _p(path))
        return np.ascontiguousarray(path)

    def retrieve_xgeneration(self, lag):
        xg = np.empty((self.dimx, self.N), dtype=np.float64, order="F")
        lib().ref_tree_retrieve_xgeneration(self.h, C.c_int(lag), _p(xg))
        return np.ascontiguousarray(xg)


def create_A(alpha, d):
    A = np.empty((d, d), dtype=np.float64, order="F")
    lib().ref_create_A(C.c_double(alpha), C.c_int(d), _p(A))
    return np.ascontiguousarray(A)


def rtransition_SVLevy(Xold, t_1, t, xi, w2, lam, k, c_unif, e_unit):
    """src/SVLevy_functions.cpp:58-70.  Xold (N, 2) = (v, z); k = the rpois counts, c_unif / e_unit = the uniforms and unit
    exponentials of all jumps in particle order."""
    X = np.asfortranarray(_f64(Xold)); N = X.shape[0]; out = np.empty((N, 2), dtype=np.float64, order="F")
    _queue(unif=[c_unif] if len(c_unif) else (), exp=[e_unit] if len(e_unit) else (), pois=[k])
    lib().ref_rtransition_SVLevy(_p(X), C.c_int(N), C.c_double(t_1), C.c_double(t), C.c_double(xi), C.c_double(w2), C.c_double(lam), _p(out))
    _check()
    return np.ascontiguousarray(out)


def rinitial_SVLevy(N, t, xi, w2, lam, z0, k, c_unif, e_unit):
    """src/SVLevy_functions.cpp:40-55; z0 = the rgamma draws."""
    out = np.empty((N, 2), dtype=np.float64, order="F")
    _queue(unif=[c_unif] if len(c_unif) else (), exp=[e_unit] if len(e_unit) else (), pois=[k], gamma=[z0])
    lib().ref_rinitial_SVLevy(C.c_int(N), C.c_double(t), C.c_double(xi), C.c_double(w2), C.c_double(lam), _p(out))
    _check()
    return np.ascontiguousarray(out)


def dobs_SVLevy(y, Xts, mu, beta, log=True):
    """src/SVLevy_functions.cpp:73-81 (dnorm itself is the stub's restatement of R's nmath)."""
    X = np.asfortranarray(_f64(Xts)); N = X.shape[0]; out = np.empty(N, dtype=np.float64)
    _queue()
    lib().ref_dobs_SVLevy(C.c_double(y), _p(X), C.c_int(N), C.c_double(mu), C.c_double(beta), C.c_int(int(log)), _p(out))
    return out
