"""`Tree` -- the ancestor-tree path storage exposed by the reference as the Rcpp module class `module_tree$Tree`
(src/tree.h:6-39, src/tree.cpp:6-162): fields N, M, dimx, nsteps, a_star, o_star, x_star, l_star and methods init,
insert/prune (through update), update, get_path, retrieve_xgeneration.  Device-resident (cpimh_tree_*)."""
import ctypes as C

import numpy as np

from . import _lib as L


class Tree:
    def __init__(self, N, M, dimx):
        self.h = C.c_void_p()
        L.check(L.lib().cpimh_tree_create(C.c_int(int(N)), C.c_int(int(M)), C.c_int(int(dimx)), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            L.lib().cpimh_tree_destroy(self.h)
            self.h = None

    __del__ = close

    def _fields(self):
        v = [C.c_int() for _ in range(4)]
        L.check(L.lib().cpimh_tree_fields(self.h, *[C.byref(x) for x in v]))
        return [x.value for x in v]

    N = property(lambda self: self._fields()[0])
    M = property(lambda self: self._fields()[1])
    dimx = property(lambda self: self._fields()[2])
    nsteps = property(lambda self: self._fields()[3])

    def reset(self):
        L.check(L.lib().cpimh_tree_reset(self.h))

    def init(self, x_0):
        x = np.asfortranarray(np.asarray(x_0, dtype=np.float64))
        L.check(L.lib().cpimh_tree_init(self.h, x.ctypes.data_as(C.c_void_p)))

    def update(self, x, a):
        """x: d x N matrix, a: 0-based ancestors (the reference passes `ancestors - 1`, R/CPF.R:65)."""
        xx = np.asfortranarray(np.asarray(x, dtype=np.float64))
        aa = L.i32(a)
        L.check(L.lib().cpimh_tree_update(self.h, xx.ctypes.data_as(C.c_void_p), L.ptr(aa)))

    def get_path(self, n):
        N, M, d, ns = self._fields()
        path = np.empty((d, ns + 1), order="F")
        L.check(L.lib().cpimh_tree_get_path(self.h, C.c_int(int(n)), path.ctypes.data_as(C.c_void_p)))
        return np.ascontiguousarray(path)

    def retrieve_xgeneration(self, lag):
        N, M, d, ns = self._fields()
        xg = np.empty((d, N), order="F")
        L.check(L.lib().cpimh_tree_retrieve_xgeneration(self.h, C.c_int(int(lag)), xg.ctypes.data_as(C.c_void_p)))
        return np.ascontiguousarray(xg)

    def _arrays(self):
        N, M, d, ns = self._fields()
        a = np.empty(M, dtype=np.int32); o = np.empty(M, dtype=np.int32); l = np.empty(N, dtype=np.int32); x = np.empty((d, M), order="F")
        L.check(L.lib().cpimh_tree_arrays(self.h, L.ptr(a), L.ptr(o), L.ptr(l), x.ctypes.data_as(C.c_void_p)))
        return a, o, l, np.ascontiguousarray(x)

    a_star = property(lambda self: self._arrays()[0])
    o_star = property(lambda self: self._arrays()[1])
    l_star = property(lambda self: self._arrays()[2])
    x_star = property(lambda self: self._arrays()[3])
