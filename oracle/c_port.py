"""ctypes binding of the C restatement (oracle/c/plf_oracle.c).  Test infrastructure / CPU baseline only."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libplf_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError("%s missing: run make -C oracle/c" % LIB)
        _lib = ctypes.CDLL(LIB)
        _lib.plf_oracle_max_threads.restype = ctypes.c_int
        _lib.plf_oracle_ll_deriv.restype = ctypes.c_int
    return _lib


def max_threads():
    return load().plf_oracle_max_threads()


def ll_deriv(indptr, indices, preorder, P, D, prior, root_mode, root_vec, codes, defs, w=None,
             nthreads=0, want_deriv=True, want_site_ll=True):
    lib = load()
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    preorder = np.ascontiguousarray(preorder, dtype=np.int32)
    P = np.ascontiguousarray(P, dtype=np.float64)
    C, E, n, _ = P.shape
    N = E + 1
    D = np.ascontiguousarray(D, dtype=np.float64) if D is not None else P
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    rv = np.ascontiguousarray(root_vec if root_vec is not None else np.ones(n), dtype=np.float64)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    defs = np.ascontiguousarray(defs, dtype=np.float64)
    S = codes.shape[0]
    site_ll = np.empty(S) if want_site_ll else None
    sum_ll = np.zeros(1)
    sum_d = np.zeros(E) if want_deriv else None
    wv = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
    vp = ctypes.c_void_p

    def p(a):
        return None if a is None else a.ctypes.data_as(vp)
    rc = lib.plf_oracle_ll_deriv(ctypes.c_int(N), p(indptr), p(indices), p(preorder), ctypes.c_int(n), ctypes.c_int(C),
                                 p(P), p(D), p(prior), ctypes.c_int(root_mode), p(rv), ctypes.c_int64(S), p(codes),
                                 ctypes.c_int(defs.shape[0]), p(defs), p(wv), ctypes.c_int(nthreads),
                                 p(site_ll), p(sum_ll), p(sum_d))
    if rc != 0:
        raise RuntimeError("plf_oracle_ll_deriv failed")
    return site_ll, float(sum_ll[0]), sum_d
