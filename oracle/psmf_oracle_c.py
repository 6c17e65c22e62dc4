# -*- coding: utf-8 -*-
"""ctypes wrapper of oracle/libpsmf_oracle.so (C/OpenMP restatement of the step) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpsmf_oracle.so")


def available():
    return os.path.exists(_LIB)


def _lib():
    L = C.CDLL(_LIB)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.psmf_oracle_run.restype = C.c_int64
    L.psmf_oracle_run.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, C.c_void_p, C.c_int64,
                                  C.c_void_p]
    L.psmf_oracle_threads.restype = C.c_int
    return L


def threads():
    return int(_lib().psmf_oracle_threads())


def use_all_cores():
    """Use every core this process may run on (torchrun exports OMP_NUM_THREADS=1); returns the thread count."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    L = _lib()
    L.psmf_oracle_set_threads(C.c_int(n))
    return int(L.psmf_oracle_threads())


def run(C_, x, P, V, Q, rho, lam, Y, M, robust=True, cupdate_vt=True, want_X=True):
    """Arrays are copied; returns dict(C, x, P, V, Q, rho, lam, X, bad)."""
    L = _lib()
    Cc = np.ascontiguousarray(C_, dtype=np.float64).copy()
    d, r = Cc.shape
    xs, Ps, Vs, Qs = (np.ascontiguousarray(a, dtype=np.float64).copy() for a in (x, P, V, Q))
    scal = np.array([rho, lam], dtype=np.float64)
    Yc = np.ascontiguousarray(Y, dtype=np.float64)
    T = Yc.shape[0]
    Mc = None if M is None else np.ascontiguousarray(M, dtype=np.uint8)
    X = np.zeros((T, r)) if want_X else None
    bad = L.psmf_oracle_run(d, r, int(robust), int(cupdate_vt), Cc, xs, Ps, Vs, Qs, scal, Yc,
                            None if Mc is None else Mc.ctypes.data_as(C.c_void_p), T,
                            None if X is None else X.ctypes.data_as(C.c_void_p))
    return dict(C=Cc, x=xs, P=Ps, V=Vs, Q=Qs, rho=float(scal[0]), lam=float(scal[1]), X=X, bad=int(bad))
