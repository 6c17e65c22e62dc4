# -*- coding: utf-8 -*-
"""Drop-in replacements for the two flat model functions of the imputation experiment.

    robust_PSMF                                  ExperimentImpute/rPSMF.py:39-148
    ProbabilisticSequentialMatrixFactorizer      ExperimentImpute/PSMF.py:39-95

Same positional signatures, same return tuple ``(Epred, Efull, RunTime, InsideBars)``,
same side effect (``X`` is overwritten in place with the filtered x_t, rPSMF.py:104;
``C`` is left untouched).  The per-timestep loop bodies (rPSMF.py:81-135, PSMF.py:60-84)
run on the GPU through libpsmf_b200.so; the d x d temporaries of the reference are gone.
"""

from __future__ import annotations

import time

import numpy as np
import torch

from . import _capi
from .engine import FilterEngine


def _uniform_diag(R, d, name):
    """The CUDA path supports R = rho * I only (every experiment of the reference uses that)."""
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 0:
        return float(R)
    if R.shape != (d, d):
        raise ValueError("%s must be (d, d)" % name)
    dg = np.diagonal(R)
    if np.count_nonzero(R - np.diag(dg)) != 0:
        raise NotImplementedError("%s must be diagonal (the reference itself assumes this, rPSMF.py:91)" % name)
    if not np.all(dg == dg[0]):
        raise NotImplementedError("%s must be rho * I: non-uniform diagonals are not supported" % name)
    return float(dg[0])


def RMSEM(Y1, Y2, M):
    """common.py:79-84 on device tensors."""
    n = M.sum()
    return torch.sqrt((((Y1 - Y2) * M) ** 2).sum() / n)


def _fit(Y, C, X, d, n, r, M, Mmiss, V, Q0, R0, P, lambda0, sig, Iter, YorigInt, Einit, robust, device=None,
         dtype=torch.float64, return_details=False):
    Y = np.asarray(Y); M = np.asarray(M)
    if Y.shape != (d, n) or M.shape != (d, n):
        raise ValueError("Y and M must be (d, n)")
    if X.shape != (r, n) or np.asarray(C).shape != (d, r):
        raise ValueError("C must be (d, r) and X (r, n)")
    rho0 = _uniform_diag(R0, d, "R")
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    f64 = torch.float64

    Epred = np.zeros([1, Iter + 1]); Efull = np.zeros([1, Iter + 1])
    Epred[:, 0] = Einit; Efull[:, 0] = Einit
    RunTime = np.zeros([1, Iter + 1])
    t0 = time.time()

    # time-major device copies: y_t and m_t contiguous (the reference gathers a strided column per step)
    Yt = torch.as_tensor(np.ascontiguousarray(Y.T), dtype=dtype).to(dev)
    Mt = torch.as_tensor(np.ascontiguousarray(M.T != 0).astype(np.uint8)).to(dev)
    Yo = torch.as_tensor(np.ascontiguousarray(np.asarray(YorigInt).T), dtype=f64).to(dev)
    Mm = torch.as_tensor(np.ascontiguousarray(np.asarray(Mmiss).T), dtype=f64).to(dev)

    eng = FilterEngine(d, r, dtype=dtype, robust=robust, c_update_transpose=robust, dynamics=_capi.DYN_IDENTITY,
                       device=dev.index)
    try:
        eng.set_state(C_=C, V=V, P=P, x=np.ascontiguousarray(X[:, n - 1]), Q=Q0, rho=[rho0], lam=[lambda0 if robust else 0.0])
        out = None
        for i in range(Iter):
            if robust:
                eng.set_state(Q=Q0, rho=[rho0], lam=[lambda0])                 # rPSMF.py:77-79
            # x_bar of t = 0 wraps to X[:, n-1] (rPSMF.py:86): that is the engine's carried x
            out = eng.run(Yt, Mt, k0=1, want_X=True, want_Yrec=True, want_scal=True)
            bad = eng.status()
            Xd = out["X"]                                                       # (n, r)
            X[:, :] = Xd.T.cpu().numpy()                                        # rPSMF.py:104 (in place)
            Cd = eng.get_state()["C"].to(f64)
            Yrec = out["Yrec"].to(f64)                                          # (n, d)
            Yrec2 = Xd @ Cd.T                                                   # (C @ X).T, rPSMF.py:137
            Epred[:, i + 1] = float(RMSEM(Yrec, Yo, Mm))
            Efull[:, i + 1] = float(RMSEM(Yrec2, Yo, Mm))
            if bad >= 0:
                Epred[:, i + 1] = np.nan; Efull[:, i + 1] = np.nan
            RunTime[:, i + 1] = time.time() - t0
        if out is None:                                                         # Iter = 0: nothing was filtered
            return (Epred, Efull, RunTime, 0.0, {}) if return_details else (Epred, Efull, RunTime, 0.0)   # bounds are all zero
        sc = out["scal"]
        if robust:
            U = sc[:, _capi.SCAL_NAMES.index("a")].unsqueeze(1) * Mt.to(f64) + sc[:, _capi.SCAL_NAMES.index("eta")].unsqueeze(1)
        else:
            U = sc[:, _capi.SCAL_NAMES.index("N")].unsqueeze(1).expand(n, d)
        sq = sig * torch.sqrt(U)                                                # rPSMF.py:121-123 / PSMF.py:83-84
        lo, hi = Yrec - sq, Yrec + sq
        inside = ((Mm == 1) & (Yo < hi) & (lo < Yo)).sum().to(f64) / Mm.sum()   # common.py:87-94
        InsideBars = float(inside)
        if return_details:
            return Epred, Efull, RunTime, InsideBars, dict(C=Cd.cpu().numpy(), Yrec=Yrec.T.cpu().numpy(),
                                                          scal=sc.cpu().numpy(), state=eng.get_state())
    finally:
        eng.close()
    return Epred, Efull, RunTime, InsideBars


def robust_PSMF(Y, C, X, d, n, r, M, Mmiss, V, Q0, R0, P, lambda0, sig, Iter, YorigInt, Einit):
    """ExperimentImpute/rPSMF.py:39-58 signature."""
    return _fit(Y, C, X, d, n, r, M, Mmiss, V, Q0, R0, P, lambda0, sig, Iter, YorigInt, Einit, robust=True)


def ProbabilisticSequentialMatrixFactorizer(Y, C, X, d, n, r, M, Mmiss, lam, V, Q, R, P, sig, Iter, YorgInt, Einit):
    """ExperimentImpute/PSMF.py:40-42 signature (``lam`` is unused there too)."""
    return _fit(Y, C, X, d, n, r, M, Mmiss, V, Q, R, P, 0.0, sig, Iter, YorgInt, Einit, robust=False)
