# -*- coding: utf-8 -*-
"""Linear state-space form of the PSMF step with an observation selector H (ExperimentChange/PSMF.m:6-45).

The change-point experiment of the reference (Matlab) runs PSMF with a linear-Gaussian state
``x_t = A x_{t-1} + noise`` of dimension s = 2 r (a Matern-3/2 GP per latent factor) of which only the r positions
selected by ``H`` (r x s) enter the observation: ``y_t ~ C H x_t``.  Every place where the Matlab code uses ``C`` it is
multiplied by ``H`` and every place where it uses ``V`` it is sandwiched as ``H' V H`` (PSMF.m:15-24,32-41), so the step
is EXACTLY the standard PSMF step of rank s with

    C_eff = C H   (m x s),     V_eff = H' V H   (s x s),     dynamics x_bar = A x  (PSMF_DYN_LINEAR)

-- the rank-1 update ``C += e (x_bar' H' V) / N`` keeps ``C_eff = C H`` of that form, and ``V_eff`` keeps the form
``H' V H``.  This module does that embedding on top of the CUDA engine and maps the result back (``C = C_eff H'`` for a
selector with orthonormal rows).  One deliberate deviation: PSMF.m:16,32 adds the scalar x'H'VHx to EVERY entry of R (Matlab broadcasting of matrix +
scalar); the model -- and the reference's Python implementations (psmf.py:141-143, rPSMF.py:92-98) -- add it to the
diagonal, R_bar = R + (x'Vx) I.  This module follows the model.  The engine supports ranks up to 16, i.e. r <= 8 here (the experiment's r = 10 would need
s = 20).  The reference for this file is Matlab and cannot run in this repository's environment: parity is checked against
a line-by-line numpy restatement of PSMF.m in the tests ("parity unpinned" beyond that).
"""

from __future__ import annotations

import numpy as np
import torch

from . import _capi
from .engine import FilterEngine


def embed_selector(C, V, H):
    """(C, V, H) -> (C_eff, V_eff) = (C H, H' V H)."""
    C, V, H = (np.asarray(a, dtype=np.float64) for a in (C, V, H))
    return C @ H, H.T @ V @ H


def PSMF(r, Y, Q, A, R, H, V, P0, C, X, m, n, x0=None, device=None):
    """``X = PSMF(r, Y, Q, A, R, H, V, P0, C, X, m, n)`` of ExperimentChange/PSMF.m:6 on the GPU.

    Y (m, n); Q, A, P0 (2r, 2r); R (m, m) = rho * I; H (r, 2r); V (r, r); C (m, r); X (2r, n) is overwritten column by
    column with the filtered states and returned.  The Matlab code draws its initial state at random
    (``X0 = chol(Q)' * randn``, PSMF.m:12); pass it as ``x0`` (default zeros).  Also returns nothing else -- like the
    Matlab function -- but the filtered dictionary is available as ``PSMF.last["C"]``."""
    Y = np.asarray(Y, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    H = np.asarray(H, dtype=np.float64)
    s = A.shape[0]
    if s > _capi_max_rank():
        raise NotImplementedError("state dimension %d exceeds the engine's maximal rank %d (r <= %d with a 2r-dim state)"
                                  % (s, _capi_max_rank(), _capi_max_rank() // 2))
    if H.shape != (r, s) or Y.shape != (m, n):
        raise ValueError("H must be (r, 2r) and Y (m, n)")
    if not np.allclose(H @ H.T, np.eye(r), atol=1e-12):
        raise NotImplementedError("H must have orthonormal rows (a selector), as in ExperimentChange/main.m:70")
    Rm = np.asarray(R, dtype=np.float64)
    rho = float(Rm) if Rm.ndim == 0 else float(Rm[0, 0])
    if Rm.ndim == 2 and (np.count_nonzero(Rm - rho * np.eye(m)) != 0):
        raise NotImplementedError("R must be rho * I (main.m:79)")
    C_eff, V_eff = embed_selector(C, V, H)
    eng = FilterEngine(m, s, robust=False, c_update_transpose=False, dynamics=_capi.DYN_LINEAR, device=device)
    try:
        eng.set_linear_dynamics(A)
        eng.set_state(C_=C_eff, V=V_eff, P=P0, x=np.zeros(s) if x0 is None else np.asarray(x0, dtype=np.float64).reshape(s), Q=Q,
                      rho=[rho], lam=[0.0])
        Yt = torch.as_tensor(np.ascontiguousarray(Y.T)).to(eng.device)
        out = eng.run(Yt, None, k0=1, want_X=True)
        bad = eng.status()
        if bad >= 0:
            raise FloatingPointError("non-finite filter state at step %d" % bad)
        X[:, :] = out["X"].cpu().numpy().T
        st = eng.get_state()
        PSMF.last = dict(C=st["C"].cpu().numpy() @ H.T, V=H @ st["V"].cpu().numpy() @ H.T, P=st["P"].cpu().numpy())
    finally:
        eng.close()
    return X


def _capi_max_rank():
    return 16
