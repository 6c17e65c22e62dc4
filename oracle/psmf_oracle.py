# -*- coding: utf-8 -*-
"""CPU oracle for the PSMF / rPSMF per-timestep filter  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's filter step.  It is the
checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``rpsmf_b200/`` does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this file against
(a) the published per-repeat goldens of the reference
    (``ExperimentImpute/output/LondonAir_PM25_30_{PSMF,rPSMF}.json``: input
    hashes and ``error_full`` / ``error_predict`` / ``inside_sig``), replayed
    through fixtures under ``tests/golden/`` that were produced by importing
    the *unmodified* reference (``tests/golden/make_golden.py``), and
(b) per-step trajectories dumped from the reference's own loop bodies.

What is restated (reference file:line, relative to /root/reference):

* masked rPSMF step     ExperimentImpute/rPSMF.py:81-135  (+ compute_Sinv :30-36)
* masked PSMF step      ExperimentImpute/PSMF.py:60-84    (+ compute_Sinv :30-36)
* sweep carry-over      ExperimentImpute/rPSMF.py:75-79,86 ; PSMF.py:59,65
* general (pypsmf) step pypsmf/psmf/psmf.py:90-165 ; rpsmf.py:116-171
* simplified step       ExperimentSynthetic/synthetic_psmf.py:78-100,
                        synthetic_rpsmf.py:82-118
* metrics               ExperimentImpute/common.py:79-94
* theta gradient        pypsmf/psmf/psmf.py:48-66,167-177 ; rpsmf.py:53-73

The reference materialises d x d matrices (``np.diag(M[:, t])``, ``Ri``,
``Skinv``); this restatement uses the algebraically identical O(d r^2)
sufficient-statistic form:

    w_i = 1 / (m_i rho_i + a)                 a = xbar' V xbar
    G   = sum_i m_i w_i c_i c_i'              b = sum_i m_i w_i e_i c_i
    s   = sum_i w_i e_i^2                     G0 = sum_i m_i c_i c_i'
    K   = (Pbar^-1 + G)^-1 = (I + Pbar G)^-1 Pbar
    x   = xbar + K b        e' S^-1 e = s - b' K b        P = beta omega K
"""

from __future__ import annotations

import dataclasses
import numpy as np

# dynamics ids shared with include/psmf_b200.h
DYN_IDENTITY = 0      # pypsmf/psmf/nonlinearities.py:42-56  (RandomWalk)
DYN_COS = 1           # ExperimentSynthetic/synthetic_psmf.py:105-106
DYN_LINEAR = 2        # x_bar = A x + c, F = A (psmf.py:104-115 with a linear f; ExperimentChange/PSMF.m:29-30)
DYN_EXTERNAL = 3      # xbar and F supplied by the caller for every step


@dataclasses.dataclass
class OracleConfig:
    robust: bool = True          # rPSMF (Student-t scales omega, phi) vs PSMF
    simplified: bool = False     # ExperimentSynthetic overrides (P_bar=P, eta=tr(R)/d, no x update)
    c_update_transpose: bool = True   # True: C += e (V xbar)'/N  (rPSMF.py:111, psmf.py:132)
    #                                   False: C += e (V' xbar)'/N (PSMF.py:80)
    bounds_rpsmf: bool = True    # True: sqrt(a m_i + eta) (rPSMF.py:112,121); False: sqrt(N) (PSMF.py:83)
    alpha: float = 1.0           # rpsmf.py:45-51
    beta: float = 1.0
    fixed_lambda: bool = False   # rpsmf.py:36-40
    sig: float = 2.0
    dynamics: int = DYN_IDENTITY
    d_global: int | None = None  # denominator d (differs from local rows only when sharded)
    lin_A: np.ndarray | None = None   # DYN_LINEAR
    lin_c: np.ndarray | None = None


@dataclasses.dataclass
class OracleState:
    C: np.ndarray        # (d, r)
    x: np.ndarray        # (r,)
    P: np.ndarray        # (r, r)
    V: np.ndarray        # (r, r)
    Q: np.ndarray        # (r, r)
    rho: float           # R = rho * I_d (float: uniform diagonal, every experiment uses this) or a (d,) vector = diag(R)
    lam: float
    theta: np.ndarray | None = None

    def copy(self) -> "OracleState":
        return OracleState(
            self.C.copy(), self.x.copy(), self.P.copy(), self.V.copy(),
            self.Q.copy(), float(self.rho) if np.ndim(self.rho) == 0 else np.array(self.rho, dtype=np.float64), float(self.lam),
            None if self.theta is None else self.theta.copy(),
        )


def dynamics(kind: int, theta, x, k, cfg=None):
    """Return (xbar, F) with F = d f / d x   (psmf.py:104-115)."""
    r = x.shape[0]
    if kind == DYN_LINEAR:
        A = np.asarray(cfg.lin_A, dtype=np.float64)
        xb = A @ x
        if cfg.lin_c is not None:
            xb = xb + np.asarray(cfg.lin_c, dtype=np.float64).reshape(-1)
        return xb, A
    if kind == DYN_IDENTITY:
        return x.copy(), np.eye(r)
    if kind == DYN_COS:
        th = np.asarray(theta, dtype=np.float64).reshape(-1)
        arg = 2.0 * np.pi * th * k + x
        return np.cos(arg), np.diag(-np.sin(arg))
    raise ValueError("unknown dynamics id %r" % kind)


def local_stats(C, xbar, a, rho, y, m):
    """Row pass: everything that needs a sweep over the d rows.

    Returns yhat (unmasked prediction), e, and the dict of sufficient
    statistics.  This is the part that is summed across row shards.
    """
    yhat = C @ xbar                                   # rPSMF.py:89
    e = y - m * yhat                                  # rPSMF.py:101
    w = 1.0 / (m * rho + a)                           # rPSMF.py:92,98,32
    mw = m * w
    G = C.T @ (mw[:, None] * C)                       # rPSMF.py:35  CM' Ri CM
    b = C.T @ (mw * e)                                # CM' Ri diff
    s = float(np.sum(w * e * e))                      # diff' Ri diff
    obs = m > 0
    q1 = float(np.sum(e[obs] ** 2))
    q0 = float(np.sum(e[~obs] ** 2))
    nobs = float(np.sum(m))
    S = dict(G=G, b=b, s=s, q1=q1, q0=q0, nobs=nobs)
    if np.ndim(rho) != 0:
        # non-uniform diagonal R: the unweighted Gram and sum m_i rho_i are needed for eta (rPSMF.py:95,108), and phi
        # needs sum e_i^2 / (a m_i + eta) -- with a per-row mask only, so q1 / q0 still suffice (rPSMF.py:112-114)
        S["G0"] = C.T @ (m[:, None] * C)
        S["nrho"] = float(np.sum(m * rho))
    return yhat, e, S


def small_update(st: OracleState, cfg: OracleConfig, xbar, F, vx, vxt, a, S, d):
    """r x r part of the step given the reduced statistics S (rPSMF.py:102-115,133-135)."""
    r = xbar.shape[0]
    G, b, s, q1, q0, nobs = S["G"], S["b"], S["s"], S["q1"], S["q0"], S["nobs"]
    rho, lam = st.rho, st.lam
    if cfg.simplified:
        Pbar = st.P                                   # synthetic_psmf.py:83-84
        K = None
        x_new = xbar.copy()                           # synthetic_psmf.py:93-94
        sSe = s                                       # synthetic_rpsmf.py:93-98 (S^-1 = Rbar^-1)
        eta = float(np.mean(rho))                     # tr(R)/d, synthetic_psmf.py:86-87
    else:
        Pbar = F @ st.P @ F.T + st.Q                  # rPSMF.py:87 / psmf.py:115
        K = np.linalg.solve(np.eye(r) + Pbar @ G, Pbar)
        Kb = K @ b
        x_new = xbar + Kb                             # rPSMF.py:104
        sSe = s - float(b @ Kb)                       # diff' CPinv diff, rPSMF.py:105
        if "G0" in S:                                 # non-uniform diagonal R
            eta = (S["nrho"] + float(np.sum(Pbar * S["G0"]))) / d        # rPSMF.py:108: trace(M R M + CM PP CM') / d
        else:
            G0 = (rho + a) * G                        # unweighted Gram (uniform rho)
            eta = (rho * nobs + float(np.sum(Pbar * G0))) / d     # rPSMF.py:108
    if cfg.robust:
        omega = (lam + sSe) / (lam + d)               # rPSMF.py:105
    else:
        omega = 1.0
    N = a + eta                                       # rPSMF.py:109
    if cfg.robust:
        # rPSMF.py:112-114.  A step with every row missing has eta = 0 and the reference evaluates
        # 0 * inf = NaN there (Ui = 1/diag(U), rPSMF.py:113); mirrored here with numpy semantics.
        with np.errstate(divide="ignore", invalid="ignore"):
            phi = float((lam + np.float64(q1) / (a + eta) + np.float64(q0) / np.float64(eta)) / (lam + d))
    else:
        phi = 1.0
    if cfg.simplified:
        P_new = Pbar                                  # synthetic_rpsmf.py:109
        Q_new = st.Q                                  # synthetic_rpsmf.py:112
    else:
        P_new = (cfg.beta * omega) * K                # rPSMF.py:106 / rpsmf.py:160-167
        Q_new = omega * st.Q                          # rPSMF.py:133
    V_new = (cfg.alpha * phi) * (st.V - np.outer(vx, vxt) / N)   # rPSMF.py:115
    rho_new = omega * rho                             # rPSMF.py:134
    lam_new = lam if (cfg.fixed_lambda or not cfg.robust) else lam + d   # rPSMF.py:135
    g = (vx if cfg.c_update_transpose else vxt) / N   # rank-1 direction, rPSMF.py:111
    scal = dict(a=a, eta=eta, N=N, omega=omega, phi=phi, sSe=sSe, lam=lam, rho=float(np.mean(rho)))
    return x_new, P_new, V_new, Q_new, rho_new, lam_new, g, scal


def step(st: OracleState, cfg: OracleConfig, y, m, k=0, xbar_F=None):
    """One filter step. Mutates nothing; returns (new_state, out dict)."""
    d = st.C.shape[0]
    dg = cfg.d_global or d
    y = np.asarray(y, dtype=np.float64)
    m = np.ones(d) if m is None else np.asarray(m, dtype=np.float64)
    if xbar_F is not None:
        xbar, F = xbar_F
    else:
        xbar, F = dynamics(cfg.dynamics, st.theta, st.x, k, cfg)
    vx = st.V @ xbar
    vxt = st.V.T @ xbar
    a = float(xbar @ vx)                              # rPSMF.py:93
    yhat, e, S = local_stats(st.C, xbar, a, st.rho, y, m)
    x_new, P_new, V_new, Q_new, rho_new, lam_new, g, scal = small_update(
        st, cfg, xbar, F, vx, vxt, a, S, dg)
    C_new = st.C + np.outer(e, g)                     # rPSMF.py:111
    if cfg.bounds_rpsmf:
        U = a * m + scal["eta"]                       # rPSMF.py:112,121
    else:
        U = np.full(d, scal["N"])                     # PSMF.py:83-84
    sq = cfg.sig * np.sqrt(U)
    new = OracleState(C_new, x_new, P_new, V_new, Q_new, rho_new, lam_new, st.theta)
    out = dict(yhat=yhat, lo=yhat - sq, hi=yhat + sq, e=e, xbar=xbar, stats=S, **scal)
    return new, out


def run(st: OracleState, cfg: OracleConfig, Y, M=None, k0=1, record=None):
    """Run T steps over time-major Y (T, d) / M (T, d). Returns (state, X (T,r), Yrec (T,d), scal (T,8)).

    ``record`` (optional list) receives the per-step out dicts.
    """
    T = Y.shape[0]
    r = st.x.shape[0]
    X = np.zeros((T, r))
    Yrec = np.zeros(Y.shape)
    scal = np.zeros((T, 8))
    for t in range(T):
        st, out = step(st, cfg, Y[t], None if M is None else M[t], k=k0 + t)
        X[t] = st.x
        Yrec[t] = out["yhat"]
        scal[t] = [out[n] for n in SCALAR_NAMES]
        if record is not None:
            record.append(out)
    return st, X, Yrec, scal


# per-step scalar record layout, shared with the C ABI (include/psmf_b200.h PSMF_SCAL_*)
SCALAR_NAMES = ("a", "eta", "N", "omega", "phi", "sSe", "lam", "rho")


# ---------------------------------------------------------------------------
# Metrics  (ExperimentImpute/common.py:79-94)
# ---------------------------------------------------------------------------

def rmsem(Y1, Y2, Mm):
    n = np.sum(Mm)
    return float(np.sqrt(np.sum(((Y1 - Y2) * Mm) ** 2) / n))      # common.py:79-84


def inside_bars(Mm, Yorg, lo, hi):
    sel = Mm == 1
    return float(np.sum((Yorg[sel] < hi[sel]) & (lo[sel] < Yorg[sel])) / np.sum(Mm))   # common.py:87-94


# ---------------------------------------------------------------------------
# The two flat model functions of ExperimentImpute, restated on top of step()
# ---------------------------------------------------------------------------

def impute_fit(Y, C, X, M, Mmiss, V, Q0, rho0, P, lam0, sig, Iter, YorigInt, Einit, robust):
    """robust_PSMF (rPSMF.py:39-148) / ProbabilisticSequentialMatrixFactorizer (PSMF.py:39-95).

    Y, M, Mmiss, YorigInt are (d, n) as in the reference; X (r, n) is mutated in
    place like the reference does (rPSMF.py:104).  Returns
    (Epred, Efull, InsideBars, final_state, Yrec, YrecL, YrecH).
    """
    d, n = Y.shape
    cfg = OracleConfig(robust=robust, c_update_transpose=robust, bounds_rpsmf=robust, sig=sig)
    Epred = np.zeros((1, Iter + 1)); Efull = np.zeros((1, Iter + 1))
    Epred[:, 0] = Einit; Efull[:, 0] = Einit
    Yrec = np.zeros((d, n)); lo = np.zeros((d, n)); hi = np.zeros((d, n))
    rho_of = (lambda: float(rho0)) if np.ndim(rho0) == 0 else (lambda: np.array(rho0, dtype=np.float64))
    st = OracleState(C.copy(), X[:, n - 1].copy(), P.copy(), V.copy(), Q0.copy(), rho_of(), float(lam0))
    Mf = M.astype(np.float64)
    for i in range(Iter):
        st.Q = Q0.copy(); st.rho = rho_of(); st.lam = float(lam0)         # rPSMF.py:77-79
        st.x = X[:, n - 1].copy()                                          # rPSMF.py:86 wrap-around
        for t in range(n):
            st, out = step(st, cfg, Y[:, t], Mf[:, t])
            X[:, t] = st.x
            Yrec[:, t] = out["yhat"]; lo[:, t] = out["lo"]; hi[:, t] = out["hi"]
        Epred[:, i + 1] = rmsem(Yrec, YorigInt, Mmiss)                     # rPSMF.py:139
        Efull[:, i + 1] = rmsem(st.C @ X, YorigInt, Mmiss)                 # rPSMF.py:137,140
    ib = inside_bars(Mmiss, YorigInt, lo, hi)
    return Epred, Efull, ib, st, Yrec, lo, hi


# ---------------------------------------------------------------------------
# theta-gradient of the incremental negative log-likelihood
# (psmf.py:57-64 / rpsmf.py:62-71), closed form for f = cos(2 pi theta t + x)
# ---------------------------------------------------------------------------

def dll_df(robust, f, V, Cte, q, eta, lam, d):
    """d ell / d f where Cte = C'(y - C f), q = ||y - C f||^2, s = f'Vf + eta."""
    Vs = 0.5 * (V + V.T)
    Vf = Vs @ f
    s = float(f @ V @ f) + eta
    if not robust:
        return (d / s - q / (s * s)) * Vf - Cte / s
    gq = 1.0 + q / (lam * s)
    return d * Vf / s - (d + lam) / (gq * lam * s) * (Cte + (q / s) * Vf)


def theta_grad_cos(robust, theta, mu_prev, k, y, C, V, eta, lam, d):
    th = np.asarray(theta, dtype=np.float64).reshape(-1)
    arg = 2.0 * np.pi * th * k + mu_prev
    f = np.cos(arg)
    e = y - C @ f
    g_f = dll_df(robust, f, V, C.T @ e, float(e @ e), eta, lam, d)
    return g_f * (-np.sin(arg)) * (2.0 * np.pi * k)


def predict(st: OracleState, cfg: OracleConfig, T, n_pred):
    """PSMFIter.predict (psmf.py:182-188): roll mu through f for n_pred steps from the filtered state at T and emit
    C mu.  Returns (mu_pred (n_pred, r), y_pred (n_pred, d))."""
    x = st.x.copy()
    mus = []
    for k in range(T + 1, T + n_pred + 1):
        x, _ = dynamics(cfg.dynamics, st.theta, x, k, cfg)
        mus.append(x.copy())
    mus = np.stack(mus)
    return mus, mus @ st.C.T


def psmf_statespace_m(r, Y, Q, A, R, H, V, P0, C, X, m, n, X0):
    """Line-by-line numpy restatement of ExperimentChange/PSMF.m:6-45 (linear dynamics A, observation selector H), with
    the random initial state X0 of PSMF.m:12 passed in.  The Matlab reference cannot run here: this restatement is the
    only checker of the (A, H) form -- PARITY UNPINNED against the Matlab code itself.  Returns (X, C, V, P)."""
    C = np.array(C, dtype=np.float64); V = np.array(V, dtype=np.float64); X = np.array(X, dtype=np.float64)
    P = np.array(P0, dtype=np.float64)
    xprev = np.asarray(X0, dtype=np.float64).reshape(-1, 1)
    for t in range(n):
        Xp = A @ xprev                                                    # PSMF.m:13,29
        PP = A @ P @ A.T + Q                                              # PSMF.m:14,30  (t = 1: A P0 A' + Q)
        HX = H @ Xp
        # PSMF.m:16,32 writes `barR = R + (H*Xp)'*V*(H*Xp)`: in Matlab a scalar added to a matrix lands on EVERY entry, which
        # is not the model's Rbar = R + (x'Vx) I (psmf.py:141-143, rPSMF.py:92-98).  The model's form is restated here -- and
        # implemented by rpsmf_b200.statespace; a literal port would add a rank-one all-ones term to S.
        barR = R + (HX.T @ V @ HX).item() * np.eye(m)
        S = C @ (H @ PP @ H.T) @ C.T + barR                               # PSMF.m:17,33
        e = Y[:, [t]] - C @ HX
        Kg = PP @ H.T @ C.T @ np.linalg.inv(S)
        xnew = Xp + Kg @ e                                                # PSMF.m:18,34
        P = PP - Kg @ C @ H @ PP                                          # PSMF.m:19,35
        eta = np.trace(C @ (H @ PP @ H.T) @ C.T + R) / m                  # PSMF.m:21,37
        Nt = (HX.T @ V @ HX).item() + eta                                   # PSMF.m:22,38
        C = C + (e @ Xp.T @ H.T @ V) / Nt                                 # PSMF.m:24,40
        V = V - (V @ H @ (Xp @ Xp.T) @ H.T @ V) / Nt                      # PSMF.m:25,41
        X[:, [t]] = xnew
        xprev = xnew
    return X, C, V, P
