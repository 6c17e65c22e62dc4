"""The algebra behind the software pipeline of psmf_stream.cuh, checked on the CPU against the oracle.

The pipelined kernel never sums over C_t directly: pass t sums over C_{t-1}, e_{t-1}, y_t, m_t (all available
before the solve of step t-1 has produced g_{t-1}) and the control CTA assembles the statistics of step t as

    A_t  = A0 + u g' + g u' + kappa g g'      A0 = sum m_t c c',  u = sum m_t e c,  kappa = sum m_t e^2
    h_t  = h0 + psi g                         h0 = sum m_t y_t c, psi = sum m_t y_t e
    bu_t = h_t - A_t xbar_t,  q1_t = gamma - 2 xbar_t'h_t + xbar_t'A_t xbar_t,  gamma = sum m_t y_t^2
    G = w1 A_t,  b = w1 bu_t,  s = w1 q1 + w0 q0,  q0 = sum_{m_t = 0} y_t^2

with c = rows of C_{t-1}, e = e_{t-1}, g = g_{t-1} (C_t = C_{t-1} + e_{t-1} g_{t-1}').  This test replays a run with
exactly that data flow in numpy and compares every statistic with oracle.local_stats on the true C_t."""

import numpy as np
import pytest

from oracle import psmf_oracle as po
from synth import impute_init, make_problem


@pytest.mark.parametrize("robust", [True, False])
def test_lagged_sums_reproduce_the_step_statistics(robust):
    d, r, T = 400, 7, 12
    Y, M, C0, x0 = make_problem(d, r, T, seed=31)
    M = M.astype(np.float64)
    init = impute_init(r)
    cfg = po.OracleConfig(robust=robust, c_update_transpose=robust)
    st = po.OracleState(C0.copy(), x0.copy(), init["P"], init["V"], init["Q"], init["rho"], init["lam"])
    C_prev = C0.copy()                   # C_{t-1}: what the pass of step t reads
    e_prev = np.zeros(d)                 # e_{t-1}
    g_prev = np.zeros(r)                 # g_{t-1}
    for t in range(T):
        y, m = Y[t], M[t]
        # --- pass t: sums that do not involve g_{t-1} ---
        mc = m[:, None] * C_prev
        A0 = C_prev.T @ mc
        u = mc.T @ e_prev
        h0 = mc.T @ y
        kappa, psi, gamma = np.sum(m * e_prev ** 2), np.sum(m * y * e_prev), np.sum(m * y ** 2)
        q0 = np.sum((1 - m) * y ** 2)
        nobs = np.sum(m)
        # --- control CTA: assemble with g_{t-1} and xbar_t ---
        xbar = st.x.copy()               # identity dynamics
        A = A0 + np.outer(u, g_prev) + np.outer(g_prev, u) + kappa * np.outer(g_prev, g_prev)
        h = h0 + psi * g_prev
        bu = h - A @ xbar
        q1 = gamma - 2 * xbar @ h + xbar @ A @ xbar
        a = float(xbar @ st.V @ xbar)
        w1, w0 = 1.0 / (st.rho + a), 1.0 / a
        # --- the oracle on the true C_t ---
        C_t = st.C
        assert np.allclose(C_t, C_prev + np.outer(e_prev, g_prev), rtol=0, atol=1e-13)
        _, e_t, S = po.local_stats(C_t, xbar, a, st.rho, y, m)
        scale = max(1.0, float(np.abs(S["G"]).max()))
        assert np.abs(w1 * A - S["G"]).max() < 1e-11 * scale
        assert np.abs(w1 * bu - S["b"]).max() < 1e-11 * max(1.0, float(np.abs(S["b"]).max()))
        assert abs(q1 - S["q1"]) < 1e-10 * max(1.0, S["q1"])
        assert abs(q0 - S["q0"]) < 1e-12 * max(1.0, S["q0"]) and nobs == S["nobs"]
        assert abs((w1 * q1 + w0 * q0) - S["s"]) < 1e-10 * max(1.0, S["s"])
        # --- advance: the oracle step yields x_t and C_{t+1}; recover g_t from the rank-1 update ---
        C_before = st.C.copy()
        st, _ = po.step(st, cfg, y, m)
        dC = st.C - C_before             # = e_t g_t'
        i = int(np.argmax(np.abs(e_t)))
        g_prev = dC[i] / e_t[i]
        assert np.allclose(dC, np.outer(e_t, g_prev), rtol=0, atol=1e-12 * max(1.0, float(np.abs(dC).max())))
        C_prev, e_prev = C_before, e_t
