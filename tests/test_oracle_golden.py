"""Pin the oracle (oracle/psmf_oracle.py) against the reference's golden vectors.

* published per-repeat goldens of the reference run (PM25 30 %, repeat 0;
  ExperimentImpute/output/LondonAir_PM25_30_{PSMF,rPSMF}.json)
* reference-generated fixtures (tests/golden/make_golden.py): X trajectory,
  per-step eta / a / P+Q captured from the unmodified loop body
* pypsmf class fixtures (full step with cos dynamics, robust, scaled, random walk,
  simplified synthetic classes)

fp64 tolerance: 1e-9 norm-wise per quantity (max|delta| / max|ref|).
"""

import numpy as np
import pytest

from conftest import impute_case, load_golden, relerr
from oracle import psmf_oracle as po

TOL = 1e-9


def _run_impute(g, method):
    c = impute_case(g)
    r = c["r"]
    d, n = c["Y"].shape
    X = c["X0"].copy()
    Einit = po.rmsem(c["C0"] @ X, c["YorigInt"], c["Mmiss"])
    rec = []
    robust = method == "rPSMF"
    # replay with a per-step record
    cfg = po.OracleConfig(robust=robust, c_update_transpose=robust, bounds_rpsmf=robust, sig=2.0)
    st = po.OracleState(c["C0"].copy(), X[:, n - 1].copy(), np.eye(r), 2 * np.eye(r), 0.1 * np.eye(r), 10.0, 1.8)
    Mf = c["M"].astype(float)
    eta, a, PP = [], [], []
    for i in range(c["Iter"]):
        st.Q = 0.1 * np.eye(r); st.rho = 10.0; st.lam = 1.8
        st.x = X[:, n - 1].copy()
        for t in range(n):
            PP.append(st.P + st.Q)
            st, out = po.step(st, cfg, c["Y"][:, t], Mf[:, t])
            X[:, t] = st.x
            eta.append(out["eta"]); a.append(out["a"])
    Ep, Ef, ib, st2, *_ = po.impute_fit(c["Y"], c["C0"], c["X0"].copy(), c["M"], c["Mmiss"], 2 * np.eye(r),
                                        0.1 * np.eye(r), 10.0, np.eye(r), 1.8, 2.0, c["Iter"], c["YorigInt"],
                                        Einit, robust)
    return dict(X=X, eta=np.array(eta), a=np.array(a), PP=np.stack(PP), Ep=Ep, Ef=Ef, ib=ib, Einit=Einit)


@pytest.mark.parametrize("method", ["rPSMF", "PSMF"])
def test_published_golden_pm25(method):
    g = load_golden("impute_pm25_30")
    res = _run_impute(g, method)
    pub = g["rep0_%s_published" % method]          # error_predict, error_full, inside_sig
    assert abs(res["Ep"][0, -1] - pub[0]) / pub[0] < 1e-10
    assert abs(res["Ef"][0, -1] - pub[1]) / pub[1] < 1e-10
    assert abs(res["ib"] - pub[2]) < 1e-12
    pre = "rep0_%s_" % method
    assert relerr(res["X"], g[pre + "X_final"]) < TOL
    assert relerr(res["eta"], g[pre + "eta"]) < TOL
    assert relerr(res["a"], g[pre + "a"]) < TOL
    assert relerr(res["PP"][g[pre + "PP_idx"]], g[pre + "PP"]) < TOL
    assert relerr(res["Ep"], g[pre + "Epred"]) < TOL
    assert relerr(res["Ef"], g[pre + "Efull"]) < TOL


@pytest.mark.parametrize("name", ["impute_pm10_head_20", "impute_sp500_head_30"])
@pytest.mark.parametrize("method", ["rPSMF", "PSMF"])
def test_reference_fixture(name, method):
    g = load_golden(name)
    res = _run_impute(g, method)
    pre = "rep0_%s_" % method
    assert relerr(res["X"], g[pre + "X_final"]) < TOL
    assert relerr(res["eta"], g[pre + "eta"]) < TOL
    assert relerr(res["a"], g[pre + "a"]) < TOL
    assert relerr(res["PP"][g[pre + "PP_idx"]], g[pre + "PP"]) < TOL
    assert relerr(res["Ep"], g[pre + "Epred"]) < TOL
    assert relerr(res["Ef"], g[pre + "Efull"]) < TOL
    assert abs(res["ib"] - float(g[pre + "inside"])) < 1e-12


def _pypsmf_run(g, tag, robust, dyn, simplified=False):
    Y = g[tag + "_Y"]
    T, d = Y.shape
    C0 = g[tag + "_C0"]
    r = C0.shape[1]
    theta = g[tag + "_theta0"] if (tag + "_theta0") in g else None
    alpha = float(g[tag + "_alpha"]) if (tag + "_alpha") in g else 1.0
    beta = float(g[tag + "_beta"]) if (tag + "_beta") in g else 1.0
    cfg = po.OracleConfig(robust=robust, simplified=simplified, c_update_transpose=True, bounds_rpsmf=robust,
                          alpha=alpha, beta=beta, dynamics=dyn)
    lam0 = float(g[tag + "_lam0"]) if (tag + "_lam0") in g else 0.0
    st = po.OracleState(C0.copy(), g[tag + "_mu0"].copy(), g[tag + "_P0"].copy(), g[tag + "_V0"].copy(),
                        g[tag + "_Q"].copy(), float(g[tag + "_rho"]), lam0, theta)
    st, X, Yrec, scal = po.run(st, cfg, Y, None, k0=1)
    return st, X, Yrec, scal


@pytest.mark.parametrize("tag,robust,dyn", [
    ("psmf_full", False, po.DYN_COS),
    ("rpsmf_full", True, po.DYN_COS),
    ("rpsmf_scaled", True, po.DYN_COS),
    ("psmf_rw", False, po.DYN_IDENTITY),
])
def test_pypsmf_fixture(tag, robust, dyn):
    g = load_golden("pypsmf_cases")
    st, X, Yrec, scal = _pypsmf_run(g, tag, robust, dyn)
    assert relerr(Yrec, g[tag + "_ypred"]) < TOL
    assert relerr(st.C, g[tag + "_C"]) < TOL
    assert relerr(st.x, g[tag + "_mu"]) < TOL
    assert relerr(st.P, g[tag + "_P"]) < TOL
    assert relerr(st.V, g[tag + "_V"]) < TOL
    if robust:
        assert relerr(st.lam, g[tag + "_lam_T"]) < TOL
        assert relerr(st.rho, g[tag + "_rho_T"]) < TOL
        assert relerr(st.Q, g[tag + "_Q_T"]) < TOL


@pytest.mark.parametrize("tag,robust", [("syn_psmf", False), ("syn_rpsmf", True)])
def test_simplified_synthetic_first_sweep(tag, robust):
    """Configs 1-2 (shortened): first sweep with theta0 (synthetic_psmf.py:78-100, synthetic_rpsmf.py:82-118)."""
    g = load_golden("pypsmf_cases")
    Y = g[tag + "_Y"]
    T, d = Y.shape
    C0 = g[tag + "_C0"]
    r = C0.shape[1]
    cfg = po.OracleConfig(robust=robust, simplified=True, dynamics=po.DYN_COS)
    st = po.OracleState(C0.copy(), np.zeros(r), np.zeros((r, r)), g[tag + "_V0"].copy(), np.zeros((r, r)), 1.0,
                        1.8 if robust else 0.0, g[tag + "_thetas"][0])
    st, X, Yrec, scal = po.run(st, cfg, Y, None, k0=1)
    assert relerr(st.C, g[tag + "_Cs"][0]) < TOL
    assert relerr(st.x, g[tag + "_mus"][0]) < TOL
    # the general (non-simplified) formulas coincide when P0 = 0 and Q = 0
    cfg2 = po.OracleConfig(robust=robust, simplified=False, dynamics=po.DYN_COS)
    st2 = po.OracleState(C0.copy(), np.zeros(r), np.zeros((r, r)), g[tag + "_V0"].copy(), np.zeros((r, r)), 1.0,
                         1.8 if robust else 0.0, g[tag + "_thetas"][0])
    st2, *_ = po.run(st2, cfg2, Y, None, k0=1)
    assert relerr(st2.C, st.C) < 1e-12


@pytest.mark.parametrize("tag,robust", [("syn_psmf", False), ("syn_rpsmf", True)])
def test_theta_gradient_closed_form(tag, robust):
    """sum_k d ell_k / d theta over the first sweep vs the reference's _gradsum (finite-difference shim: 1e-6)."""
    g = load_golden("pypsmf_cases")
    Y = g[tag + "_Y"]
    T, d = Y.shape
    C0 = g[tag + "_C0"]
    r = C0.shape[1]
    theta = g[tag + "_thetas"][0]
    cfg = po.OracleConfig(robust=robust, simplified=True, dynamics=po.DYN_COS)
    st = po.OracleState(C0.copy(), np.zeros(r), np.zeros((r, r)), g[tag + "_V0"].copy(), np.zeros((r, r)), 1.0,
                        1.8 if robust else 0.0, theta)
    gs = np.zeros(r)
    for k in range(1, T + 1):
        eta = st.rho
        gs += po.theta_grad_cos(robust, theta, st.x, k, Y[k - 1], st.C, st.V, eta, st.lam, d)
        st, _ = po.step(st, cfg, Y[k - 1], None, k=k)
    assert relerr(gs, g[tag + "_grads"][0]) < 1e-6


@pytest.mark.parametrize("robust", [True, False])
def test_c_oracle_matches_numpy_oracle(robust):
    """oracle/psmf_oracle_c.c (multi-threaded CPU baseline of bench.py) against the pinned numpy oracle."""
    from oracle import psmf_oracle_c as pc
    if not pc.available():
        pytest.skip("oracle/libpsmf_oracle.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    from synth import impute_init, make_problem
    d, r, T = 5000, 16, 30
    Y, M, C0, x0 = make_problem(d, r, T, seed=99)
    init = impute_init(r)
    res = pc.run(C0, x0, init["P"], init["V"], init["Q"], init["rho"], init["lam"], Y, M, robust=robust, cupdate_vt=robust)
    st = po.OracleState(C0.copy(), x0.copy(), init["P"], init["V"], init["Q"], init["rho"], init["lam"])
    st, X, _, _ = po.run(st, po.OracleConfig(robust=robust, c_update_transpose=robust), Y, M.astype(float))
    assert res["bad"] == -1
    assert relerr(res["X"], X) < TOL and relerr(res["C"], st.C) < TOL
    assert relerr(res["P"], st.P) < TOL and relerr(res["V"], st.V) < TOL
    assert relerr(res["rho"], st.rho) < TOL and relerr(res["lam"], st.lam) < TOL


# ---------------------------------------------------------------------------
# round 2: linear dynamics + forecast, and diagonal non-uniform R, pinned against the unmodified reference
# (fixtures: tests/golden/make_golden.py make_linear / make_diagR)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("tag,robust", [("lin_psmf", False), ("lin_rpsmf", True)])
def test_oracle_linear_dynamics_and_predict(tag, robust):
    g = load_golden("pypsmf_linear")
    Y = g[tag + "_Y"]
    T, d = Y.shape
    cfg = po.OracleConfig(robust=robust, dynamics=po.DYN_LINEAR, lin_A=g[tag + "_A"], lin_c=g[tag + "_c"])
    st = po.OracleState(g[tag + "_C0"].copy(), g[tag + "_mu0"].copy(), g[tag + "_P0"].copy(), g[tag + "_V0"].copy(),
                        g[tag + "_Q"].copy(), float(g[tag + "_rho"]), float(g[tag + "_lam0"]) if robust else 0.0)
    st, X, Yrec, scal = po.run(st, cfg, Y, None, k0=1)
    assert relerr(Yrec, g[tag + "_ypred"]) < 1e-9
    assert relerr(st.C, g[tag + "_C"]) < 1e-9 and relerr(st.x, g[tag + "_mu"]) < 1e-9
    assert relerr(st.P, g[tag + "_P"]) < 1e-9 and relerr(st.V, g[tag + "_V"]) < 1e-9
    mus, yp = po.predict(st, cfg, T, g[tag + "_ypred_future"].shape[0])
    assert relerr(mus, g[tag + "_mu_future"]) < 1e-9 and relerr(yp, g[tag + "_ypred_future"]) < 1e-9


@pytest.mark.parametrize("tag,robust", [("diag_psmf", False), ("diag_rpsmf", True)])
def test_oracle_diagonal_nonuniform_R_classes(tag, robust):
    g = load_golden("diag_R_cases")
    Y = g[tag + "_Y"]
    cfg = po.OracleConfig(robust=robust, dynamics=po.DYN_COS)
    st = po.OracleState(g[tag + "_C0"].copy(), g[tag + "_mu0"].copy(), g[tag + "_P0"].copy(), g[tag + "_V0"].copy(),
                        g[tag + "_Q"].copy(), g[tag + "_rho_vec"].copy(), float(g[tag + "_lam0"]) if robust else 0.0,
                        g[tag + "_theta0"].copy())
    st, X, Yrec, scal = po.run(st, cfg, Y, None, k0=1)
    assert relerr(Yrec, g[tag + "_ypred"]) < 1e-9
    assert relerr(st.C, g[tag + "_C"]) < 1e-9 and relerr(st.x, g[tag + "_mu"]) < 1e-9
    assert relerr(st.P, g[tag + "_P"]) < 1e-9 and relerr(st.V, g[tag + "_V"]) < 1e-9
    if robust:
        assert relerr(st.rho, g[tag + "_rho_vec_T"]) < 1e-9 and abs(st.lam - float(g[tag + "_lam_T"])) < 1e-9 * st.lam


@pytest.mark.parametrize("method", ["rPSMF", "PSMF"])
def test_oracle_diagonal_nonuniform_R_impute(method):
    g = load_golden("diag_R_cases")
    Yorig = g["imp_Yorig"]
    Mmiss = g["imp_Mmiss"].astype(np.float64)
    Ymiss = Yorig.copy(); Ymiss[Mmiss == 1] = np.nan
    M = (~np.isnan(Ymiss)).astype(np.int64)
    Y = Ymiss.copy(); Y[np.isnan(Y)] = 0
    YorigInt = Yorig.copy(); YorigInt[np.isnan(YorigInt)] = 0
    r = g["imp_C0"].shape[1]
    X = g["imp_X0"].copy()
    pre = "imp_%s_" % method
    ep, ef, ib, st, *_ = po.impute_fit(Y, g["imp_C0"], X, M, Mmiss, 2 * np.eye(r), 0.1 * np.eye(r), g["imp_rho_vec"], np.eye(r), 1.8, 2, 2,
                                       YorigInt, float(g[pre + "Einit"]), robust=method == "rPSMF")
    assert relerr(ep, g[pre + "Epred"]) < 1e-9 and relerr(ef, g[pre + "Efull"]) < 1e-9
    assert abs(ib - float(g[pre + "inside"])) < 1e-12
    assert relerr(X, g[pre + "X_final"]) < 1e-9


def test_selector_form_of_PSMF_m_is_the_rank_2r_standard_step():
    """ExperimentChange/PSMF.m (linear dynamics A, observation selector H) restated line by line equals the standard PSMF
    step of rank 2r on C_eff = C H, V_eff = H'VH with linear dynamics: the identity rpsmf_b200.statespace relies on."""
    rng = np.random.RandomState(12)
    r, m, n = 4, 20, 60
    s = 2 * r
    A = np.kron(np.eye(r), np.array([[0.95, 0.05], [-0.1, 0.9]]))
    Q = np.kron(np.eye(r), np.array([[0.02, 0.005], [0.005, 0.03]]))
    H = np.kron(np.eye(r), np.array([[1.0, 0.0]]))
    V, P0, C = np.eye(r), np.eye(s), rng.randn(m, r)
    Y, X0, R = rng.randn(m, n), rng.randn(s), 0.001 * np.eye(m)
    Xo, Co, Vo, Po = po.psmf_statespace_m(r, Y, Q, A, R, H, V, P0, C, np.zeros((s, n)), m, n, X0)
    cfg = po.OracleConfig(robust=False, c_update_transpose=False, dynamics=po.DYN_LINEAR, lin_A=A)
    st = po.OracleState(C @ H, X0.copy(), P0.copy(), H.T @ V @ H, Q.copy(), 0.001, 0.0)
    st, X, _, _ = po.run(st, cfg, Y.T.copy(), None)
    assert relerr(X.T, Xo) < 1e-12 and relerr(st.C @ H.T, Co) < 1e-12 and relerr(H @ st.V @ H.T, Vo) < 1e-12 and relerr(st.P, Po) < 1e-12
