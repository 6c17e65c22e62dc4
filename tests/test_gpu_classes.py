"""GPU parity tests of the pypsmf class surface (rpsmf_b200.psmf / rpsmf) against fixtures produced by the
unmodified reference classes (tests/golden/make_golden.py).

fp64 tolerance 1e-9 norm-wise for everything that does not involve the reference's theta gradient; the
fixtures' gradients come from a central-difference stand-in for autograd (oracle/ref_loader.py), so
quantities downstream of a theta update are compared at 1e-5.
"""

import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle import psmf_oracle as po
from synth import impute_init, make_problem

pytestmark = pytest.mark.gpu

TOL = 1e-9
GTOL = 1e-5


def cosnl(theta, x, t):
    return np.cos(2 * np.pi * theta * t + x)      # same expression as synthetic_psmf.py:105-106


def _ydict(Y):
    T, d = Y.shape
    return {k + 1: Y[k].reshape(d, 1) for k in range(T)}


def test_psmfiter_full_step_cos():
    from rpsmf_b200 import PSMFIter
    g = load_golden("pypsmf_cases")
    tag = "psmf_full"
    Y = g[tag + "_Y"]; T, d = Y.shape; r = g[tag + "_C0"].shape[1]
    Qs = {k: g[tag + "_Q"] for k in range(T + 1)}
    Rs = {k: float(g[tag + "_rho"]) * np.eye(d) for k in range(T + 1)}
    o = PSMFIter(g[tag + "_theta0"].reshape(r, 1), g[tag + "_C0"], g[tag + "_V0"], g[tag + "_mu0"].reshape(r, 1),
                 g[tag + "_P0"], Qs, Rs, cosnl)
    o.step(_ydict(Y), 1, T)
    yp = np.stack([o._y_pred[k].reshape(d) for k in range(1, T + 1)])
    assert relerr(yp, g[tag + "_ypred"]) < TOL
    assert relerr(o._C[T], g[tag + "_C"]) < TOL
    assert relerr(o._mu[T].reshape(-1), g[tag + "_mu"]) < TOL
    assert relerr(o._P[T], g[tag + "_P"]) < TOL
    assert relerr(o._V[T], g[tag + "_V"]) < TOL
    assert o._mu[T].shape == (r, 1) and o._y_pred[1].shape == (d, 1)
    assert relerr(o._gradsum.reshape(-1), g[tag + "_gradsum"]) < 1e-6      # closed form vs finite differences
    o.close()


@pytest.mark.parametrize("tag,scaling", [("rpsmf_full", False), ("rpsmf_scaled", True)])
def test_rpsmfiter_full_step(tag, scaling):
    from rpsmf_b200 import rPSMFIter
    g = load_golden("pypsmf_cases")
    Y = g[tag + "_Y"]; T, d = Y.shape; r = g[tag + "_C0"].shape[1]
    o = rPSMFIter(g[tag + "_theta0"].reshape(r, 1), g[tag + "_C0"], g[tag + "_V0"], g[tag + "_mu0"].reshape(r, 1),
                  g[tag + "_P0"], g[tag + "_Q"], float(g[tag + "_rho"]) * np.eye(d), float(g[tag + "_lam0"]), cosnl,
                  use_scaling=scaling)
    if scaling:     # compute_scaling_factor (rpsmf.py:75-104) is a host scalar: same mpmath root
        assert abs(o._alpha - float(g[tag + "_alpha"])) < 1e-12 and abs(o._beta - float(g[tag + "_beta"])) < 1e-12
    o.step(_ydict(Y), 1, T)
    yp = np.stack([o._y_pred[k].reshape(d) for k in range(1, T + 1)])
    assert relerr(yp, g[tag + "_ypred"]) < TOL
    assert relerr(o._C[T], g[tag + "_C"]) < TOL
    assert relerr(o._mu[T].reshape(-1), g[tag + "_mu"]) < TOL
    assert relerr(o._P[T], g[tag + "_P"]) < TOL
    assert relerr(o._V[T], g[tag + "_V"]) < TOL
    assert relerr(o._lambda[T], g[tag + "_lam_T"]) < TOL
    assert relerr(o._R[T][0, 0], g[tag + "_rho_T"]) < TOL
    assert relerr(o._Q[T], g[tag + "_Q_T"]) < TOL
    assert relerr(o._gradsum.reshape(-1), g[tag + "_gradsum"]) < 1e-6
    o.close()


def test_psmfiter_random_walk_rank1():
    from rpsmf_b200 import PSMFIter
    from rpsmf_b200.nonlinearities import RandomWalk
    g = load_golden("pypsmf_cases")
    tag = "psmf_rw"
    Y = g[tag + "_Y"]; T, d = Y.shape; r = 1
    o = PSMFIter(np.zeros((1, 1)), g[tag + "_C0"], g[tag + "_V0"], g[tag + "_mu0"].reshape(r, 1), g[tag + "_P0"],
                 {k: g[tag + "_Q"] for k in range(T + 1)}, {k: float(g[tag + "_rho"]) * np.eye(d) for k in range(T + 1)},
                 RandomWalk())
    o.run(_ydict(Y), T, 1, 3)
    assert relerr(o._C[T], g[tag + "_C"]) < TOL
    assert relerr(o._P[T], g[tag + "_P"]) < TOL
    assert np.array_equal(o._mu_pred[T + 3], o._mu[T])          # random walk forecast
    assert relerr(o._y_pred[T + 2], o._C[T] @ o._mu[T]) < 1e-15
    o.close()


@pytest.mark.parametrize("tag,robust", [("syn_psmf", False), ("syn_rpsmf", True)])
def test_synthetic_experiment_sweeps_with_theta_learning(tag, robust):
    """Configs 1-2 of BASELINE.json, shortened (3 sweeps of T = 150): simplified step, V re-initialised each
    sweep (the experiments' step_reset override), Adam on theta."""
    from rpsmf_b200 import PSMFIter, rPSMFIter
    g = load_golden("pypsmf_cases")
    Y = g[tag + "_Y"]; T, d = Y.shape; C0 = g[tag + "_C0"]; r = C0.shape[1]
    theta0 = g[tag + "_theta0"].reshape(r, 1)
    base = rPSMFIter if robust else PSMFIter

    class Synthetic(base):
        def step_reset(self):                      # synthetic_psmf.py:78-81
            super().step_reset()
            self._V = {0: self.V0}

    V0 = g[tag + "_V0"]; mu0 = np.zeros((r, 1)); P0 = np.zeros((r, r))
    if robust:
        o = Synthetic(theta0, C0, V0, mu0, P0, 0 * np.eye(r), np.eye(d), 1.8, cosnl, simplified=True)
    else:
        o = Synthetic(theta0, C0, V0, mu0, P0, {k: 0 * np.eye(r) for k in range(T + 1)},
                      {k: np.eye(d) for k in range(T + 1)}, cosnl, simplified=True)
    y = _ydict(Y)
    o.adam_init(gam=1e-3)
    n_iter = g[tag + "_Cs"].shape[0]
    for i in range(1, n_iter + 1):
        o.step(y, i, T)
        o.predict(i, T, 10)
        assert relerr(o._gradsum.reshape(-1), g[tag + "_grads"][i - 1]) < GTOL
        o.adam_update(i)
        assert relerr(o._theta[i].reshape(-1), g[tag + "_thetas"][i]) < GTOL
        assert relerr(o._C[T], g[tag + "_Cs"][i - 1]) < (TOL if i == 1 else GTOL)
        assert relerr(o._mu[T].reshape(-1), g[tag + "_mus"][i - 1]) < (TOL if i == 1 else GTOL)
    ypp = np.stack([o._y_pred[k].reshape(d) for k in range(T + 1, T + 11)])
    assert relerr(ypp, g[tag + "_ypred_future"]) < GTOL
    o.close()


@pytest.mark.parametrize("tag,robust,ue", [("rec_psmf_1", False, 1), ("rec_psmf_7", False, 7), ("rec_rpsmf_4", True, 4)])
def test_recursive_classes(tag, robust, ue):
    from rpsmf_b200 import PSMFRecursive, rPSMFRecursive
    g = load_golden("pypsmf_recursive")
    Y = g[tag + "_Y"]; T, d = Y.shape; C0 = g[tag + "_C0"]; r = C0.shape[1]
    theta0 = g[tag + "_theta0"].reshape(r, 1)
    args = (theta0, C0, g[tag + "_V0"], g[tag + "_mu0"].reshape(r, 1), g[tag + "_P0"])
    if robust:
        o = rPSMFRecursive(*args, g[tag + "_Q"], np.eye(d), 1.8, cosnl)
    else:
        o = PSMFRecursive(*args, {k: g[tag + "_Q"] for k in range(T + 1)}, {k: np.eye(d) for k in range(T + 1)}, cosnl)
    o.run(_ydict(Y), T, 5, update_every=ue)
    thetas = np.stack([o._theta[k].reshape(-1) for k in range(T + 1)])
    assert relerr(thetas, g[tag + "_thetas"]) < GTOL
    assert relerr(o._C[T], g[tag + "_C"]) < GTOL
    assert relerr(o._mu[T].reshape(-1), g[tag + "_mu"]) < GTOL
    assert relerr(o._P[T], g[tag + "_P"]) < GTOL
    ypp = np.stack([o._y_pred[k].reshape(d) for k in range(T + 1, T + 6)])
    assert relerr(ypp, g[tag + "_ypred_future"]) < GTOL
    o.close()


def test_rpsmfitermissing_masked_sweeps():
    """Masked class surface = masked semantics of ExperimentImpute/rPSMF.py (checked against the oracle)."""
    from rpsmf_b200 import rPSMFIterMissing
    from rpsmf_b200.nonlinearities import RandomWalk
    d, r, T = 60, 5, 80
    Y, M, C0, x0 = make_problem(d, r, T, seed=8)
    init = impute_init(r)
    o = rPSMFIterMissing(np.zeros((1, 1)), C0, init["V"], x0.reshape(r, 1), init["P"], init["Q"], init["rho"] * np.eye(d),
                         init["lam"], RandomWalk())
    y = _ydict(Y)
    m = {k + 1: M[k].reshape(d, 1) for k in range(T)}
    o.run(y, T, 2, 0, m=m)
    st = po.OracleState(C0.copy(), x0.copy(), init["P"], init["V"], init["Q"], init["rho"], init["lam"])
    cfg = po.OracleConfig(robust=True)
    for sweep in range(2):
        st.Q = init["Q"]; st.rho = init["rho"]; st.lam = init["lam"]       # rpsmf.py:106-114
        st, X, Yrec, scal = po.run(st, cfg, Y, M.astype(float))
    assert relerr(o._C[T], st.C) < TOL
    assert relerr(o._mu[T].reshape(-1), st.x) < TOL
    assert relerr(o._P[T], st.P) < TOL
    assert relerr(o._V[T], st.V) < TOL
    yp = np.stack([o._y_pred[k].reshape(d) for k in range(1, T + 1)])
    assert relerr(yp, Yrec * M) < TOL                                       # masked prediction, rpsmf.py:229-230
    o.close()


def test_arbitrary_callable_uses_external_path():
    from rpsmf_b200 import PSMFIter
    d, r, T = 40, 4, 25
    Y, M, C0, x0 = make_problem(d, r, T, seed=13, missing=0.0)
    rng = np.random.RandomState(1)
    A = np.eye(r) + 0.1 * rng.randn(r, r)

    def f(theta, x, t):
        return np.tanh(A @ x) + theta

    theta = 0.05 * np.ones((r, 1))
    Q = 0.05 * np.eye(r)
    o = PSMFIter(theta, C0, 0.5 * np.eye(r), x0.reshape(r, 1), np.eye(r), {k: Q for k in range(T + 1)},
                 {k: 2.0 * np.eye(d) for k in range(T + 1)}, f)
    o.step(_ydict(Y), 1, T)
    st = po.OracleState(C0.copy(), x0.copy(), np.eye(r), 0.5 * np.eye(r), Q, 2.0, 0.0)
    cfg = po.OracleConfig(robust=False)
    for k in range(1, T + 1):
        xb = np.tanh(A @ st.x) + theta.reshape(-1)
        F = (1 - np.tanh(A @ st.x) ** 2)[:, None] * A
        st, out = po.step(st, cfg, Y[k - 1], None, xbar_F=(xb, F))
    assert relerr(o._C[T], st.C) < TOL
    assert relerr(o._mu[T].reshape(-1), st.x) < TOL
    assert relerr(o._P[T], st.P) < TOL
    o.close()


@pytest.mark.parametrize("robust", [False, True])
def test_external_dynamics_learn_theta(robust):
    """An arbitrary callable gets theta learning as well (the reference uses autograd for any f, psmf.py:167-177):
    the kernel returns d ell_k / d f, the host closes the chain rule with J_theta.  Checked against the device-side
    closed-form gradient of the built-in cos dynamics: same callable, forced through the external path."""
    from rpsmf_b200 import PSMFIter, rPSMFIter, _capi
    d, r, T = 30, 4, 40
    Y, M, C0, x0 = make_problem(d, r, T, seed=17, missing=0.0)
    theta0 = 0.001 * np.arange(1, r + 1).reshape(r, 1)

    def f_ext(theta, x, t):
        return np.cos(2 * np.pi * theta * t + x)
    f_ext._psmf_dynamics = _capi.DYN_EXTERNAL

    def make(fn):
        Q, R = 0.05 * np.eye(r), 2.0 * np.eye(d)
        if robust:
            return rPSMFIter(theta0.copy(), C0, 0.5 * np.eye(r), x0.reshape(r, 1), np.eye(r), Q, R, 1.8, fn)
        return PSMFIter(theta0.copy(), C0, 0.5 * np.eye(r), x0.reshape(r, 1), np.eye(r), {k: Q for k in range(T + 1)},
                        {k: R for k in range(T + 1)}, fn)
    a, b = make(cosnl), make(f_ext)
    assert a._dyn == _capi.DYN_COS and b._dyn == _capi.DYN_EXTERNAL
    y = _ydict(Y)
    a.run(y, T, 2, 0)
    b.run(y, T, 2, 0)
    assert np.max(np.abs(a._gradsum)) > 0
    assert relerr(b._gradsum, a._gradsum) < 1e-8
    assert relerr(b._theta[2], a._theta[2]) < 1e-8
    assert relerr(b._C[T], a._C[T]) < 1e-8
    a.close(); b.close()


def test_hook_override_is_rejected():
    from rpsmf_b200 import PSMFIter

    class Bad(PSMFIter):
        def _compute_eta_k(self, k, P_bar):
            return 1.0

    with pytest.raises(NotImplementedError):
        Bad(np.zeros((2, 1)), np.zeros((4, 2)), np.eye(2), np.zeros((2, 1)), np.eye(2), {0: np.eye(2)}, {0: np.eye(4)}, cosnl)


def test_non_diagonal_R_is_rejected():
    from rpsmf_b200 import PSMFIter
    R = np.diag([1.0, 2.0, 1.0, 1.0])
    R[0, 1] = R[1, 0] = 0.1
    o = PSMFIter(np.zeros((2, 1)), np.zeros((4, 2)), np.eye(2), np.zeros((2, 1)), np.eye(2), {0: np.eye(2), 1: np.eye(2)},
                 {0: R, 1: R}, cosnl)
    with pytest.raises(NotImplementedError):
        o.step({1: np.zeros((4, 1))}, 1, 1)


@pytest.mark.parametrize("tag,robust", [("diag_psmf", False), ("diag_rpsmf", True)])
def test_diagonal_non_uniform_R_against_reference_fixture(tag, robust):
    """R = diag(rho_i) with different rho_i (psmf.py:144-152 keeps the Woodbury branch for any diagonal R): the class
    surface against the unmodified reference classes (tests/golden/make_golden.py make_diagR), incl. the theta gradient."""
    from rpsmf_b200 import PSMFIter, rPSMFIter
    g = load_golden("diag_R_cases")
    Y = g[tag + "_Y"]
    T, d = Y.shape
    C0 = g[tag + "_C0"]
    r = C0.shape[1]
    R = np.diag(g[tag + "_rho_vec"])
    th = g[tag + "_theta0"].reshape(r, 1)
    if robust:
        o = rPSMFIter(th, C0, g[tag + "_V0"], g[tag + "_mu0"].reshape(r, 1), g[tag + "_P0"], g[tag + "_Q"], R, float(g[tag + "_lam0"]), cosnl)
    else:
        o = PSMFIter(th, C0, g[tag + "_V0"], g[tag + "_mu0"].reshape(r, 1), g[tag + "_P0"], {k: g[tag + "_Q"] for k in range(T + 1)},
                     {k: R for k in range(T + 1)}, cosnl)
    o.step(_ydict(Y), 1, T)
    yp = np.stack([o._y_pred[k].reshape(d) for k in range(1, T + 1)])
    assert relerr(yp, g[tag + "_ypred"]) < TOL
    assert relerr(o._C[T], g[tag + "_C"]) < TOL and relerr(o._mu[T].reshape(-1), g[tag + "_mu"]) < TOL
    assert relerr(o._P[T], g[tag + "_P"]) < TOL and relerr(o._V[T], g[tag + "_V"]) < TOL
    assert relerr(o._gradsum.reshape(-1), g[tag + "_gradsum"]) < GTOL       # the fixture's gradient is a finite difference
    if robust:
        assert relerr(np.diagonal(o._R[T]), g[tag + "_rho_vec_T"]) < TOL
        assert abs(o._lambda[T] - float(g[tag + "_lam_T"])) < TOL * o._lambda[T]
    o.close()
