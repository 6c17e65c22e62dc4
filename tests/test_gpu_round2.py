"""GPU parity tests of the round-2 kernels and rows (through the C ABI): the resident batch kernel (BASELINE.json
configs[4]), the NaN-encoded mask, the fused evaluation metrics, device ingest and the missing-segment generator, the
forecast and the linear dynamics, and the caller-driven statistics exchange.

Tolerances as everywhere: fp64 1e-9 norm-wise per quantity, fp32 storage 1e-4, mask / index handling exact.
"""

import numpy as np
import pytest

from conftest import impute_case, load_golden, relerr
from oracle import psmf_oracle as po
from synth import impute_init, make_problem

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _torch():
    import torch
    return torch


def _oracle_series(Y, M, C0, x0, init, robust=True):
    st = po.OracleState(C0.copy(), x0.copy(), init["P"].copy(), init["V"].copy(), init["Q"].copy(), init["rho"], init["lam"])
    return po.run(st, po.OracleConfig(robust=robust, c_update_transpose=robust), Y, None if M is None else M.astype(float))


@pytest.mark.parametrize("S,d,r,T,dtype,tol", [
    (64, 512, 8, 40, "f64", TOL),        # the config-5 shape (a 64-series slice of the 4096)
    (3, 512, 8, 40, "f64", TOL),         # few series: the 8-warp variant
    (5, 505, 10, 30, "f64", TOL),        # S&P500-shaped dictionary (the 100-repeat loop of the imputation experiment)
    (9, 96, 16, 30, "f64", TOL),
    (4, 27, 10, 50, "f64", TOL),         # PM25-shaped: a single, partly filled tile
    (16, 512, 8, 40, "f32", 1e-4),
])
def test_batch_kernel_series_against_oracle(S, d, r, T, dtype, tol):
    torch = _torch()
    from rpsmf_b200 import FilterEngine
    tdt = torch.float64 if dtype == "f64" else torch.float32
    Y, M, C0, x0 = make_problem(d, r, T, seed=100 + S, S=S)
    init = impute_init(r)
    eng = FilterEngine(d, r, n_series=S, dtype=tdt, robust=True)
    eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
    Yd, Md = torch.as_tensor(Y, dtype=tdt).cuda(), torch.as_tensor(M).cuda()
    # two launches: C goes back to HBM and returns to shared memory in between
    h = T // 2
    o1 = eng.run(Yd[:, :h], Md[:, :h], k0=1, want_X=True, want_Yrec=True, want_scal=True)
    o2 = eng.run(Yd[:, h:], Md[:, h:], k0=1 + h, want_X=True, want_Yrec=True, want_scal=True)
    assert eng.status() == -1
    info = eng.launch_info()
    assert info["kernel"] == "batch", info
    st = eng.get_state()
    X = torch.cat([o1["X"], o2["X"]], 1).cpu().numpy()
    Yrec = torch.cat([o1["Yrec"], o2["Yrec"]], 1).double().cpu().numpy()
    scal = torch.cat([o1["scal"], o2["scal"]], 1).cpu().numpy()
    rnd = (lambda a: a.astype(np.float32).astype(np.float64)) if dtype == "f32" else (lambda a: a)
    for s in range(S):
        ost, oX, oYrec, oscal = _oracle_series(rnd(Y[s]), M[s], rnd(C0[s]), x0[s], init)
        assert relerr(X[s], oX) < tol and relerr(Yrec[s], oYrec) < tol
        assert relerr(st["C"][s].double().cpu().numpy(), ost.C) < tol
        assert relerr(st["P"][s].cpu().numpy(), ost.P) < tol and relerr(st["V"][s].cpu().numpy(), ost.V) < tol
        for k in range(8):
            assert relerr(scal[s][:, k], oscal[:, k]) < tol, po.SCALAR_NAMES[k]
    eng.close()


def test_batch_and_direct_kernels_agree_and_single_series_uses_batch():
    torch = _torch()
    from rpsmf_b200 import FilterEngine
    d, r, T = 505, 10, 60
    Y, M, C0, x0 = make_problem(d, r, T, seed=71)
    init = impute_init(r)
    res = {}
    for kernel in (0, 1, 3):
        eng = FilterEngine(d, r, robust=True, kernel=kernel)
        eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
        out = eng.run(torch.as_tensor(Y).cuda(), torch.as_tensor(M).cuda(), want_X=True)
        assert eng.status() == -1
        res[kernel] = (out["X"].cpu().numpy(), eng.get_state()["C"].cpu().numpy(), eng.launch_info()["kernel"])
        eng.close()
    assert res[0][2] == "batch" and res[1][2] == "direct" and res[3][2] == "batch"
    assert relerr(res[3][0], res[1][0]) < 1e-12 and relerr(res[3][1], res[1][1]) < 1e-12
    ost, oX, _, _ = _oracle_series(Y, M, C0, x0, init)
    assert relerr(res[3][0], oX) < TOL and relerr(res[3][1], ost.C) < TOL


@pytest.mark.parametrize("kernel,d,S", [(1, 2000, 1), (2, 7680, 1), (3, 300, 4)])
def test_nan_encoded_mask_is_exactly_the_byte_mask(kernel, d, S):
    """PSMF_NAN_MASK: missing entries are NaN in Y and there is no mask stream (rPSMF.py:160-164 before :198-202).  The
    arithmetic is the same, so the results must be BIT-identical to the zero-filled Y + mask byte form."""
    torch = _torch()
    from rpsmf_b200 import FilterEngine
    r, T = 8, 16
    Y, M, C0, x0 = make_problem(d, r, T, seed=5 + kernel, S=None if S == 1 else S)
    init = impute_init(r)
    Yn = Y.copy()
    Yn[M == 0] = np.nan
    outs = []
    for nan_mask in (False, True):
        eng = FilterEngine(d, r, n_series=S, robust=True, kernel=kernel, nan_mask=nan_mask)
        eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
        if nan_mask:
            out = eng.run(torch.as_tensor(Yn).cuda(), None, want_X=True, want_Yrec=True)
        else:
            out = eng.run(torch.as_tensor(Y).cuda(), torch.as_tensor(M).cuda(), want_X=True, want_Yrec=True)
        assert eng.status() == -1
        assert eng.launch_info()["kernel"] == {1: "direct", 2: "tma", 3: "batch"}[kernel]
        outs.append((out["X"].cpu().numpy(), out["Yrec"].cpu().numpy(), eng.get_state()["C"].cpu().numpy()))
        eng.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    # a mask pointer together with the flag is a caller error
    eng = FilterEngine(d, r, n_series=S, robust=True, kernel=kernel, nan_mask=True)
    from rpsmf_b200 import _capi
    with pytest.raises(_capi.PsmfError):
        eng.run(torch.as_tensor(Yn).cuda(), torch.as_tensor(M).cuda())
    eng.close()


@pytest.mark.parametrize("kernel", [1, 3])
@pytest.mark.parametrize("robust", [True, False])
def test_fused_evaluation_matches_oracle(kernel, robust):
    """Epred sums and the 2-sigma coverage count accumulated inside the filter pass, and Efull from psmf_eval_full,
    against the oracle's (d, n) restatement of common.py:79-94 on an imputation fixture."""
    torch = _torch()
    from rpsmf_b200 import FilterEngine, ingest, transpose_mask, _capi
    g = load_golden("impute_pm10_head_20")
    c = impute_case(g)
    r = c["r"]
    d, n = c["Y"].shape
    V, Q, P = 2 * np.eye(r), 0.1 * np.eye(r), np.eye(r)
    X = c["X0"].copy()
    ep, ef, ib, ost, Yrec, lo, hi = po.impute_fit(c["Y"], c["C0"], X, c["M"], c["Mmiss"], V, Q, 10.0, P, 1.8, 2.0, 1, c["YorigInt"], 0.0, robust)
    inside_cnt = float(np.sum((c["Mmiss"] == 1) & (c["YorigInt"] < hi) & (lo < c["YorigInt"])))
    Yt, _ = ingest(c["Y"], want_mask=False)
    Yo, _ = ingest(c["YorigInt"], want_mask=False)
    Mt, Et = transpose_mask(c["M"]), transpose_mask(c["Mmiss"])
    eng = FilterEngine(d, r, robust=robust, c_update_transpose=robust, kernel=kernel)
    eng.set_state(C_=c["C0"], V=V, P=P, x=c["X0"][:, n - 1].copy(), Q=Q, rho=[10.0], lam=[1.8 if robust else 0.0])
    # in three launches: the coverage of the last step of a launch is scored in its flush
    bounds = [0, 1, n // 3, n]
    ev = np.zeros(_capi.NEVAL)
    Xs = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        out = eng.run(Yt[a:b], Mt[a:b], k0=1 + a, want_X=True, Yorig=Yo[a:b], E=Et[a:b], sig=2.0)
        assert eng.status() == -1
        ev += out["eval"].cpu().numpy().reshape(-1)
        Xs.append(out["X"])
    assert eng.launch_info()["kernel"] == {1: "direct", 3: "batch"}[kernel]
    assert ev[_capi.EVAL_COUNT] == float(np.sum(c["Mmiss"]))                       # exact
    assert ev[_capi.EVAL_INSIDE] == inside_cnt                                     # exact: an integer count
    assert abs(np.sqrt(ev[_capi.EVAL_SSE] / ev[_capi.EVAL_COUNT]) - ep[0, 1]) < TOL * ep[0, 1]
    full = eng.eval_full(torch.cat(Xs), Yo, Et).cpu().numpy().reshape(-1)
    assert full[1] == float(np.sum(c["Mmiss"]))
    assert abs(np.sqrt(full[0] / full[1]) - ef[0, 1]) < TOL * ef[0, 1]
    eng.close()


def test_ingest_transpose_and_missing_segments_are_exact():
    """f3: (d, n) + NaN -> time-major on the device, and prepare_missing (common.py:50-76) with the host drawing the same
    random stream as the reference and the device applying the segments: masks, NaN positions and counts are exact."""
    torch = _torch()
    from rpsmf_b200 import ingest, transpose_mask, prepare_missing
    from rpsmf_b200.experiment import prepare_missing_device
    rng = np.random.RandomState(3)
    d, n = 75, 333
    Yorig = rng.randn(d, n)
    Yorig[rng.rand(d, n) < 0.05] = np.nan
    for dtype in (torch.float64, torch.float32):
        Yz, M = ingest(Yorig, dtype=dtype, keep_nan=False)
        Yk, _ = ingest(Yorig, dtype=dtype, keep_nan=True, want_mask=False)
        ref = torch.as_tensor(np.nan_to_num(Yorig, nan=0.0).T.copy()).to(dtype)
        assert torch.equal(Yz.cpu(), ref)
        assert np.array_equal(M.cpu().numpy(), (~np.isnan(Yorig)).T.astype(np.uint8))
        assert np.array_equal(np.isnan(Yk.cpu().numpy()), np.isnan(Yorig).T)
        assert torch.equal(torch.nan_to_num(Yk, nan=0.0).cpu(), ref)
    Mi = (rng.rand(d, n) < 0.3).astype(np.int64)
    assert np.array_equal(transpose_mask(Mi).cpu().numpy(), Mi.T.astype(np.uint8))
    # prepare_missing: host reference stream vs device application
    for pct in (0.2, 0.4):
        np.random.seed(123)
        Ymiss = Yorig.copy()
        ratio, Mmiss = prepare_missing(Ymiss, pct)
        np.random.seed(123)
        Ytm, Etm, ratio_d = prepare_missing_device(Yorig, pct)
        assert ratio_d == ratio
        assert np.array_equal(Etm.cpu().numpy(), Mmiss.T.astype(np.uint8))
        assert np.array_equal(np.isnan(Ytm.cpu().numpy()), np.isnan(Ymiss).T)
        # the random stream was consumed identically: the next draws agree
        a = np.random.rand(3)
        np.random.seed(123)
        prepare_missing(Yorig.copy(), pct)
        assert np.array_equal(a, np.random.rand(3))


@pytest.mark.parametrize("tag,robust", [("lin_psmf", False), ("lin_rpsmf", True)])
@pytest.mark.parametrize("kernel", [0, 1])
def test_linear_dynamics_and_forecast_against_reference_fixture(tag, robust, kernel):
    """f4: PSMF_DYN_LINEAR (x_bar = A x + c) and the device forecast against the unmodified reference classes."""
    torch = _torch()
    from rpsmf_b200 import FilterEngine, _capi
    g = load_golden("pypsmf_linear")
    Y = g[tag + "_Y"]
    T, d = Y.shape
    C0 = g[tag + "_C0"]
    r = C0.shape[1]
    eng = FilterEngine(d, r, robust=robust, dynamics=_capi.DYN_LINEAR, ll_student=robust, kernel=kernel)
    eng.set_linear_dynamics(g[tag + "_A"], g[tag + "_c"])
    eng.set_state(C_=C0, V=g[tag + "_V0"], P=g[tag + "_P0"], x=g[tag + "_mu0"], Q=g[tag + "_Q"], rho=[float(g[tag + "_rho"])],
                  lam=[float(g[tag + "_lam0"]) if robust else 0.0])
    out = eng.run(torch.as_tensor(Y).cuda(), None, k0=1, want_X=True, want_Yrec=True)
    assert eng.status() == -1
    st = eng.get_state()
    assert relerr(out["Yrec"].cpu().numpy(), g[tag + "_ypred"]) < TOL
    assert relerr(st["C"].cpu().numpy(), g[tag + "_C"]) < TOL and relerr(st["x"].cpu().numpy(), g[tag + "_mu"]) < TOL
    assert relerr(st["P"].cpu().numpy(), g[tag + "_P"]) < TOL and relerr(st["V"].cpu().numpy(), g[tag + "_V"]) < TOL
    n_pred = g[tag + "_ypred_future"].shape[0]
    Xp, Yp = eng.predict(n_pred, T + 1)
    assert relerr(Xp.cpu().numpy(), g[tag + "_mu_future"]) < TOL
    assert relerr(Yp.cpu().numpy(), g[tag + "_ypred_future"]) < TOL
    eng.close()


def test_linear_dynamics_pipelined_kernel_against_oracle():
    """The TMA-staged kernel evaluates A x + c in two places (published x_bar, control copy): both must agree bitwise,
    else the statistics drift.  d = 7680 rows, masked, against the oracle and the direct kernel."""
    torch = _torch()
    from rpsmf_b200 import FilterEngine, _capi
    d, r, T = 7680, 12, 20
    Y, M, C0, x0 = make_problem(d, r, T, seed=91)
    init = impute_init(r)
    rng = np.random.RandomState(4)
    A = 0.97 * np.eye(r) + 0.03 * rng.randn(r, r)
    c = 0.01 * rng.randn(r)
    cfg = po.OracleConfig(robust=True, dynamics=po.DYN_LINEAR, lin_A=A, lin_c=c)
    ost = po.OracleState(C0.copy(), x0.copy(), init["P"], init["V"], init["Q"], init["rho"], init["lam"])
    ost, oX, oYrec, _ = po.run(ost, cfg, Y, M.astype(float))
    for kernel in (1, 2):
        eng = FilterEngine(d, r, robust=True, dynamics=_capi.DYN_LINEAR, kernel=kernel)
        eng.set_linear_dynamics(A, c)
        eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
        out = eng.run(torch.as_tensor(Y).cuda(), torch.as_tensor(M).cuda(), want_X=True, want_Yrec=True)
        assert eng.status() == -1
        assert relerr(out["X"].cpu().numpy(), oX) < TOL and relerr(out["Yrec"].cpu().numpy(), oYrec) < TOL
        assert relerr(eng.get_state()["C"].cpu().numpy(), ost.C) < TOL
        mus, yp = po.predict(ost, cfg, T, 5)
        Xp, Yp = eng.predict(5, T + 1)
        assert relerr(Xp.cpu().numpy(), mus) < TOL and relerr(Yp.cpu().numpy(), yp) < TOL
        eng.close()


def test_class_surface_forecast_on_device():
    """PSMFIter.predict now runs on the device (roll-out + one d x r . r x n_pred product): synthetic-class fixture with
    10 forecast steps from the reference."""
    from rpsmf_b200 import PSMFIter
    g = load_golden("pypsmf_cases")
    tag = "syn_psmf"
    Y = g[tag + "_Y"]
    T, d = Y.shape
    C0 = g[tag + "_C0"]
    r = C0.shape[1]

    def cosnl(theta, x, t):
        return np.cos(2 * np.pi * theta * t + x)
    o = PSMFIter(g[tag + "_theta0"].reshape(r, 1), C0, g[tag + "_V0"], np.zeros((r, 1)), np.zeros((r, r)),
                 {k: np.zeros((r, r)) for k in range(T + 1)}, {k: np.eye(d) for k in range(T + 1)}, cosnl, simplified=True)

    class Syn(type(o)):
        def step_reset(self):                       # synthetic_psmf.py:78-81 re-initialises V each sweep
            super().step_reset()
            self._V = {0: self.V0}
    o.__class__ = Syn
    o.adam_init(gam=1e-3)
    for i in range(1, 4):
        o.step({k + 1: Y[k].reshape(d, 1) for k in range(T)}, i, T)
        o.predict(i, T, 10)
        o.adam_update(i)
    ypp = np.stack([o._y_pred[k].reshape(d) for k in range(T + 1, T + 11)])
    assert relerr(ypp, g[tag + "_ypred_future"]) < 1e-5      # downstream of the reference's finite-difference theta gradient
    o.close()


def test_caller_driven_exchange_equals_fused_step():
    """PSMF_XCHG_EXTERNAL (the NCCL-baseline / multi-node path): pass + reduction, caller's all-reduce, update -- with a
    no-op all-reduce on one GPU it must reproduce the fused kernel and the oracle."""
    torch = _torch()
    from rpsmf_b200 import FilterEngine
    d, r, T = 5000, 16, 14
    Y, M, C0, x0 = make_problem(d, r, T, seed=15)
    init = impute_init(r)
    Yd, Md = torch.as_tensor(Y).cuda(), torch.as_tensor(M).cuda()
    eng = FilterEngine(d, r, robust=True, exchange="external")
    eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
    X = torch.empty((T, r), dtype=torch.float64, device="cuda")
    calls = []
    for t in range(T):
        eng.run_split(Yd[t:t + 1], Md[t:t + 1], 1 + t, lambda buf: calls.append(int(buf.numel())), X_out=X[t:t + 1].unsqueeze(0))
    assert eng.status() == -1 and len(calls) == T and calls[0] == 16 * 17 // 2 + 16 + 4
    st = eng.get_state()
    ost, oX, _, _ = _oracle_series(Y, M, C0, x0, init)
    assert relerr(X.cpu().numpy(), oX) < TOL and relerr(st["C"].cpu().numpy(), ost.C) < TOL
    assert relerr(st["P"].cpu().numpy(), ost.P) < TOL and relerr(st["V"].cpu().numpy(), ost.V) < TOL
    eng.close()


def test_full_size_batch_slice_properties():
    """Config 5 at its own per-GPU size (512 series of d = 512, r = 8 = one of 8 GPUs), T = 100: size-independent
    properties -- bit-identical reruns, series are independent (a series filtered alone gives the same bits), a
    never-observed row keeps its C row exactly -- plus the oracle on a sample of series."""
    torch = _torch()
    from rpsmf_b200 import FilterEngine
    S, d, r, T = 512, 512, 8, 100
    rng = np.random.RandomState(8)
    Ct = rng.randn(S, d, r)
    x = rng.randn(S, r)
    Y = np.empty((S, T, d))
    for t in range(T):
        x = x + 0.1 * rng.randn(S, r)
        Y[:, t] = np.einsum("sdr,sr->sd", Ct, x) + np.sqrt(0.1) * rng.standard_t(3, (S, d))
    M = (rng.rand(S, T, d) >= 0.2).astype(np.uint8)
    M[:, :, 17] = 0
    Y *= M
    C0 = rng.rand(S, d, r)
    x0 = rng.rand(S, r)
    init = impute_init(r)
    Yd, Md = torch.as_tensor(Y).cuda(), torch.as_tensor(M).cuda()

    def run(sl):
        eng = FilterEngine(d, r, n_series=len(sl), robust=True, kernel=3)
        eng.set_state(C_=C0[sl], V=init["V"], P=init["P"], x=x0[sl], Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
        out = eng.run(Yd[sl], Md[sl], want_X=True)
        assert eng.status() == -1
        res = out["X"].cpu().numpy().reshape(len(sl), T, r), eng.get_state()["C"].cpu().numpy().reshape(len(sl), d, r)
        eng.close()
        return res
    allS = list(range(S))
    Xa, Ca = run(allS)
    Xb, Cb = run(allS)
    assert np.array_equal(Xa, Xb) and np.array_equal(Ca, Cb)
    assert np.array_equal(Ca[:, 17], C0[:, 17])
    pick = [3, 255, 511]
    Xs, Cs = run(pick)                      # 3 series take the 8-warp variant: another summation tree, same math
    assert relerr(Xs, Xa[pick]) < 1e-12 and relerr(Cs, Ca[pick]) < 1e-12
    for s in pick:
        ost, oX, _, _ = _oracle_series(Y[s], M[s], C0[s], x0[s], init)
        assert relerr(Xa[s], oX) < TOL and relerr(Ca[s], ost.C) < TOL


@pytest.mark.parametrize("method", ["rPSMF", "PSMF"])
def test_impute_functions_with_non_uniform_diagonal_R(method):
    """The two flat model functions with R = diag(rho_i), rho_i all different, against the UNMODIFIED reference functions
    (fixture diag_R_cases.npz): Epred / Efull per sweep, the coverage, and the filtered X."""
    from rpsmf_b200 import ProbabilisticSequentialMatrixFactorizer, robust_PSMF
    g = load_golden("diag_R_cases")
    Yorig = g["imp_Yorig"]
    Mmiss = g["imp_Mmiss"].astype(np.float64)
    Ymiss = Yorig.copy(); Ymiss[Mmiss == 1] = np.nan
    M = (~np.isnan(Ymiss)).astype(np.int64)
    Y = Ymiss.copy(); Y[np.isnan(Y)] = 0
    YorigInt = Yorig.copy(); YorigInt[np.isnan(YorigInt)] = 0
    d, n = Y.shape
    r = g["imp_C0"].shape[1]
    X = g["imp_X0"].copy()
    pre = "imp_%s_" % method
    R = np.diag(g["imp_rho_vec"])
    V, Q, P = 2 * np.eye(r), 0.1 * np.eye(r), np.eye(r)
    if method == "rPSMF":
        ep, ef, rt, ib = robust_PSMF(Y, g["imp_C0"], X, d, n, r, M, Mmiss, V, Q, R, P, 1.8, 2, 2, YorigInt, float(g[pre + "Einit"]))
    else:
        ep, ef, rt, ib = ProbabilisticSequentialMatrixFactorizer(Y, g["imp_C0"], X, d, n, r, M, Mmiss, 0, V, Q, R, P, 2, 2, YorigInt,
                                                                 float(g[pre + "Einit"]))
    assert relerr(ep, g[pre + "Epred"]) < TOL and relerr(ef, g[pre + "Efull"]) < TOL
    assert abs(ib - float(g[pre + "inside"])) < 1e-12
    assert relerr(X, g[pre + "X_final"]) < TOL


def test_statespace_selector_form_against_restatement_of_PSMF_m():
    """The (A, H) state-space form of ExperimentChange/PSMF.m through the rank-2r embedding C_eff = C H, V_eff = H'VH on the
    CUDA engine, against a line-by-line numpy restatement of the Matlab function (oracle.psmf_statespace_m; the Matlab
    reference cannot run here: parity unpinned beyond the restatement)."""
    from rpsmf_b200 import statespace
    rng = np.random.RandomState(12)
    r, m, n = 4, 20, 150
    s = 2 * r
    a2 = np.array([[0.95, 0.05], [-0.1, 0.9]])
    A = np.kron(np.eye(r), a2)                                   # main.m:68-70: kron(eye(r), .) blocks
    Q = np.kron(np.eye(r), np.array([[0.02, 0.005], [0.005, 0.03]]))
    H = np.kron(np.eye(r), np.array([[1.0, 0.0]]))               # selects the first component of every 2-dim block
    V = np.eye(r)
    P0 = np.eye(s)
    C = rng.randn(m, r)
    Ct = rng.randn(m, r)
    x = rng.randn(s)
    Y = np.zeros((m, n))
    for t in range(n):
        x = A @ x + 0.1 * rng.randn(s)
        Y[:, t] = Ct @ (H @ x) + 0.05 * rng.randn(m)
    X0 = np.linalg.cholesky(Q) @ rng.randn(s)
    R = 0.001 * np.eye(m)
    Xo, Co, Vo, Po = po.psmf_statespace_m(r, Y, Q, A, R, H, V, P0, C, np.zeros((s, n)), m, n, X0)
    X = statespace.PSMF(r, Y, Q, A, R, H, V, P0, C, np.zeros((s, n)), m, n, x0=X0)
    assert relerr(X, Xo) < TOL
    assert relerr(statespace.PSMF.last["C"], Co) < TOL and relerr(statespace.PSMF.last["V"], Vo) < TOL
    assert relerr(statespace.PSMF.last["P"], Po) < TOL
