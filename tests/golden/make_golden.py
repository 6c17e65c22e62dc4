# -*- coding: utf-8 -*-
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (imported from /root/reference through oracle/ref_loader.py).

Run here (CPU container, reference mounted):   python tests/golden/make_golden.py

Nothing in this script is needed at test time: the tests only read the .npz
files it writes.  Each fixture stores the exact inputs handed to the reference
function and what the reference returned / mutated, plus per-step internals
captured by ``RecordingNumpy`` (no reference source is edited or copied).
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import ref_loader  # noqa: E402

REF = ref_loader.REF_ROOT


def _impute_inputs(common, Yorig, pct, r):
    """Exactly the draw order of rPSMF.py:196-205 / PSMF.py main."""
    Ymiss = np.copy(Yorig)
    missRatio, missMask = common.prepare_missing(Ymiss, pct / 100)
    M = np.array(np.invert(np.isnan(Ymiss)), dtype=int)
    Y = np.copy(Ymiss)
    Y[np.isnan(Y)] = 0
    d, T = Yorig.shape
    C = np.random.rand(d, r)
    X = np.random.rand(r, T)
    return Y, M, missMask, C, X, missRatio


def _run_impute(mod, method, Y, C, X, M, Mmiss, YorigInt, Iter, r=10, sig=2, rho=10, v=2, q=0.1, p=1.0, lam0=1.8):
    """Call the reference flat function (joblib cache bypassed with .func) under a recording np."""
    common = sys.modules["ref_impute_common"] if "ref_impute_common" in sys.modules else None
    d, n = Y.shape
    V = v * np.eye(r); Q = q * np.eye(r); R = rho * np.eye(d); P = p * np.eye(r)
    Xw = X.copy()
    Einit = mod.RMSEM(C @ Xw, YorigInt, Mmiss)
    rec = ref_loader.RecordingNumpy()
    real_np = mod.np
    mod.np = rec
    try:
        if method == "rPSMF":
            ep, ef, rt, ib = mod.robust_PSMF.func(Y, C, Xw, d, n, r, M, Mmiss, V, Q, R, P, lam0, sig, Iter, YorigInt, Einit)
        else:
            ep, ef, rt, ib = mod.ProbabilisticSequentialMatrixFactorizer.func(
                Y, C, Xw, d, n, r, M, Mmiss, 0, V, Q, R, P, sig, Iter, YorigInt, Einit)
    finally:
        mod.np = real_np
    nst = Iter * n
    assert len(rec.inv_args) == 2 * nst and len(rec.trace_vals) == nst
    eta = np.array(rec.trace_vals) / d
    PP = np.stack(rec.inv_args[0::2])              # (nst, r, r)  P + Q entering each step
    if method == "rPSMF":
        U = np.stack(rec.sqrt_args)[:, :]          # (nst, d) = a*m + eta
        assert U.shape == (nst, d)
        Mt = np.tile(M.T.astype(float), (Iter, 1))
        # a from any observed row: U = a + eta there
        a = np.array([(U[t][Mt[t] > 0][0] - eta[t]) if (Mt[t] > 0).any() else np.nan for t in range(nst)])
    else:
        Nt = np.array([float(np.asarray(x).squeeze()) for x in rec.sqrt_args[0::2]])
        a = Nt - eta
    keep = np.arange(0, nst, 16)
    return dict(Epred=ep, Efull=ef, inside=float(ib), Einit=float(Einit), X_final=Xw,
                eta=eta, a=a, PP_idx=keep, PP=PP[keep])


def make_impute(name, csv, pct, ncols, repeats, Iter=2, published=None):
    common = ref_loader.impute_module("common")
    sys.modules["common"] = common           # rPSMF.py / PSMF.py do `from common import ...`
    mods = {"rPSMF": ref_loader.impute_module("rPSMF"), "PSMF": ref_loader.impute_module("PSMF")}
    Yorig = np.genfromtxt(os.path.join(REF, "ExperimentImpute", "data", csv), delimiter=",")
    if ncols is not None:
        Yorig = np.ascontiguousarray(Yorig[:, :ncols])
    YorigInt = np.copy(Yorig); YorigInt[np.isnan(YorigInt)] = 0
    out = dict(Yorig=Yorig, pct=pct, Iter=Iter, r=10)
    np.random.seed(123)                           # Makefile:176
    for rep in range(repeats):
        Y, M, Mmiss, C, X, ratio = _impute_inputs(common, Yorig, pct, 10)
        hashes = dict(Y=common.matrix_hash(Y), C=common.matrix_hash(C), X=common.matrix_hash(X))
        out["rep%d_Mmiss" % rep] = Mmiss.astype(np.uint8)
        out["rep%d_C0" % rep] = C
        out["rep%d_X0" % rep] = X
        out["rep%d_hashes" % rep] = np.array([hashes["Y"], hashes["C"], hashes["X"]])
        for method, mod in mods.items():
            res = _run_impute(mod, method, Y, C, X, M, Mmiss, YorigInt, Iter)
            for k, v in res.items():
                out["rep%d_%s_%s" % (rep, method, k)] = v
            if published is not None:
                pj = json.load(open(os.path.join(REF, "ExperimentImpute", "output", published % method)))
                assert pj["hashes"]["Y"][rep] == hashes["Y"], "Y hash mismatch vs published run"
                assert pj["hashes"]["C"][rep] == hashes["C"] and pj["hashes"]["X"][rep] == hashes["X"]
                out["rep%d_%s_published" % (rep, method)] = np.array([
                    pj["results"]["error_predict"][rep], pj["results"]["error_full"][rep],
                    pj["results"]["inside_sig"][rep]])
                print(name, method, rep, "published", out["rep%d_%s_published" % (rep, method)],
                      "replayed", res["Epred"][0, -1], res["Efull"][0, -1], res["inside"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


# ---------------------------------------------------------------------------
# pypsmf class surface
# ---------------------------------------------------------------------------

def _cosnl(theta, x, t):
    return np.cos(2 * np.pi * theta * t + x)      # same expression as synthetic_psmf.py:105-106


def _collect(obj, T, d):
    ypred = np.stack([np.asarray(obj._y_pred[k]).reshape(d) for k in range(1, T + 1)])
    res = dict(C=obj._C[T], mu=obj._mu[T].reshape(-1), P=obj._P[T], V=obj._V[T], ypred=ypred)
    if hasattr(obj, "_lambda"):
        res["lam_T"] = float(obj._lambda[T])
        res["rho_T"] = float(obj._R[T][0, 0])
        res["Q_T"] = obj._Q[T]
    return res


def make_pypsmf():
    psmf = ref_loader.pypsmf()
    syn_p = ref_loader.synthetic_module("synthetic_psmf")
    syn_r = ref_loader.synthetic_module("synthetic_rpsmf")
    data = sys.modules["data"] if "data" in sys.modules else ref_loader.synthetic_module("data")
    out = {}

    def gen(seed, d, r, T, student):
        np.random.seed(seed)
        gen_fn = data.generate_t_data if student else data.generate_normal_data
        dat = gen_fn(_cosnl, d=d, T=T, n_pred=0, r=r, var=0.1)
        C0 = 0.1 * np.random.randn(d, r)
        theta0 = 0.1 * np.random.rand(r, 1)
        y = dat["y_train"]
        Ymat = np.stack([y[k].reshape(d) for k in range(1, T + 1)])
        return y, Ymat, C0, theta0

    # (1) full PSMFIter step, cos dynamics, P0 > 0, Q > 0  (psmf.py:90-165)
    d, r, T = 20, 6, 120
    y, Ymat, C0, theta0 = gen(35853, d, r, T, False)
    V0 = 0.1 * np.eye(r); mu0 = 0.3 * np.ones((r, 1)); P0 = 0.5 * np.eye(r)
    Q = 0.01 * np.eye(r); rho = 1.0
    Qs = {k: Q for k in range(T + 1)}; Rs = {k: rho * np.eye(d) for k in range(T + 1)}
    o = psmf.PSMFIter(theta0, C0, V0, mu0, P0, Qs, Rs, _cosnl)
    o.step(y, 1, T)
    res = _collect(o, T, d)
    out.update({"psmf_full_" + k: v for k, v in dict(Y=Ymat, C0=C0, theta0=theta0.reshape(-1), V0=V0, mu0=mu0.reshape(-1),
                                                     P0=P0, Q=Q, rho=rho, gradsum=o._gradsum.reshape(-1), **res).items()})

    # (2) full rPSMFIter step (rpsmf.py:116-171), with and without the KL scaling factors
    for tag, scaling in (("rpsmf_full", False), ("rpsmf_scaled", True)):
        y, Ymat, C0, theta0 = gen(35833, d, r, T, True)
        o = psmf.rPSMFIter(theta0, C0, V0, mu0, P0, Q, rho * np.eye(d), 1.8, _cosnl, use_scaling=scaling)
        o.step(y, 1, T)
        res = _collect(o, T, d)
        out.update({tag + "_" + k: v for k, v in dict(Y=Ymat, C0=C0, theta0=theta0.reshape(-1), V0=V0, mu0=mu0.reshape(-1),
                                                      P0=P0, Q=Q, rho=rho, lam0=1.8, alpha=o._alpha, beta=o._beta,
                                                      gradsum=o._gradsum.reshape(-1), **res).items()})

    # (3) random-walk PSMFIter, rank 1, d = 3 (the ExperimentBeijing shape, beijing_psmf.py:119-140)
    from psmf.nonlinearities import RandomWalk
    np.random.seed(2151)
    d3, r3, T3 = 3, 1, 200
    Y3 = np.cumsum(0.1 * np.random.randn(T3, d3), axis=0) + 1.0
    y3 = {k + 1: Y3[k].reshape(d3, 1) for k in range(T3)}
    C03 = np.random.randn(d3, r3); th3 = np.zeros((1, 1))
    o = psmf.PSMFIter(th3, C03, 0.5 * np.eye(r3), np.ones((r3, 1)), 1.0 * np.eye(r3),
                      {k: 0.05 * np.eye(r3) for k in range(T3 + 1)}, {k: 0.2 * np.eye(d3) for k in range(T3 + 1)},
                      RandomWalk())
    o.step(y3, 1, T3)
    res = _collect(o, T3, d3)
    out.update({"psmf_rw_" + k: v for k, v in dict(Y=Y3, C0=C03, V0=0.5 * np.eye(r3), mu0=np.ones(r3), P0=np.eye(r3),
                                                   Q=0.05 * np.eye(r3), rho=0.2, **res).items()})

    # (4) the simplified synthetic classes, 3 sweeps with theta-learning (configs 1-2, shortened)
    d, r, T, n_iter = 20, 6, 150, 3
    for tag, cls, student, seed in (("syn_psmf", syn_p.PSMFIterSynthetic, False, 35853),
                                    ("syn_rpsmf", syn_r.rPSMFIterSynthetic, True, 35833)):
        y, Ymat, C0, theta0 = gen(seed, d, r, T, student)
        V0 = 0.1 * np.eye(r); mu0 = np.zeros((r, 1)); P0 = np.zeros((r, r))
        if student:
            o = cls(theta0, C0, V0, mu0, P0, 0 * np.eye(r), np.eye(d), 1.8, _cosnl)
        else:
            o = cls(theta0, C0, V0, mu0, P0, {k: 0 * np.eye(r) for k in range(T + 1)},
                    {k: np.eye(d) for k in range(T + 1)}, _cosnl)
        o.adam_init(gam=1e-3)
        thetas = [theta0.reshape(-1)]; grads = []; Cs = []; mus = []
        for i in range(1, n_iter + 1):
            o.step(y, i, T)
            o.predict(i, T, 10)
            grads.append(o._gradsum.reshape(-1).copy())
            o.adam_update(i)
            thetas.append(o._theta[i].reshape(-1).copy())
            Cs.append(o._C[T].copy()); mus.append(o._mu[T].reshape(-1).copy())
        res = _collect(o, T, d)
        ypp = np.stack([np.asarray(o._y_pred[k]).reshape(d) for k in range(T + 1, T + 11)])
        out.update({tag + "_" + k: v for k, v in dict(Y=Ymat, C0=C0, theta0=theta0.reshape(-1), V0=V0,
                                                      thetas=np.stack(thetas), grads=np.stack(grads), Cs=np.stack(Cs),
                                                      mus=np.stack(mus), ypred_future=ypp, **res).items()})
    np.savez_compressed(os.path.join(HERE, "pypsmf_cases.npz"), **out)
    print("wrote pypsmf_cases", os.path.getsize(os.path.join(HERE, "pypsmf_cases.npz")) // 1024, "KiB")


def make_recursive():
    """PSMFRecursive / rPSMFRecursive (psmf.py:275-331, rpsmf.py:290-334): full step, theta updated every
    `update_every` steps inside the sweep (the reference's gradient comes from the finite-difference shim)."""
    psmf = ref_loader.pypsmf()
    data = ref_loader.synthetic_module("data")
    out = {}
    d, r, T = 10, 3, 90
    for tag, robust, ue in (("rec_psmf_1", False, 1), ("rec_psmf_7", False, 7), ("rec_rpsmf_4", True, 4)):
        np.random.seed(5535)
        dat = (data.generate_t_data if robust else data.generate_normal_data)(_cosnl, d=d, T=T, n_pred=0, r=r, var=0.1)
        y = dat["y_train"]
        Ymat = np.stack([y[k].reshape(d) for k in range(1, T + 1)])
        C0 = 0.1 * np.random.randn(d, r)
        theta0 = 0.1 * np.random.rand(r, 1)
        V0 = 0.1 * np.eye(r); mu0 = 0.2 * np.ones((r, 1)); P0 = 0.3 * np.eye(r); Q = 0.02 * np.eye(r)
        if robust:
            o = psmf.rPSMFRecursive(theta0, C0, V0, mu0, P0, Q, np.eye(d), 1.8, _cosnl)
        else:
            o = psmf.PSMFRecursive(theta0, C0, V0, mu0, P0, {k: Q for k in range(T + 1)},
                                   {k: np.eye(d) for k in range(T + 1)}, _cosnl)
        o.run(y, T, 5, update_every=ue)
        res = _collect(o, T, d)
        thetas = np.stack([o._theta[k].reshape(-1) for k in range(T + 1)])
        ypp = np.stack([np.asarray(o._y_pred[k]).reshape(d) for k in range(T + 1, T + 6)])
        out.update({tag + "_" + k: v for k, v in dict(Y=Ymat, C0=C0, theta0=theta0.reshape(-1), V0=V0, mu0=mu0.reshape(-1),
                                                      P0=P0, Q=Q, thetas=thetas, ypred_future=ypp, **res).items()})
    np.savez_compressed(os.path.join(HERE, "pypsmf_recursive.npz"), **out)
    print("wrote pypsmf_recursive", os.path.getsize(os.path.join(HERE, "pypsmf_recursive.npz")) // 1024, "KiB")


def make_linear():
    """Linear dynamics x_bar = A x + c through the reference classes (psmf.py:104-115 with a linear callable; the
    state-space form of ExperimentChange/PSMF.m:29-30 with H = I), full step + predict, PSMF and rPSMF."""
    psmf = ref_loader.pypsmf()
    out = {}
    d, r, T, n_pred = 24, 5, 90, 8
    rng = np.random.RandomState(77)
    A = 0.95 * np.eye(r) + 0.05 * rng.randn(r, r)
    c = 0.02 * rng.randn(r, 1)

    def lin(theta, x, t):
        return A @ x + c

    Ct = rng.randn(d, r)
    x = rng.randn(r, 1)
    Y = np.zeros((T, d))
    for t in range(T):
        x = A @ x + c + 0.1 * rng.randn(r, 1)
        Y[t] = (Ct @ x).reshape(d) + 0.3 * rng.standard_t(3, d)
    y = {k + 1: Y[k].reshape(d, 1) for k in range(T)}
    C0 = 0.1 * rng.randn(d, r)
    V0 = 0.2 * np.eye(r); mu0 = 0.1 * np.ones((r, 1)); P0 = 0.5 * np.eye(r); Q = 0.02 * np.eye(r); rho = 0.5
    th = np.zeros((1, 1))
    for tag, robust in (("lin_psmf", False), ("lin_rpsmf", True)):
        if robust:
            o = psmf.rPSMFIter(th, C0, V0, mu0, P0, Q, rho * np.eye(d), 1.8, lin)
        else:
            o = psmf.PSMFIter(th, C0, V0, mu0, P0, {k: Q for k in range(T + 1)}, {k: rho * np.eye(d) for k in range(T + 1)}, lin)
        o.step(y, 1, T)
        o.predict(1, T, n_pred)
        res = _collect(o, T, d)
        ypp = np.stack([np.asarray(o._y_pred[k]).reshape(d) for k in range(T + 1, T + n_pred + 1)])
        mupp = np.stack([np.asarray(o._mu_pred[k]).reshape(r) for k in range(T + 1, T + n_pred + 1)])
        out.update({tag + "_" + k: v for k, v in dict(Y=Y, C0=C0, V0=V0, mu0=mu0.reshape(-1), P0=P0, Q=Q, rho=rho, A=A, c=c.reshape(-1),
                                                      lam0=1.8, ypred_future=ypp, mu_future=mupp, **res).items()})
    np.savez_compressed(os.path.join(HERE, "pypsmf_linear.npz"), **out)
    print("wrote pypsmf_linear", os.path.getsize(os.path.join(HERE, "pypsmf_linear.npz")) // 1024, "KiB")


def make_diagR():
    """Diagonal, NON-uniform R (psmf.py:144-152 takes the Woodbury branch for any diagonal R; the flat functions accept
    any (d, d) R): the class surface (cos dynamics, full step) and the two masked flat functions on a PM25 head."""
    psmf = ref_loader.pypsmf()
    data = ref_loader.synthetic_module("data")
    out = {}
    d, r, T = 20, 6, 100
    np.random.seed(4242)
    dat = data.generate_t_data(_cosnl, d=d, T=T, n_pred=0, r=r, var=0.1)
    y = dat["y_train"]
    Ymat = np.stack([y[k].reshape(d) for k in range(1, T + 1)])
    C0 = 0.1 * np.random.randn(d, r)
    theta0 = 0.1 * np.random.rand(r, 1)
    rho_vec = 0.5 + 1.5 * np.random.rand(d)
    V0 = 0.1 * np.eye(r); mu0 = 0.3 * np.ones((r, 1)); P0 = 0.5 * np.eye(r); Q = 0.01 * np.eye(r)
    for tag, robust in (("diag_psmf", False), ("diag_rpsmf", True)):
        if robust:
            o = psmf.rPSMFIter(theta0, C0, V0, mu0, P0, Q, np.diag(rho_vec), 1.8, _cosnl)
        else:
            o = psmf.PSMFIter(theta0, C0, V0, mu0, P0, {k: Q for k in range(T + 1)}, {k: np.diag(rho_vec) for k in range(T + 1)}, _cosnl)
        o.step(y, 1, T)
        res = _collect(o, T, d)
        if robust:
            res["rho_vec_T"] = np.diagonal(o._R[T]).copy()
        out.update({tag + "_" + k: v for k, v in dict(Y=Ymat, C0=C0, theta0=theta0.reshape(-1), V0=V0, mu0=mu0.reshape(-1), P0=P0, Q=Q,
                                                      rho_vec=rho_vec, lam0=1.8, gradsum=o._gradsum.reshape(-1), **res).items()})
    # masked flat functions with a non-uniform diagonal R
    common = ref_loader.impute_module("common")
    Yorig = np.genfromtxt(os.path.join(REF, "ExperimentImpute", "data", "LondonAir_PM25.csv"), delimiter=",")[:, :500]
    np.random.seed(123)
    rr = 10
    Y, M, Mmiss, C, X, _ = _impute_inputs(common, Yorig, 30, rr)
    dd, n = Y.shape
    YorigInt = np.copy(Yorig); YorigInt[np.isnan(YorigInt)] = 0
    rho_i = 5.0 + 10.0 * np.random.rand(dd)
    out.update(imp_Yorig=Yorig, imp_Mmiss=Mmiss.astype(np.uint8), imp_C0=C, imp_X0=X, imp_rho_vec=rho_i)
    for method in ("rPSMF", "PSMF"):
        mod = ref_loader.impute_module(method)
        V = 2 * np.eye(rr); Qm = 0.1 * np.eye(rr); R = np.diag(rho_i); P = np.eye(rr)
        Xw = X.copy()
        Einit = mod.RMSEM(C @ Xw, YorigInt, Mmiss)
        if method == "rPSMF":
            ep, ef, rt, ib = mod.robust_PSMF.func(Y, C, Xw, dd, n, rr, M, Mmiss, V, Qm, R, P, 1.8, 2, 2, YorigInt, Einit)
        else:
            ep, ef, rt, ib = mod.ProbabilisticSequentialMatrixFactorizer.func(Y, C, Xw, dd, n, rr, M, Mmiss, 0, V, Qm, R, P, 2, 2, YorigInt, Einit)
        out.update({"imp_%s_%s" % (method, k): v for k, v in dict(Einit=Einit, Epred=ep, Efull=ef, inside=ib, X_final=Xw).items()})
    np.savez_compressed(os.path.join(HERE, "diag_R_cases.npz"), **out)
    print("wrote diag_R_cases", os.path.getsize(os.path.join(HERE, "diag_R_cases.npz")) // 1024, "KiB")


if __name__ == "__main__":
    if not ref_loader.available():
        raise SystemExit("reference tree not found at %s" % REF)
    which = sys.argv[1:] or ["pm25", "pm10", "sp500", "pypsmf", "recursive", "linear", "diagR"]
    if "pm25" in which:
        make_impute("impute_pm25_30", "LondonAir_PM25.csv", 30, None, 1,
                    published="LondonAir_PM25_30_%s.json")
    if "pm10" in which:
        make_impute("impute_pm10_head_20", "LondonAir_PM10.csv", 20, 600, 1)
    if "sp500" in which:
        make_impute("impute_sp500_head_30", "sp500_closing_prices.csv", 30, 160, 1)
    if "pypsmf" in which:
        make_pypsmf()
    if "recursive" in which:
        make_recursive()
    if "linear" in which:
        make_linear()
    if "diagR" in which:
        make_diagR()
