# -*- coding: utf-8 -*-
"""Host-side caller of the imputation path: what ``main()`` of ExperimentImpute/rPSMF.py:150-287 and
PSMF.py:98-222 does around the model function -- random missing mask, initial C / X, the fit, the result record.

The random stream is consumed exactly like the reference (``np.random.seed(seed)``, then per repeat: one
``randint`` per row and sweep of ``prepare_missing``, ``rand(d, r)``, ``rand(r, T)``), so a run with the same
seed reproduces the reference's inputs bit for bit; the ``hashes`` of the record are the reference's
``matrix_hash`` (common.py:108-111) and can be compared with the published JSON files.

    python -m rpsmf_b200.experiment -i data.csv -o out.json -m rPSMF -p 30 -s 123 -r 100
"""

from __future__ import annotations

import argparse
import hashlib
import json
import socket

import numpy as np


def prepare_missing(Ymiss, missRatio, misSeg=20):
    """Random missing mask in segments of ``misSeg`` time steps (common.py:50-76).

    ``Ymiss`` (d, n) is modified in place (new missing entries become NaN); returns ``(ratio, Mmiss)`` with
    ``Mmiss`` = 1 where an observed entry was removed.  One ``np.random.randint(1, n - misSeg)`` per row and
    sweep, drawn in row order, exactly as the reference does; the segment itself is filled with a slice."""
    d, n = Ymiss.shape
    n_default = int(np.sum(np.isnan(Ymiss)))
    Mmiss = np.zeros_like(Ymiss)
    ratio = n_default / (d * n)
    while ratio < missRatio:
        for i in range(d):
            start = np.random.randint(1, n - misSeg)
            seg = slice(start, start + misSeg)
            newly = ~np.isnan(Ymiss[i, seg])
            Mmiss[i, seg][newly] = 1
            Ymiss[i, seg] = np.nan
        ratio = (n_default + np.sum(Mmiss)) / (d * n)
    return ratio, Mmiss


def prepare_missing_device(Yorig, missRatio, misSeg=20, dtype=None, device=None):
    """``prepare_missing`` with the data on the GPU: ``Yorig`` (d, n) with NaN = missing is
    ingested once (transpose to time-major on the device, NaN kept); per sweep the host draws the d segment starts from
    the SAME random stream as the reference (one ``randint(1, n - misSeg)`` per row, in row order -- a vectorised draw
    consumes the legacy generator identically) and the device removes the segments and counts.  Returns
    ``(Y_tm, E_tm, ratio)``: time-major NaN-encoded observations (feed them to an engine created with ``nan_mask=True``),
    the uint8 mask of artificially removed entries (``Mmiss`` transposed) and the achieved ratio -- bit-identical to the
    reference's ``(Ymiss, Mmiss, ratio)``; only d integers per sweep cross PCIe."""
    import torch
    from .engine import count_nan, ingest, missing_segments
    Y_tm, _ = ingest(Yorig, dtype=dtype or torch.float64, keep_nan=True, want_mask=False, device=device)
    n, d = Y_tm.shape
    E_tm = torch.zeros((n, d), dtype=torch.uint8, device=Y_tm.device)
    n_default = count_nan(Y_tm)
    removed = 0
    ratio = n_default / (d * n)
    while ratio < missRatio:
        starts = np.random.randint(1, n - misSeg, size=d)          # == d scalar draws in row order (common.py:70)
        removed += missing_segments(Y_tm, E_tm, starts, misSeg)
        ratio = (n_default + float(removed)) / (d * n)
    return Y_tm, E_tm, ratio


def matrix_hash(A):
    """blake2b-128 of the raw bytes (common.py:108-111): the ``hashes`` of the published result files."""
    h = hashlib.blake2b(digest_size=16)
    h.update(np.ascontiguousarray(A).tobytes())
    return h.hexdigest()


def rmsem(Y1, Y2, M):
    """RMSE over the entries marked in M (common.py:79-84)."""
    return float(np.sqrt(np.sum(((Y1 - Y2) * M) ** 2) / np.sum(M)))


# hyper-parameters of the imputation experiment (rPSMF.py:170-183, PSMF.py:118-129)
DEFAULTS = dict(r=10, sig=2, Iter=2, rho=10, v=2, q=0.1, p=1.0, lambda0=1.8)


def _default_fit(method):
    from .impute import ProbabilisticSequentialMatrixFactorizer, robust_PSMF      # needs the CUDA library
    return robust_PSMF if method == "rPSMF" else ProbabilisticSequentialMatrixFactorizer


def run_impute_experiment(Yorig, method="rPSMF", percentage=30, seed=None, repeats=1, fit=None, log=None, batched=False, **hyper):
    """Repeat the imputation fit ``repeats`` times on fresh random masks / initialisations.

    ``Yorig`` (d, T) float64 with NaN at originally missing entries.  ``fit`` defaults to the CUDA model function of
    ``method`` (``robust_PSMF`` / ``ProbabilisticSequentialMatrixFactorizer``); any callable with the same positional
    signature works (the CPU tests pass the oracle).  Returns the result record of ``prepare_output``
    (common.py:114-146) without the host / script provenance fields: ``method, seed, missing_percentage,
    missing_ratio, parameters, hashes{Y,C,X}, results{error_predict, error_full, runtime, inside_sig}``.

    ``batched=True`` draws the inputs of all repeats first (the fits consume no random numbers, so the stream -- and the
    ``hashes`` -- are unchanged) and filters the repeats side by side on the resident batch kernel
    (``rpsmf_b200.impute.fit_repeats``): the 100-repeat loop of the reference as one launch per sweep."""
    if method not in ("rPSMF", "PSMF"):
        raise ValueError("method must be 'rPSMF' or 'PSMF'")
    hp = dict(DEFAULTS)
    unknown = set(hyper) - set(hp)
    if unknown:
        raise TypeError("unknown hyper-parameters: %s" % sorted(unknown))
    hp.update(hyper)
    if batched and fit is not None:
        raise ValueError("batched=True uses the CUDA batch kernel; it takes no custom fit")
    fit = fit or (None if batched else _default_fit(method))
    seed = seed or np.random.randint(10000)                   # rPSMF.py:154
    np.random.seed(seed)
    Yorig = np.asarray(Yorig, dtype=np.float64)
    YorigInt = np.copy(Yorig)
    YorigInt[np.isnan(YorigInt)] = 0
    d, T = Yorig.shape
    r, Iter = hp["r"], hp["Iter"]
    V = hp["v"] * np.eye(r)
    Q = hp["q"] * np.eye(r)
    R = hp["rho"] * np.eye(d)
    P = hp["p"] * np.eye(r)

    res = dict(error_predict=[], error_full=[], runtime=[], inside_sig=[])
    hashes = dict(Y=[], C=[], X=[])
    missRatio = float("nan")
    if batched:
        from .impute import fit_repeats
        Ys, Cs, Xs, Ms, Mms, Eis = [], [], [], [], [], []
        for i in range(repeats):
            Ymiss = np.copy(Yorig)
            missRatio, missMask = prepare_missing(Ymiss, percentage / 100)
            M = np.array(np.invert(np.isnan(Ymiss)), dtype=int)
            Y = np.copy(Ymiss)
            Y[np.isnan(Y)] = 0
            C = np.random.rand(d, r)
            X = np.random.rand(r, T)
            hashes["Y"].append(matrix_hash(Y)); hashes["C"].append(matrix_hash(C)); hashes["X"].append(matrix_hash(X))
            Ys.append(Y); Cs.append(C); Xs.append(X); Ms.append(M); Mms.append(missMask)
            Eis.append(rmsem(C @ X, YorigInt, missMask))
        eps, efs, rts, ibs = fit_repeats(Ys, Cs, Xs, Ms, Mms, V, Q, R, P, hp["lambda0"], hp["sig"], Iter, YorigInt, Eis,
                                         robust=method == "rPSMF")
        for ep, ef, rt, ib in zip(eps, efs, rts, ibs):
            e_pred, e_full = float(ep[:, Iter].item()), float(ef[:, Iter].item())
            bad = np.isnan(e_pred) or np.isnan(e_full)
            res["error_predict"].append(e_pred)
            res["error_full"].append(e_full)
            res["runtime"].append(float("nan") if bad else float(rt[:, Iter].item()))
            res["inside_sig"].append(float("nan") if bad else float(ib))
        repeats = 0
    for i in range(repeats):
        Ymiss = np.copy(Yorig)
        missRatio, missMask = prepare_missing(Ymiss, percentage / 100)
        M = np.array(np.invert(np.isnan(Ymiss)), dtype=int)
        Y = np.copy(Ymiss)
        Y[np.isnan(Y)] = 0
        C = np.random.rand(d, r)
        X = np.random.rand(r, T)
        hashes["Y"].append(matrix_hash(Y)); hashes["C"].append(matrix_hash(C)); hashes["X"].append(matrix_hash(X))
        Einit = rmsem(C @ X, YorigInt, missMask)
        if method == "rPSMF":                                  # rPSMF.py:214-232
            ep, ef, rt, ib = fit(Y, C, X, d, T, r, M, missMask, V, Q, R, P, hp["lambda0"], hp["sig"], Iter, YorigInt, Einit)
        else:                                                  # PSMF.py:160-178 (the unused `lam` argument is 0 there)
            ep, ef, rt, ib = fit(Y, C, X, d, T, r, M, missMask, 0, V, Q, R, P, hp["sig"], Iter, YorigInt, Einit)
        e_pred, e_full = float(ep[:, Iter].item()), float(ef[:, Iter].item())
        bad = np.isnan(e_pred) or np.isnan(e_full)             # rPSMF.py:236-243
        res["error_predict"].append(e_pred)
        res["error_full"].append(e_full)
        res["runtime"].append(float("nan") if bad else float(rt[:, Iter].item()))
        res["inside_sig"].append(float("nan") if bad else float(ib))
        if log:
            log("finished repeat %d of %d" % (i + 1, repeats))
    params = {k: hp[k] for k in ("r", "sig", "rho", "q", "p", "Iter", "v")}
    if method == "rPSMF":
        params["lambda0"] = hp["lambda0"]
    return dict(method=method, seed=int(seed), missing_percentage=percentage, missing_ratio=float(missRatio),
                parameters=params, hashes=hashes, results=res)


def main(argv=None):
    ap = argparse.ArgumentParser(description="PSMF / rPSMF imputation experiment on the GPU (ExperimentImpute/*.py main)")
    ap.add_argument("-i", "--input", required=True, help="CSV file, one row per series (d x T), empty = missing")
    ap.add_argument("-o", "--output", help="JSON result file (default: stdout)")
    ap.add_argument("-m", "--method", default="rPSMF", choices=["rPSMF", "PSMF"])
    ap.add_argument("-p", "--percentage", type=int, default=30, help="percentage of missing entries to create")
    ap.add_argument("-s", "--seed", type=int, default=None)
    ap.add_argument("-r", "--repeats", type=int, default=1)
    args = ap.parse_args(argv)
    Yorig = np.genfromtxt(args.input, delimiter=",")
    out = run_impute_experiment(Yorig, args.method, args.percentage, args.seed, args.repeats, log=print)
    out["dataset"] = args.input
    out["hostname"] = socket.gethostname()
    text = json.dumps(out, indent=1)
    if args.output:
        with open(args.output, "w") as fp:
            fp.write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
