"""Replay EVERY replayable published imputation experiment of the reference on the GPU and compare every repeat.

18 files (LondonAir_PM25, LondonAir_PM10, sp500_closing_prices x {20, 30, 40} % x {PSMF, rPSMF}), 100 repeats each, seed 123:
the PSMF / rPSMF rows of the paper's imputation table for the data sets that ship with the reference (Makefile:160 ff.).
Inputs come from tests/golden/ (tests/golden/make_published_fixture.py); nothing under /root/reference is read.

    python scratch/replay_published.py [name-filter]        # GPU box; writes one line per file
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpsmf_b200 import experiment as ex          # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
pubs = json.load(open(os.path.join(G, "published_results.json")))["files"]
flt = sys.argv[1] if len(sys.argv) > 1 else ""
data = {}


def dataset(name):
    if name not in data:
        f = "impute_pm25_30.npz" if name == "LondonAir_PM25" else "dataset_%s.npz" % name
        data[name] = np.load(os.path.join(G, f))["Yorig"]
    return data[name]


tot_pub = tot_gpu = 0.0
worst = 0.0
nfail = 0
for name, pub in pubs.items():
    if flt not in name:
        continue
    Yorig = dataset(pub["dataset"])
    method = pub["method"]
    hyper = {k: v for k, v in pub["parameters"].items() if k != "lambda0" or method == "rPSMF"}
    reps = len(pub["results"]["error_full"])
    t0 = time.perf_counter()
    out = ex.run_impute_experiment(Yorig, method, pub["missing_percentage"], seed=pub["seed"], repeats=reps, batched=True, **hyper)
    wall = time.perf_counter() - t0
    same_inputs = out["hashes"] == pub["hashes"]
    errs = {}
    for key in ("error_predict", "error_full"):
        a, b = np.asarray(out["results"][key], dtype=float), np.asarray(pub["results"][key], dtype=float)
        errs[key] = float(np.max(np.abs(a - b) / np.abs(b)))
    cov = float(np.max(np.abs(np.asarray(out["results"]["inside_sig"], dtype=float) - np.asarray(pub["results"]["inside_sig"], dtype=float))))
    fit_gpu = float(np.nansum(out["results"]["runtime"]))
    fit_pub = float(np.nansum(pub["results"]["runtime"]))
    # 1e-6: the sp500 series (d = 505, prices of order 1e2 - 1e3) contain repeats on which the reference itself is not
    # reproducible beyond ~3e-7 across machines -- on repeat 18 of sp500_40_rPSMF the UNMODIFIED reference run in the build
    # container differs from its own published value by 3.3e-7 (this library: 2.5e-7; oracle vs that reference run: 7e-8)
    ok = same_inputs and max(errs.values()) < 1e-6 and cov < 1e-4
    above = int(np.sum(np.abs(np.asarray(out["results"]["error_full"], dtype=float) - np.asarray(pub["results"]["error_full"], dtype=float))
                       / np.abs(np.asarray(pub["results"]["error_full"], dtype=float)) > 1e-8))
    nfail += 0 if ok else 1
    worst = max(worst, max(errs.values()))
    tot_pub += fit_pub
    tot_gpu += wall
    print("%-34s d x n = %3d x %4d  %3d repeats  inputs %s  max rel err: error_predict %.1e error_full %.1e  inside_sig max abs diff %.1e  "
          "repeats above 1e-8: %d  published fits %7.1f s  here %6.2f s (fits %.2f s + host-side input generation, hashing, metrics)  %s"
          % (name, Yorig.shape[0], Yorig.shape[1], reps, "identical (hashes)" if same_inputs else "DIFFER", errs["error_predict"],
             errs["error_full"], cov, above, fit_pub, wall, fit_gpu, "OK" if ok else "MISMATCH"), flush=True)
print("total: published fits %.0f s (%.2f h on the authors' machine), here %.1f s; worst relative error %.1e; %d file(s) failed"
      % (tot_pub, tot_pub / 3600, tot_gpu, worst, nfail))
sys.exit(1 if nfail else 0)
