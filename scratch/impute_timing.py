"""Wall time of one imputation fit at the reference's own sizes (config 3, SURVEY 8(d) "I-shaped"):
the unmodified reference on the host CPU (only where /root/reference exists) or the drop-in functions on the GPU.

    python scratch/impute_timing.py reference      # this container (no GPU): ExperimentImpute/{rPSMF,PSMF}.py
    python scratch/impute_timing.py b200           # GPU box: rpsmf_b200.robust_PSMF / ProbabilisticSequentialMatrixFactorizer
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import impute_case, load_golden

impl = sys.argv[1]
if impl == "reference":
    from oracle import ref_loader
    os.chdir("/tmp")
    fr = ref_loader.impute_module("rPSMF").robust_PSMF.func
    fp = ref_loader.impute_module("PSMF").ProbabilisticSequentialMatrixFactorizer.func
else:
    from rpsmf_b200 import ProbabilisticSequentialMatrixFactorizer as fp, robust_PSMF as fr

for name in ("impute_pm25_30", "impute_pm10_head_20", "impute_sp500_head_30"):
    g = load_golden(name)
    c = impute_case(g)
    r = c["r"]; d, n = c["Y"].shape
    V, Q, R, P = 2 * np.eye(r), 0.1 * np.eye(r), 10 * np.eye(d), np.eye(r)
    for method in ("rPSMF", "PSMF"):
        best = 1e30
        for rep in range(3 if impl == "b200" else 1):
            X = c["X0"].copy(); C = c["C0"].copy()
            Einit = float(g["rep0_%s_Einit" % method])
            t0 = time.perf_counter()
            if method == "rPSMF":
                out = fr(c["Y"], C, X, d, n, r, c["M"], c["Mmiss"], V.copy(), Q.copy(), R.copy(), P.copy(), 1.8, 2, c["Iter"], c["YorigInt"], Einit)
            else:
                out = fp(c["Y"], C, X, d, n, r, c["M"], c["Mmiss"], 0, V.copy(), Q.copy(), R.copy(), P.copy(), 2, c["Iter"], c["YorigInt"], Einit)
            best = min(best, time.perf_counter() - t0)
        steps = n * c["Iter"]
        print("%-9s %-22s d=%4d n=%5d Iter=%d  %8.3f s per fit  %9.0f filter steps/s  Efull=%.6f"
              % (impl, name + "/" + method, d, n, c["Iter"], best, steps / best, float(out[1][0, -1])), flush=True)
