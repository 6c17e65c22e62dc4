import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def relerr(a, b):
    """Norm-wise relative error max|a-b| / max|b| (the 1e-9 bar is norm-wise per quantity)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = float(np.max(np.abs(b))) if b.size else 0.0
    if den == 0.0:
        den = 1.0
    return float(np.max(np.abs(a - b))) / den if a.size else 0.0


def impute_case(g, rep=0):
    """Rebuild the reference call arguments of one fixture repeat (rPSMF.py:196-205)."""
    Yorig = g["Yorig"]
    Mmiss = g["rep%d_Mmiss" % rep].astype(np.float64)
    Ymiss = Yorig.copy()
    Ymiss[Mmiss == 1] = np.nan
    M = (~np.isnan(Ymiss)).astype(np.int64)
    Y = Ymiss.copy()
    Y[np.isnan(Y)] = 0
    YorigInt = Yorig.copy()
    YorigInt[np.isnan(YorigInt)] = 0
    return dict(Y=Y, M=M, Mmiss=Mmiss, YorigInt=YorigInt, C0=g["rep%d_C0" % rep], X0=g["rep%d_X0" % rep],
                Iter=int(g["Iter"]), r=int(g["r"]))
