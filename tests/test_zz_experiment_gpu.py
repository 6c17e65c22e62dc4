"""The experiment driver end to end on the GPU (default fit = the CUDA model functions) against the same driver
with the oracle as the fit.  Runs last (file name) so that a problem here cannot hide the parity tests under -x."""

import numpy as np
import pytest

from rpsmf_b200 import experiment as ex
from test_experiment_cpu import _oracle_fit

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", ["rPSMF", "PSMF"])
def test_experiment_driver_gpu_vs_oracle(method):
    rng = np.random.RandomState(7)
    d, T, r = 12, 300, 4
    Ct = rng.randn(d, r)
    x = np.cumsum(0.1 * rng.randn(T, r), axis=0)
    Yorig = (x @ Ct.T).T + 0.3 * rng.standard_t(3, (d, T))
    Yorig[rng.rand(d, T) < 0.02] = np.nan                       # a few originally missing entries
    a = ex.run_impute_experiment(Yorig, method, 20, seed=11, repeats=2, r=r)
    b = ex.run_impute_experiment(Yorig, method, 20, seed=11, repeats=2, r=r, fit=_oracle_fit(method == "rPSMF"))
    assert a["hashes"] == b["hashes"] and a["missing_ratio"] == b["missing_ratio"]
    for k in ("error_predict", "error_full"):
        assert np.allclose(a["results"][k], b["results"][k], rtol=1e-9, atol=0)
    assert np.allclose(a["results"]["inside_sig"], b["results"]["inside_sig"], atol=2e-3)
    assert all(t > 0 for t in a["results"]["runtime"])
