"""Seeded synthetic inputs shared by the tests and bench.py (SURVEY.md 8(d) generator, host version)."""
import numpy as np


def make_problem(d, r, T, seed=0, missing=0.2, student=True, q=0.01, var=0.1, S=None):
    rng = np.random.RandomState(seed)
    n = 1 if S is None else S
    out = []
    for _ in range(n):
        Ct = rng.randn(d, r)
        x = rng.randn(r)
        Y = np.zeros((T, d))
        for t in range(T):
            x = x + np.sqrt(q) * rng.randn(r)
            noise = rng.standard_t(3, d) if student else rng.randn(d)
            Y[t] = Ct @ x + np.sqrt(var) * noise
        M = (rng.rand(T, d) >= missing).astype(np.uint8)
        Y = Y * M
        C0 = rng.rand(d, r)
        x0 = rng.rand(r)
        out.append((Y, M, C0, x0))
    if S is None:
        return out[0]
    return tuple(np.stack([o[k] for o in out]) for k in range(4))


def impute_init(r):
    """V=2I, Q=0.1I, rho=10, P=I, lambda0=1.8 (rPSMF.py:170-183)."""
    return dict(V=2.0 * np.eye(r), Q=0.1 * np.eye(r), rho=10.0, P=np.eye(r), lam=1.8)
