"""Wall time of the synthetic experiments at the reference's own sizes (BASELINE.json configs 1-2, SURVEY 8(d) "S":
d = 20, r = 6, T = 500, cos dynamics, Adam on theta, simplified step), per sweep of T filter steps + predict + Adam update.

    python scratch/synthetic_timing.py reference [sweeps]   # this container (no GPU): ExperimentSynthetic classes, unmodified,
                                                            # autograd replaced by the finite-difference shim of oracle/ref_loader.py
    python scratch/synthetic_timing.py b200 [sweeps]        # GPU box: rpsmf_b200.PSMFIter / rPSMFIter (simplified=True)

Both arms end with the Recursive variants (theta updated inside the sweep, psmf.py:275-331): one pass over the T steps with
update_every = 1 and 10 (full step, not the simplified one: the experiments have no simplified Recursive class).
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

impl = sys.argv[1]
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 20
d, r, T = 20, 6, 500


def cosnl(theta, x, t):
    return np.cos(2 * np.pi * theta * t + x)


def gen(seed, student):
    """Data in the spirit of ExperimentSynthetic/data.py:34-60 (same shapes; the values do not matter for timing)."""
    rng = np.random.RandomState(seed)
    C = rng.randn(d, r)
    theta = 1e-3 * (1 + np.arange(r)).reshape(r, 1)
    x = np.zeros((r, 1))
    Y = np.zeros((T, d))
    for k in range(1, T + 1):
        x = cosnl(theta, x, k)
        noise = rng.standard_t(3, size=d) if student else rng.randn(d)
        Y[k - 1] = (C @ x).reshape(d) + np.sqrt(0.1) * noise
    y = {k: Y[k - 1].reshape(d, 1) for k in range(1, T + 1)}
    return y, rng.rand(d, r), 1e-3 * rng.rand(r, 1)


if impl == "reference":
    from oracle import ref_loader
    os.chdir("/tmp")
    classes = {False: ref_loader.synthetic_module("synthetic_psmf").PSMFIterSynthetic, True: ref_loader.synthetic_module("synthetic_rpsmf").rPSMFIterSynthetic}
    kw = {}
else:
    from rpsmf_b200 import PSMFIter, rPSMFIter

    def _mk(base):
        class Synthetic(base):
            def step_reset(self):                      # synthetic_psmf.py:78-81
                super().step_reset()
                self._V = {0: self.V0}
        return Synthetic
    classes = {False: _mk(PSMFIter), True: _mk(rPSMFIter)}
    kw = dict(simplified=True)

for student, seed in ((False, 35853), (True, 35833)):
    y, C0, theta0 = gen(seed, student)
    V0 = 0.1 * np.eye(r); mu0 = np.zeros((r, 1)); P0 = np.zeros((r, r))
    if student:
        o = classes[True](theta0, C0, V0, mu0, P0, 0 * np.eye(r), np.eye(d), 1.8, cosnl, **kw)
    else:
        o = classes[False](theta0, C0, V0, mu0, P0, {k: 0 * np.eye(r) for k in range(T + 1)}, {k: np.eye(d) for k in range(T + 1)}, cosnl, **kw)
    o.adam_init(gam=1e-3)
    times = []
    for i in range(1, n_iter + 1):
        t0 = time.perf_counter()
        o.step(y, i, T)
        o.predict(i, T, 10)
        o.adam_update(i)
        times.append(time.perf_counter() - t0)
    best = float(np.median(times[1:])) if n_iter > 1 else times[0]
    print("%-9s %-6s d=%d r=%d T=%d  %8.2f ms per sweep (median of %d)  %9.0f filter steps/s  theta[0]=%.9f"
          % (impl, "rPSMF" if student else "PSMF", d, r, T, best * 1e3, n_iter - 1, T / best, float(np.asarray(o._theta[n_iter]).reshape(-1)[0])), flush=True)
    if hasattr(o, "close"):
        o.close()


# ---- Recursive variants: one pass, theta updated every `update_every` steps ----
if impl == "reference":
    mod = ref_loader.pypsmf()
    rec = {False: mod.PSMFRecursive, True: mod.rPSMFRecursive}
else:
    from rpsmf_b200 import PSMFRecursive, rPSMFRecursive
    rec = {False: PSMFRecursive, True: rPSMFRecursive}
for student, seed in ((False, 35853), (True, 35833)):
    y, C0, theta0 = gen(seed, student)
    V0 = 0.1 * np.eye(r); mu0 = 0.3 * np.ones((r, 1)); P0 = 0.5 * np.eye(r); Q = 0.01 * np.eye(r)
    for ue in (1, 10):
        best = 1e30
        for rep in range(3 if impl == "b200" else 1):
            if student:
                o = rec[True](theta0, C0, V0, mu0, P0, Q, np.eye(d), 1.8, cosnl)
            else:
                o = rec[False](theta0, C0, V0, mu0, P0, {k: Q for k in range(T + 1)}, {k: np.eye(d) for k in range(T + 1)}, cosnl)
            t0 = time.perf_counter()
            o.run(y, T, 10, update_every=ue)
            best = min(best, time.perf_counter() - t0)
            th = float(np.asarray(o._theta[T]).reshape(-1)[0])
            if hasattr(o, "close"):
                o.close()
        print("%-9s %-15s update_every=%-2d d=%d r=%d T=%d  %8.2f ms per pass  %9.0f filter steps/s  theta[0]=%.9f"
              % (impl, "rPSMFRecursive" if student else "PSMFRecursive", ue, d, r, T, best * 1e3, T / best, th), flush=True)
