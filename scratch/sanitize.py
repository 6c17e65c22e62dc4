"""Small parity run of every kernel for compute-sanitizer (racecheck / synccheck / memcheck):
    PSMF_SPIN_TIMEOUT_MS=600000 compute-sanitizer --tool racecheck python scratch/sanitize.py [stream|direct|batch ...]"""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import psmf_oracle as po
from synth import impute_init, make_problem
from rpsmf_b200 import FilterEngine
which = sys.argv[1:] or ["stream", "direct", "batch"]
cases = dict(stream=(20000, 16, 5, 1, 2), direct=(20000, 16, 5, 1, 1), batch=(512, 8, 6, 3, 3), stream_resident=(1600, 16, 5, 1, 2))
for name in which:
    d, r, T, S, kernel = cases[name]
    Y, M, C0, x0 = make_problem(d, r, T, seed=3, S=None if S == 1 else S)
    init = impute_init(r)
    eng = FilterEngine(d, r, n_series=S, robust=True, kernel=kernel, ctas=2 if name == "stream_resident" else 0)
    eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
    out = eng.run(torch.as_tensor(Y).cuda(), torch.as_tensor(M).cuda(), want_X=True)
    assert eng.status() == -1
    X = out["X"].cpu().numpy().reshape(S, T, r)
    errs = []
    for s in range(S):
        ost = po.OracleState((C0 if S == 1 else C0[s]).copy(), (x0 if S == 1 else x0[s]).copy(), init["P"], init["V"], init["Q"], init["rho"], init["lam"])
        ost, oX, _, _ = po.run(ost, po.OracleConfig(robust=True), Y if S == 1 else Y[s], (M if S == 1 else M[s]).astype(float))
        errs.append(float(np.max(np.abs(X[s] - oX)) / np.max(np.abs(oX))))
    print("%s: kernel %s, d=%d r=%d T=%d S=%d, max rel err vs oracle %.2e" % (name, eng.launch_info(), d, r, T, S, max(errs)), flush=True)
    assert max(errs) < 1e-9
    eng.close()
print("sanitize.py: all cases ok")
