import sys, os, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from synth import impute_init, make_problem
from oracle import psmf_oracle as po
from rpsmf_b200 import FilterEngine
d, ctas, r, T = int(sys.argv[1]), int(sys.argv[2]), 16, 6
Y, M, C0, x0 = make_problem(d, r, T, seed=1)
init = impute_init(r)
eng = FilterEngine(d, r, robust=True, ctas=ctas, kernel=2)
eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
out = eng.run(torch.as_tensor(Y).cuda(), torch.as_tensor(M).cuda(), want_X=True)
print("status", eng.status(), eng.launch_info())
st = po.OracleState(C0.copy(), x0.copy(), init["P"], init["V"], init["Q"], init["rho"], init["lam"])
st, X, _, _ = po.run(st, po.OracleConfig(robust=True), Y, M.astype(float))
print("relerr X", float(np.max(np.abs(out["X"].cpu().numpy() - X)) / np.max(np.abs(X))))
