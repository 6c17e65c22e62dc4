import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import psmf_oracle as po
from synth import impute_init, make_problem
from rpsmf_b200 import FilterEngine
np.set_printoptions(linewidth=200, precision=6)
for (d, r, T) in [(203, 1, 3), (203, 2, 3), (64, 8, 3)]:
    Y, M, C0, x0 = make_problem(d, r, T, seed=r)
    init = impute_init(r)
    eng = FilterEngine(d, r, robust=True)
    eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
    print("state after set", {k: (v.cpu().numpy() if k != 'C' else v.cpu().numpy()[:2]) for k, v in eng.get_state().items()})
    out = eng.run(torch.as_tensor(Y).cuda(), torch.as_tensor(M).cuda(), want_X=True, want_Yrec=True, want_scal=True)
    print("status", eng.status(), eng.launch_info())
    st = po.OracleState(C0.copy(), x0.copy(), init["P"], init["V"], init["Q"], init["rho"], init["lam"])
    rec = []
    ost, oX, oYrec, oscal = po.run(st, po.OracleConfig(robust=True), Y, M.astype(float), record=rec)
    print("x0", x0)
    print("gpu X\n", out["X"].cpu().numpy()); print("ref X\n", oX)
    print("gpu scal\n", out["scal"].cpu().numpy()); print("ref scal\n", oscal)
    print("gpu yrec", out["Yrec"].cpu().numpy()[0, :6]); print("ref yrec", oYrec[0, :6])
    print("ref stats step0", {k: (v if np.isscalar(v) else np.asarray(v).ravel()[:4]) for k, v in rec[0]["stats"].items()})
    eng.close()
