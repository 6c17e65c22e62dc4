import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from synth import impute_init, make_problem
from rpsmf_b200 import FilterEngine
d, r, T = int(sys.argv[1]), 16, int(sys.argv[2])
Y, M, C0, x0 = make_problem(d, r, T, seed=1)
init = impute_init(r)
Yd, Md = torch.as_tensor(Y).cuda(), torch.as_tensor(M).cuda()
res = {}
for k in (1, 2):
    eng = FilterEngine(d, r, robust=True, kernel=k)
    eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
    out = eng.run(Yd, Md, want_X=True)
    print("kernel", k, "status", eng.status(), eng.launch_info())
    res[k] = (out["X"].cpu().numpy(), eng.get_state()["C"].cpu().numpy())
    eng.close()
from oracle import psmf_oracle_c as pc
pc.use_all_cores()
ref = pc.run(C0, x0, init["P"], init["V"], init["Q"], init["rho"], init["lam"], Y, M, robust=True, cupdate_vt=True)
ref1 = pc.run(C0, x0, init["P"], init["V"], init["Q"], init["rho"], init["lam"], Y[:T - 1], M[:T - 1], robust=True, cupdate_vt=True)
for k in (1, 2):
    rd = np.abs(res[k][1] - ref["C"]).max(axis=1) / np.abs(ref["C"]).max()
    badk = np.nonzero(rd > 1e-9)[0]
    print("kernel", k, "vs oracle: rows differing", badk.size, "X rel", float(np.abs(res[k][0] - ref["X"]).max() / np.abs(ref["X"]).max()))
    if badk.size:
        Ck = res[k][1][badk]
        print("   of these: equal to C after T-1 steps (last update lost):", int((np.abs(Ck - ref1["C"][badk]).max(axis=1) < 1e-9).sum()),
              " equal to C0:", int((np.abs(Ck - C0[badk]).max(axis=1) < 1e-12).sum()))
        d_last = ref["C"][badk] - ref1["C"][badk]; d_first = ref1["C"][badk] - C0[badk]
        print("   equal to C0 + last update only (first update lost):", int((np.abs(Ck - (C0[badk] + d_last)).max(axis=1) < 1e-9).sum()),
              " equal to final + extra last update:", int((np.abs(Ck - (ref["C"][badk] + d_last)).max(axis=1) < 1e-9).sum()))
X1, C1 = res[1]; X2, C2 = res[2]
print("X per-step rel diff:", (np.abs(X1 - X2).max(axis=1) / np.abs(X1).max()).tolist())
rowdiff = np.abs(C1 - C2).max(axis=1) / np.abs(C1).max()
bad = np.nonzero(rowdiff > 1e-9)[0]
print("rows differing:", bad.size, "of", d)
if bad.size:
    ntiles = (d + 31) // 32; cps = 147
    tiles = np.unique(bad // 32)
    print("tiles differing:", tiles.size, "first", tiles[:12], "last", tiles[-12:])
    tb = np.array([ntiles * c // cps for c in range(cps + 1)])
    cta = np.searchsorted(tb, tiles, side="right") - 1
    within = tiles - tb[cta]
    print("CTAs involved:", np.unique(cta).size, "first", np.unique(cta)[:10])
    print("tile-in-CTA histogram (min, max, unique count):", within.min(), within.max(), np.unique(within).size, np.unique(within)[:20])
    print("max row diff", rowdiff.max())
