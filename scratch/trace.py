import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench
from rpsmf_b200 import FilterEngine
d, r, T = int(sys.argv[1]), 16, 60
kernel = int(sys.argv[2]); ctas = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
Y, M, C0, x0 = bench.make_device_data(torch, dev, d, 0, d, r, T, torch.float64)
init = bench.init_state(r)
eng = FilterEngine(d, r, robust=True, kernel=kernel, ctas=ctas)
eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
tr = eng.set_trace(T)
eng.run(Y, M, want_X=False); eng.run(Y, M, want_X=False)
print(eng.status(), eng.launch_info())
t = tr.cpu().numpy().astype(np.int64)
t = t[20:]
names = ["pass", "cta_sync", "part", "grid_bar", "tot", "gj", "rest_small", "next"]
# stamps: 0 pass start,1 pass end,2 after sync,3 before bar,4 after bar,5 tot done,7 after GJ,6 step end
seq = [0, 1, 2, 3, 4, 5, 7, 6]
dts = np.stack([t[:, seq[i + 1]] - t[:, seq[i]] for i in range(7)], 1)
step = t[1:, 0] - t[:-1, 0]
print("d=%d kernel=%d  step mean %.2f us" % (d, kernel, step.mean() / 1e3))
for n, v in zip(["pass", "cta_sync", "partials", "grid_barrier", "tot_reduce", "gauss_jordan", "small_rest"], dts.mean(0)):
    print("  %-14s %8.2f us" % (n, v / 1e3))

if kernel == 2:
    pw = t[:, 9] - t[:, 8]; pwait = t[:, 8] - t[:, 11]; wf = t[:, 10]
    print("  pass warp 0: pass %.2f us, wait for solve(t-2) %.2f us, cycles waiting for slots %.0f (%.2f us @1.9GHz)" % (pw.mean()/1e3, pwait.mean()/1e3, wf.mean(), wf.mean()/1.9e3))
