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
t = tr.cpu().numpy().astype(np.int64)[20:]
us = lambda a: a.mean() / 1e3
print("d=%d kernel=%d  step mean %.2f us" % (d, kernel, us(t[1:, 0] - t[:-1, 0])))
if kernel == 2:
    print("  reducer: wait for pass %.2f | CTA partial %.2f | grid barrier(s)+reduce %.2f (first barrier %.2f)"
          % (us(t[:, 1] - t[:, 0]), us(t[:, 2] - t[:, 1]), us(t[:, 5] - t[:, 2]), us(t[:, 4] - t[:, 3])))
    print("  solver : solve %.2f | idle before it %.2f" % (us(t[:, 13] - t[:, 12]), us(t[1:, 12] - t[:-1, 13])))
    print("  pass warp 0: pass %.2f us, wait for solve(t-2) %.2f us, waiting for slots %.2f us"
          % (us(t[:, 9] - t[:, 8]), us(t[:, 8] - t[:, 11]), t[:, 10].mean() / 1.9e3))
else:
    print("  pass %.2f | partial+barrier+reduce %.2f | solve %.2f" % (us(t[:, 1] - t[:, 0]), us(t[:, 5] - t[:, 1]), us(t[:, 6] - t[:, 5])))
