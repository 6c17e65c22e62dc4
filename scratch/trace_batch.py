"""Phase stamps of the resident batch kernel (CTA 0): python scratch/trace_batch.py S d r [kernel]"""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench_data as bd
from rpsmf_b200 import FilterEngine
S, d, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kernel = int(sys.argv[4]) if len(sys.argv) > 4 else 3
T = 200
dev = torch.device("cuda", 0)
Y, M, C0, x0 = bd.make_batch(torch, dev, 0, S, S, d, r, T, torch.float64)
init = bd.init_state(r)
eng = FilterEngine(d, r, n_series=S, robust=True, kernel=kernel)
eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
tr = eng.set_trace(T)
eng.run(Y, M, want_X=False); eng.run(Y, M, want_X=False)
print(eng.status(), eng.launch_info())
t = tr.cpu().numpy().astype(np.int64)[:T * 16].reshape(T, 16)[20:]
us = lambda a: a.mean() / 1e3
print("S=%d d=%d r=%d: step %.2f us | pass %.2f | writeout+sync %.2f | CTA sum+sync %.2f | fill+GJ %.2f | rest of update+predict %.2f"
      % (S, d, r, us(t[1:, 0] - t[:-1, 0]), us(t[:, 1] - t[:, 0]), us(t[:, 2] - t[:, 1]), us(t[:, 5] - t[:, 2]), us(t[:, 7] - t[:, 5]),
         us(t[:, 6] - t[:, 7])))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
eng.set_trace(0)
eng.run(Y, M, want_X=False)
e0.record(); eng.run(Y, M, want_X=False); e1.record(); torch.cuda.synchronize()
print("untraced: %.2f us/step, %.1f M series-steps/s" % (e0.elapsed_time(e1) * 1e3 / T, S * T / e0.elapsed_time(e1) / 1e3))
