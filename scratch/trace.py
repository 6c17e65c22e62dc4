import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench_data as bench
from rpsmf_b200 import FilterEngine
d, r, T = int(sys.argv[1]), 16, 160
kernel = int(sys.argv[2]); ctas = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
Y, M, C0, x0 = bench.make_series(torch, dev, d, 0, d, r, T, torch.float64)
init = bench.init_state(r)
eng = FilterEngine(d, r, robust=True, kernel=kernel, ctas=ctas)
eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
tr = eng.set_trace(T)
eng.run(Y, M, want_X=False); eng.run(Y, M, want_X=False)
print(eng.status(), eng.launch_info())
raw = tr.cpu().numpy().astype(np.int64)
t = raw[:T * 16].reshape(T, 16)[20:]
per_cta = raw[T * 16:T * 176].reshape(160, T)
per_cta_start = raw[T * 176:].reshape(160, T)
us = lambda a: a.mean() / 1e3
print("d=%d kernel=%d  step mean %.2f us" % (d, kernel, us(t[1:, 0] - t[:-1, 0])))
if kernel == 2:
    print("  control reducers: poll totals %.2f | exchange %.2f" % (us(t[:, 4] - t[:, 2]), us(t[:, 13] - t[:, 4])))
    print("  solvers: wait for stats %.2f | assemble+fill+eliminate %.2f | x, publish %.2f | rest of update + predict %.2f   (critical: stats->publish %.2f)"
          % (us(t[:, 1] - t[:, 0]), us(t[:, 7] - t[:, 1]), us(t[:, 6] - t[:, 7]), us(t[:, 12] - t[:, 6]), us(t[:, 6] - t[:, 1])))
    print("  chain: CTA 0 released -> stats ready %.2f -> published %.2f ; publish(t) -> CTA 0 released (t+2) %.2f"
          % (us(t[:, 13] - t[:, 9]), us(t[:, 6] - t[:, 13]), us(t[2:, 9] - t[:-2, 6])))
    print("  pass warp 0: start -> CTA partial released %.2f us (tiles %.2f | write sums %.2f | reduce warp: CTA sum, store, release %.2f), wait for params %.2f us, waiting for slots %.2f us"
          % (us(t[:, 9] - t[:, 8]), us(t[:, 14] - t[:, 8]), us(t[:, 15] - t[:, 14]), us(t[:, 9] - t[:, 15]), us(t[:, 8] - t[:, 11]), t[:, 10].mean() / 1.9e3))
    print("  data CTA 0: params seen - published %.2f" % (us(t[2:, 8] - t[:-2, 6])))
else:
    print("  pass %.2f | partial+barrier+reduce %.2f | solve %.2f" % (us(t[:, 1] - t[:, 0]), us(t[:, 5] - t[:, 1]), us(t[:, 6] - t[:, 5])))

if kernel == 2:
    nc = eng.launch_info()["ctas"] - 1
    smid = per_cta[159, :nc + 1].copy()
    pc = per_cta[:nc, 20:]
    dur = np.diff(pc, axis=1)            # per-CTA period between consecutive pass ends
    ends = pc - pc.min(axis=0, keepdims=True)
    print("  per-CTA pass period: mean %.2f us; lateness of pass end vs the earliest CTA: mean %.2f, max %.2f us"
          % (dur.mean() / 1e3, ends.mean() / 1e3, ends.max(axis=0).mean() / 1e3))
    late = ends.mean(axis=1) / 1e3
    order = np.argsort(late)
    print("  earliest CTAs", order[:8], late[order[:8]].round(1), " latest CTAs", order[-8:], late[order[-8:]].round(1))
    print("  SM ids: control CTA on SM %d; latest CTAs on SMs %s; earliest on %s" % (smid[nc], smid[order[-4:]], smid[order[:4]]))
    srt = np.argsort(late)
    print("  lateness by SM id (sorted):", [(int(smid[i]), round(float(late[i]), 1)) for i in srt[-12:]])
    # second launch in the same process: are the same SMs late?
    eng.run(Y, M, want_X=False); eng.status()
    raw2 = tr.cpu().numpy().astype(np.int64)
    pc2 = raw2[T * 16:T * 176].reshape(160, T)[:nc, 20:]
    late2 = (pc2 - pc2.min(axis=0, keepdims=True)).mean(axis=1) / 1e3
    srt2 = np.argsort(late2)
    print("  second launch            :", [(int(smid[i]), round(float(late2[i]), 1)) for i in srt2[-12:]])
    busy = (per_cta[:nc, 20:] - per_cta_start[:nc, 20:]).mean(axis=1) / 1e3
    sb = np.argsort(busy)
    print("  per-CTA busy pass time (end - start): min %.1f median %.1f max %.1f us; slowest (SM, us): %s"
          % (busy.min(), np.median(busy), busy.max(), [(int(smid[i]), round(float(busy[i]), 1)) for i in sb[-6:]]))
    z = int(np.argmax(late))
    t0 = t[20, 0]
    print("  timeline (us, relative): ctrl[start, arrivals, summed, solved] | cta0 [start,end] | latest cta %d [start,end] | all-CTA end min/max" % z)
    for k in range(20, 28):
        kk = k + 20
        print("   step %d: ctrl %7.1f %7.1f %7.1f %7.1f | cta0 %7.1f %7.1f | ctaZ %7.1f %7.1f | ends %7.1f %7.1f" % (
            kk, (t[k,2]-t0)/1e3, (t[k,4]-t0)/1e3, (t[k,13]-t0)/1e3, (t[k,6]-t0)/1e3,
            (per_cta_start[0,kk]-t0)/1e3, (per_cta[0,kk]-t0)/1e3, (per_cta_start[z,kk]-t0)/1e3, (per_cta[z,kk]-t0)/1e3,
            (per_cta[:nc,kk].min()-t0)/1e3, (per_cta[:nc,kk].max()-t0)/1e3))
