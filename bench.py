#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- PSMF/rPSMF filter steps/sec on B200 (BASELINE.json metric).

Workload (config.workload "L"): one series, d = 1,000,000 rows, r = 16, T = 10,000 time steps, rPSMF
(Student-t scales) with 20 % missing entries, fp64, synthetic data generated on the device with a seeded
generator.  One bench "step" is ONE launch of the persistent filter kernel over a window of
`--window` (default 500) consecutive filter steps; the default K = 20 steps cover the T = 10k sequence.
The metric `value` is filter steps per second over all GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W]                    # CUDA arm
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # reference arm: CPU oracle port

Multi-GPU (N > 1, launched under torchrun): rows of C are sharded across ranks (strong scaling at fixed
d), statistics are exchanged per step through NVLink mailboxes.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--d", "--rows", dest="d", type=int, default=1_000_000, help="rows of C (--rows: torchrun's parser trips over --d)")
    ap.add_argument("--r", type=int, default=16)
    ap.add_argument("--T", type=int, default=10_000, help="length of the resident synthetic sequence")
    ap.add_argument("--window", type=int, default=500, help="filter steps per kernel launch (= per bench step)")
    ap.add_argument("--mask", default="iid", choices=["iid", "segments"],
                    help="missing pattern: iid Bernoulli (default) or runs of 20 steps per row (common.py:50-76)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=0, help="filter steps of the CPU baseline sample (0 = auto)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# synthetic data (device-side, seeded): y_t = C_true x_t + sqrt(var) t_3,  x_t random walk, 20 % missing
# --------------------------------------------------------------------------------------------------
def segment_mask(torch, dev, T, d_loc, missing, generator, seg=20):
    """Missing entries in runs of `seg` consecutive time steps per row, one run per row and sweep until the
    requested ratio is reached: the pattern of the imputation experiment (common.py:50-76), time-major (T, d_loc),
    1 = observed."""
    M = torch.ones((T, d_loc), dtype=torch.uint8, device=dev)
    if T <= seg + 1 or missing <= 0:
        return M
    rows = torch.arange(d_loc, device=dev)
    while float(1.0 - M.float().mean()) < missing:
        start = torch.randint(1, T - seg, (d_loc,), generator=generator, device=dev)
        for k in range(seg):
            M[start + k, rows] = 0
    return M


def make_device_data(torch, dev, d_loc, row0, d_total, r, T, dtype, seed=20261017, q=0.01, var=0.1, missing=0.2,
                     chunk=125, mask="iid"):
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 7919 * (row0 // max(1, d_loc) + 1))
    gx = torch.Generator(device="cpu")
    gx.manual_seed(seed)                                    # the latent path is identical on every rank
    Ct = torch.randn((d_loc, r), generator=g, device=dev, dtype=torch.float64)
    x = torch.randn(r, generator=gx, dtype=torch.float64)
    steps = torch.randn((T, r), generator=gx, dtype=torch.float64) * (q ** 0.5)
    Xtrue = (x.unsqueeze(0) + torch.cumsum(steps, 0)).to(dev)
    Y = torch.empty((T, d_loc), dtype=dtype, device=dev)
    M = torch.empty((T, d_loc), dtype=torch.uint8, device=dev)
    Mseg = segment_mask(torch, dev, T, d_loc, missing, g) if mask == "segments" else None
    for a in range(0, T, chunk):
        b = min(T, a + chunk)
        n = b - a
        noise = torch.randn((n, d_loc), generator=g, device=dev, dtype=torch.float32)
        chi = torch.randn((n, d_loc), generator=g, device=dev, dtype=torch.float32).square_()
        chi += torch.randn((n, d_loc), generator=g, device=dev, dtype=torch.float32).square_()
        chi += torch.randn((n, d_loc), generator=g, device=dev, dtype=torch.float32).square_()
        noise.div_(chi.div_(3.0).sqrt_())                  # Student-t, 3 dof (ExperimentSynthetic/data.py:47)
        del chi
        if Mseg is None:
            m = torch.rand((n, d_loc), generator=g, device=dev, dtype=torch.float32) >= missing
        else:
            m = Mseg[a:b] != 0
        yc = Xtrue[a:b] @ Ct.T
        yc.add_(noise.to(torch.float64), alpha=var ** 0.5)
        yc.mul_(m)                                          # zero-filled where missing (rPSMF.py:200-202)
        Y[a:b] = yc.to(dtype)
        M[a:b] = m.to(torch.uint8)
        del noise, m, yc
    gi = torch.Generator(device=dev)
    gi.manual_seed(123 + row0)
    C0 = torch.rand((d_loc, r), generator=gi, device=dev, dtype=torch.float64)
    x0 = torch.rand(r, generator=gx, dtype=torch.float64)
    return Y, M, C0, x0


def init_state(r):
    return dict(V=2.0 * np.eye(r), Q=0.1 * np.eye(r), rho=10.0, P=np.eye(r), lam=1.8)   # rPSMF.py:170-183


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(self.samples))


def algorithmic_bytes_per_filter_step(d_loc, r, esize, masked=True):
    """SURVEY.md 8(d), regime (i): read + write C, read y_t, read the mask byte."""
    return 2 * d_loc * r * esize + d_loc * esize + (d_loc if masked else 0)


def cpu_reference_sample(d, r, nsteps, Yh, Mh, C0h, x0h):
    """Time the CPU port of the reference step on the host cores: the C/OpenMP restatement
    (oracle/psmf_oracle_c.c, all host threads) when it is built, else the numpy restatement."""
    init = init_state(r)
    from oracle import psmf_oracle_c as pc
    if pc.available():
        pc.use_all_cores()
        pc.run(C0h[:1024], x0h, init["P"], init["V"], init["Q"], init["rho"], init["lam"], Yh[:2, :1024], Mh[:2, :1024])
        Cc = np.ascontiguousarray(C0h, dtype=np.float64)
        Yc = np.ascontiguousarray(Yh[:nsteps], dtype=np.float64)
        Mc = np.ascontiguousarray(Mh[:nsteps], dtype=np.uint8)
        t0 = time.perf_counter()
        res = pc.run(Cc, x0h, init["P"], init["V"], init["Q"], init["rho"], init["lam"], Yc, Mc, want_X=False)
        dt = time.perf_counter() - t0
        assert res["bad"] == -1
        return nsteps / dt, pc.threads(), "oracle/psmf_oracle_c.c (C/OpenMP O(d r^2) restatement)"
    from oracle import psmf_oracle as po
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        threads = os.cpu_count() or 1
    st = po.OracleState(C0h.copy(), x0h.copy(), init["P"], init["V"], init["Q"], init["rho"], init["lam"])
    cfg = po.OracleConfig(robust=True)
    st, _ = po.step(st, cfg, Yh[0], Mh[0].astype(np.float64))          # warm-up step (page faults, BLAS threads)
    t0 = time.perf_counter()
    for t in range(1, nsteps):
        st, _ = po.step(st, cfg, Yh[t], Mh[t].astype(np.float64))
    dt = time.perf_counter() - t0
    return (nsteps - 1) / dt, threads, "oracle/psmf_oracle.py (numpy/OpenBLAS O(d r^2) restatement)"


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = args.gpus
    d, r, W = args.d, args.r, args.window

    if args.impl == "reference":
        if rank != 0:
            return
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from rpsmf_b200 import FilterEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    esize = 8 if args.dtype == "f64" else 4

    # rows of this rank (strong scaling at fixed d)
    from rpsmf_b200 import shard_rows
    row0, row1 = shard_rows(d, world, rank)
    d_loc = row1 - row0
    T = max(args.T // W, 1) * W
    nwin = T // W
    Y, M, C0, x0 = make_device_data(torch, dev, d_loc, row0, d, r, T, dtype, mask=args.mask)
    init = init_state(r)
    eng = FilterEngine(d_loc, r, dtype=dtype, robust=True, device=local_rank, d_global=d, world_size=world, rank=rank,
                       ctas=args.ctas)
    if world > 1:
        eng.connect(dist)
    eng.set_state(C_=C0.to(dtype), V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
    Xbuf = torch.empty((1, W, r), dtype=torch.float64, device=dev)

    def one_step(i):
        w = i % nwin
        eng.run(Y[w * W:(w + 1) * W], M[w * W:(w + 1) * W], k0=1 + i * W, want_X=False, X_out=Xbuf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    torch.cuda.profiler.start()          # ncu --profile-from-start off: skip the data generation kernels
    for i in range(args.warmup):
        one_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    for i in range(args.steps):
        one_step(args.warmup + i)
        evs[i + 1].record()
    barrier()
    clocks = sampler.stop()
    torch.cuda.profiler.stop()
    bad = eng.status()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    filter_steps = args.steps * W
    value = filter_steps / (total_ms * 1e-3)
    info = eng.launch_info()

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / mean launch duration
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per_launch = algorithmic_bytes_per_filter_step(d_loc, r, esize) * W
    mean_launch_s = float(np.mean(per_launch_ms)) * 1e-3
    achieved = bytes_per_launch / mean_launch_s / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tr_path) and world == 1 and d == 1_000_000 and r == 16 and args.dtype == "f64":
        try:
            traffic = json.load(open(tr_path)).get("dram_bytes_per_launch_scaled_to_window", {}).get(str(W))
        except Exception:
            traffic = None
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    peak_source="MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                    algorithmic_bytes_per_filter_step=algorithmic_bytes_per_filter_step(d_loc, r, esize),
                    kernel="%s<%d,%s>" % ("psmf_stream_kernel" if info.get("kernel") == "tma" else "psmf_filter_kernel", r, "double" if esize == 8 else "float"),
                    mean_launch_ms=float(np.mean(per_launch_ms)))

    # end-to-end through the public API with HOST buffers (pinned), H2D inside the timed region
    e2e = None
    if not args.no_e2e:
        nw = min(2, nwin)
        Yh = torch.empty((nw * W, d_loc), dtype=dtype).pin_memory()
        Mh = torch.empty((nw * W, d_loc), dtype=torch.uint8).pin_memory()
        Yh.copy_(Y[: nw * W]); Mh.copy_(M[: nw * W])
        eng.run_host(Yh[:W], Mh[:W], window=min(W, 125), k0=1)      # warm-up (allocates the staging buffers)
        barrier()
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        Xh = eng.run_host(Yh, Mh, window=min(W, 125), k0=1)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        e2e = dict(value=nw * W / (ms * 1e-3), unit="filter steps/s",
                   h2d_bytes_per_step=int((d_loc * esize + d_loc) * W), d2h_bytes_per_step=int(W * r * 8),
                   wall_s=time.perf_counter() - t0, note="per bench step of %d filter steps; pinned host Y/M, double-buffered H2D" % W,
                   checksum=float(Xh.sum()))
        del Yh, Mh

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:                      # reported at N = 1 only
        ns = args.cpu_steps or max(8, min(400, int(300e6 / max(d, 1))))     # ~10-20 s of host work at d = 1M
        ns = min(ns, T)
        v, threads, what = cpu_reference_sample(d_loc, r, ns, Y[:ns].double().cpu().numpy(), M[:ns].cpu().numpy(),
                                                C0.cpu().numpy(), x0.numpy())
        cpu = dict(value=v, unit="filter steps/s", cores=threads, kind="port",
                   sample="%d filter steps of the same workload prefix (d=%d rows, r=%d) through %s; the reference's "
                          "d x d form cannot run at d=1M" % (ns, d_loc, r, what))

    if rank == 0:
        line = dict(
            metric="PSMF filter steps/sec at d=1M,r=16", value=value, unit="filter steps/s", n_gpus=world,
            steps=args.steps, warmup=args.warmup, ms_per_step=total_ms / args.steps, higher_is_better=True,
            scaling="strong", vs_baseline=None, dtype=args.dtype, data="synthetic",
            config=dict(workload="L: rPSMF d=%d r=%d T=%d, 20%% missing, %s; one bench step = one kernel launch over %d filter steps"
                                 % (d, r, T, args.dtype, W),
                        d=d, r=r, T=T, window=W, missing=0.2, mask=args.mask, robust=True, rows_per_gpu=d_loc,
                        l2_policy="inputs larger than L2: C (%.0f MB) + Y/M windows (%.1f GB) stream from HBM every step"
                                  % (d_loc * r * esize / 1e6, W * d_loc * (esize + 1) / 1e9),
                        parallelism="rows of C sharded over %d GPU(s)" % world),
            e2e=e2e, gpu_launches=args.steps, roofline=roofline, cpu_baseline=cpu, clocks=clocks,
            launch=info, first_bad_step=bad)
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (oracle port; the d x d reference code cannot run at
    d = 1M) on the host cores, same config / metric / unit.  A bench step is a bounded sample of filter steps."""
    from synth import make_problem
    d, r = args.d, args.r
    est = 0.05 * d / 1e6 + 1e-4                      # seconds per filter step of the C/OpenMP port (8 threads)
    per_step = int(max(1, min(40, 40.0 / ((args.steps + args.warmup) * est))))
    total = per_step                                  # one pool of `per_step` synthetic steps, re-used by every bench step
    rng = np.random.RandomState(20261017)
    Ct = rng.randn(d, r)
    x = rng.randn(r)
    Y = np.empty((total, d)); M = np.empty((total, d), dtype=np.uint8)
    for t in range(total):
        x = x + 0.1 * rng.randn(r)
        M[t] = rng.rand(d) >= 0.2
        Y[t] = (Ct @ x + np.sqrt(0.1) * rng.standard_t(3, d)) * M[t]
    C0 = rng.rand(d, r); x0 = rng.rand(r)
    init = init_state(r)
    from oracle import psmf_oracle_c as pc
    if pc.available():
        threads = pc.use_all_cores()
        what = "oracle/psmf_oracle_c.c (C/OpenMP port, %d threads)" % threads
        st = dict(C=C0, x=x0, P=init["P"], V=init["V"], Q=init["Q"], rho=init["rho"], lam=init["lam"])

        def advance(a, b):
            res = pc.run(st["C"], st["x"], st["P"], st["V"], st["Q"], st["rho"], st["lam"], Y[a:b], M[a:b], want_X=False)
            st.update({k: res[k] for k in ("C", "x", "P", "V", "Q", "rho", "lam")})
    else:
        from oracle import psmf_oracle as po
        try:
            from threadpoolctl import threadpool_info
            threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        except Exception:
            threads = os.cpu_count() or 1
        what = "oracle/psmf_oracle.py (numpy/OpenBLAS port)"
        box = [po.OracleState(C0, x0, init["P"], init["V"], init["Q"], init["rho"], init["lam"])]
        cfg = po.OracleConfig(robust=True)

        def advance(a, b):
            for t in range(a, b):
                box[0], _ = po.step(box[0], cfg, Y[t], M[t].astype(np.float64))
    for _ in range(args.warmup):
        advance(0, per_step)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        advance(0, per_step)
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    sample = "%d filter steps per bench step, d=%d r=%d, %s" % (per_step, d, r, what)
    line = dict(impl="reference", metric="PSMF filter steps/sec at d=1M,r=16", value=v, unit="filter steps/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=dt / args.steps * 1e3,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="L: rPSMF d=%d r=%d, 20%% missing, f64 (CPU sample)" % (d, r), d=d, r=r),
                cpu_baseline=dict(value=v, unit="filter steps/s", cores=threads, kind="port", sample=sample),
                e2e=dict(value=v, unit="filter steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
