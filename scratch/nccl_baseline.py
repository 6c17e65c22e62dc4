"""Timing baseline for §8(e): what a one-launch-per-step design with an NCCL all-reduce of the statistics costs.

Every rank filters its own shard with a world_size = 1 engine, ONE step per launch, and all-reduces a
173-double tensor after every launch (the payload of the real exchange).  The numbers are not a filter result
(the statistics are not actually combined) -- this only measures launch + NCCL latency per step, the floor
of any host-driven design, next to the in-kernel NVLink mailbox of the product path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scratch/nccl_baseline.py
"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data as bench
from rpsmf_b200 import FilterEngine, shard_rows

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
d = int(os.environ.get("ROWS_PER_GPU", "125024")) * world
r, T = 16, 400
b, e = shard_rows(d, world, rank)
Y, M, C0, x0 = bench.make_series(torch, dev, e - b, b, d, r, T, torch.float64)
init = bench.init_state(r)

def timed(fn, n):
    for _ in range(20): fn(0)
    dist.barrier(); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for t in range(n): fn(t)
    t1.record(); torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) * 1e3 / n

eng = FilterEngine(e - b, r, robust=True, device=lr)
eng.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
stats = torch.zeros(173, dtype=torch.float64, device=dev)
def step_nccl(t):
    eng.run(Y[t:t + 1], M[t:t + 1], want_X=False)
    dist.all_reduce(stats)
def step_launch_only(t):
    eng.run(Y[t:t + 1], M[t:t + 1], want_X=False)
def allreduce_only(t):
    dist.all_reduce(stats)
us_nccl = timed(step_nccl, T - 1)
us_launch = timed(step_launch_only, T - 1)
us_ar = timed(allreduce_only, T - 1)
eng.close()
eng2 = FilterEngine(e - b, r, robust=True, device=lr, d_global=d, world_size=world, rank=rank)
eng2.connect(dist)
eng2.set_state(C_=C0, V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
def run_all(_):
    eng2.run(Y, M, want_X=False)
us_fused = timed(run_all, 5) / T
if rank == 0:
    print("N=%d rows/GPU=%d: one launch per step + NCCL all-reduce %.1f us/step (launch only %.1f, all-reduce only %.1f); "
          "persistent kernel with in-kernel NVLink mailbox %.2f us/step" % (world, e - b, us_nccl, us_launch, us_ar, us_fused))
eng2.close()
dist.destroy_process_group()
