"""Row-sharded filter over N GPUs vs the single-process oracle.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

Every rank filters its row shard; statistics are exchanged through the NVLink mailboxes inside the kernel.
Checks: x_t identical (bitwise) on all ranks, and x_t / C / P / V within 1e-9 of the oracle run on all rows.

    --same-device   all ranks share cuda:0 (process group on gloo; the kernels of the ranks time-slice on the GPU and
                    reach each other's mailboxes through CUDA IPC): the N > 1 CUDA path on a one-GPU box
    --quick         small cases only (the same-device mode pays a context switch per exchanged step)
    --desert RANK   rank RANK connects and then leaves without filtering: the survivors' bounded waits must expire and
                    psmf_status must report PSMF_E_STATE instead of hanging (set PSMF_SPIN_TIMEOUT_MS to keep it short)
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import psmf_oracle as po  # noqa: E402
from synth import impute_init, make_problem  # noqa: E402
from rpsmf_b200 import FilterEngine, shard_rows  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--same-device", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--desert", type=int, default=-1)
    args = ap.parse_args()
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    if args.same_device:
        lr = 0
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    if args.desert >= 0:
        return desert(args, rank, world, lr)
    ok = True
    # (d, r, T, kernel): kernel 0 = automatic.  1888 rows = 59 tiles: 29 / 30 tiles on two ranks straddle the automatic
    # threshold of the TMA-staged kernel; 100008 rows leave the last rank a shard with d % 16 != 0 (not eligible):
    # the ranks must agree on ONE kernel (psmf_mailbox_connect), whatever each would pick on its own
    cases = [(6400, 16, 30, 0), (100000, 16, 12, 2), (5000, 8, 20, 1), (777, 3, 25, 0), (1888, 16, 20, 0), (100008, 16, 10, 0)]
    if args.quick:
        cases = [(1888, 16, 8, 0), (3840, 16, 8, 2), (777, 3, 10, 0)]
    for (d, r, T, kernel) in cases:
        Y, M, C0, x0 = make_problem(d, r, T, seed=d)
        init = impute_init(r)
        b, e = shard_rows(d, world, rank)
        eng = FilterEngine(e - b, r, robust=True, device=lr, d_global=d, world_size=world, rank=rank, kernel=kernel)
        eng.connect(dist)
        eng.set_state(C_=C0[b:e], V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
        Yd = torch.as_tensor(np.ascontiguousarray(Y[:, b:e])).cuda(lr)
        Md = torch.as_tensor(np.ascontiguousarray(M[:, b:e])).cuda(lr)
        # two launches: the mailbox step counter carries over
        o1 = eng.run(Yd[: T // 2], Md[: T // 2], k0=1, want_X=True)
        o2 = eng.run(Yd[T // 2:], Md[T // 2:], k0=1 + T // 2, want_X=True)
        assert eng.status() == -1
        X = torch.cat([o1["X"], o2["X"]])
        st = eng.get_state()
        gathered = [torch.empty_like(X) for _ in range(world)]
        dist.all_gather(gathered, X)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        ost = po.OracleState(C0.copy(), x0.copy(), init["P"], init["V"], init["Q"], init["rho"], init["lam"])
        ost, oX, _, _ = po.run(ost, po.OracleConfig(robust=True), Y, M.astype(float))
        errs = dict(X=rel(X.cpu().numpy(), oX), C=rel(st["C"].cpu().numpy(), ost.C[b:e]), P=rel(st["P"].cpu().numpy(), ost.P),
                    V=rel(st["V"].cpu().numpy(), ost.V))
        good = same and max(errs.values()) < 1e-9
        ok = ok and good
        print("rank %d d=%d r=%d kernel=%s rows[%d:%d) replicas_identical=%s errs=%s %s"
              % (rank, d, r, eng.launch_info()["kernel"], b, e, same, {k: "%.1e" % v for k, v in errs.items()}, "OK" if good else "FAIL"),
              flush=True)
        eng.close()
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def desert(args, rank, world, lr):
    """Failure detection: one rank disappears after the mailboxes are connected."""
    from rpsmf_b200 import _capi
    d, r, T = 6400, 16, 10
    Y, M, C0, x0 = make_problem(d, r, T, seed=1)
    init = impute_init(r)
    b, e = shard_rows(d, world, rank)
    eng = FilterEngine(e - b, r, robust=True, device=lr, d_global=d, world_size=world, rank=rank)
    eng.connect(dist)
    if rank == args.desert:
        print("rank %d deserts" % rank, flush=True)
        dist.barrier()
        os._exit(0)
    eng.set_state(C_=C0[b:e], V=init["V"], P=init["P"], x=x0, Q=init["Q"], rho=[init["rho"]], lam=[init["lam"]])
    eng.run(torch.as_tensor(np.ascontiguousarray(Y[:, b:e])).cuda(lr), torch.as_tensor(np.ascontiguousarray(M[:, b:e])).cuda(lr))
    dist.barrier()
    try:
        eng.status()
        print("rank %d: FAIL, no error reported" % rank, flush=True)
        os._exit(1)
    except _capi.PsmfError as ex:
        good = ex.code == -4 and "no progress" in str(ex)
        print("rank %d: %s -> %s" % (rank, ex, "OK" if good else "FAIL"), flush=True)
        os._exit(0 if good else 1)


if __name__ == "__main__":
    main()
