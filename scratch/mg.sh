#!/bin/bash
# usage: scratch/mg.sh NGPUS d [d ...]
N=$1; shift
for d in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 --T 4000 --rows $d --no-cpu --no-e2e > /tmp/mg.out 2>&1
  tail -1 /tmp/mg.out | python -c '
import json,sys
try:
    j=json.loads(sys.stdin.read()); print("N=%d d=%d rows/gpu=%d %.0f steps/s %.2f us/step %s" % (j["n_gpus"], j["config"]["d"], j["config"]["rows_per_gpu"], j["value"], 1e6/j["value"], j["launch"]))
except Exception as e:
    print("bench failed:", e)' || true
  grep -E "Error|error" /tmp/mg.out | head -5
done
