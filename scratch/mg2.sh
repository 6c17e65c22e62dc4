#!/bin/bash
# scratch/mg2.sh N tag : multi-GPU validation + measurements on N GPUs of one box (outputs under gpurun_out/)
N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tests/multi_gpu_check.py > gpurun_out/${TAG}_multi_gpu_check_n$N.log 2>&1; echo "multi_gpu_check rc=$? OK lines: $(grep -c ' OK' gpurun_out/${TAG}_multi_gpu_check_n$N.log)"
$TR --master-port 29512 bench.py --gpus $N > gpurun_out/${TAG}_bench_L_n$N.json 2> gpurun_out/${TAG}_bench_L_n$N.err; echo "bench L rc=$?"
for v in $VARIANTS; do  # optional kernel variants (scratch/libs)
  PSMF_B200_LIB=$PWD/scratch/libs/lib$v.so $TR --master-port 29513 bench.py --gpus $N --steps 10 --no-e2e --no-cpu > gpurun_out/${TAG}_bench_L_n${N}_$v.json 2>/dev/null; echo "variant $v rc=$?"
done
$TR --master-port 29514 bench.py --gpus $N --workload B --steps 6 --T 1500 > gpurun_out/${TAG}_bench_B_n$N.json 2> gpurun_out/${TAG}_bench_B_n$N.err; echo "bench B rc=$?"
$TR --master-port 29515 bench.py --gpus $N --impl nccl --steps 3 --window 100 > gpurun_out/${TAG}_bench_nccl_n$N.json 2> gpurun_out/${TAG}_bench_nccl_n$N.err; echo "bench nccl rc=$?"
nvidia-smi topo -m > gpurun_out/${TAG}_topo_n$N.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*_n$N*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value=%.4g" % (j.get("value") or 0), j.get("unit"), "ms/step=%.4g" % j.get("ms_per_step",0), "parity=", (j.get("parity") or {}).get("ok"), "e2e=", (j.get("e2e") or {}).get("value"), "frac=", (j.get("roofline") or {}).get("frac"), (j.get("roofline") or {}).get("bound"))
    except Exception as e:
        print(f, "unreadable", e)
PY
