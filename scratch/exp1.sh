#!/bin/bash
mkdir -p gpurun_out
export PSMF_SPIN_TIMEOUT_MS=900000
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02d_gputests.log; tail -3 gpurun_out/r02d_gputests.log
unset PSMF_SPIN_TIMEOUT_MS
echo "== pipe off (libx0)"; PSMF_B200_LIB=$PWD/scratch/libs/libx0.so bash scratch/quick3.sh 2>&1
echo "== pipe on (main)"; bash scratch/quick3.sh 2>&1
run() { python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu "$@" 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$*', '%.0f steps/s frac=%.3f e2e=%.0f parity=%s clocks=%s' % (j['value'], j['roofline']['frac'], j['e2e']['value'], j['parity']['ok'], j['clocks']))"; }
echo "== sustained, pipe off"; PSMF_B200_LIB=$PWD/scratch/libs/libx0.so run
echo "== sustained, pipe on"; run
run --mask-encoding nan
