#!/bin/bash
# full GPU suite on the product library, then same-box A/B: scratch/exp3.sh lib1 lib2 ...
mkdir -p gpurun_out
export PSMF_SPIN_TIMEOUT_MS=900000
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02e_gputests.log; tail -3 gpurun_out/r02e_gputests.log
unset PSMF_SPIN_TIMEOUT_MS
run() { python bench.py --no-e2e --no-cpu --parity-steps 0 "$@" 2>/tmp/err.log | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); print('   %-50s %.4g /s  %.3f us/step  frac=%.3f sm=%s' % ('$*', j['value'], 1e6/j['value'], j['roofline']['frac'], j['clocks']['sm_mhz']))
except Exception as e:
    print('   failed: $*', e, open('/tmp/err.log').read()[-300:])"; }
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" = main ]; then unset PSMF_B200_LIB; else export PSMF_B200_LIB=$PWD/scratch/libs/lib$lib.so; fi
  echo "== $lib (rep $rep)"
  run --rows 125024 --T 4000 --steps 6 --warmup 3
  run --workload B --series 512 --T 1500 --steps 6 --warmup 3
  if [ $rep = 1 ]; then run --rows 250016 --T 4000 --steps 4 --warmup 2; run --steps 20 --warmup 5; fi
done
done
