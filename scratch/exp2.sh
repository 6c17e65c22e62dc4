#!/bin/bash
# same-box A/B of kernel variants: scratch/exp2.sh lib1 lib2 ... ("main" = the product library)
run() { python bench.py --no-e2e --no-cpu --parity-steps 0 "$@" 2>/tmp/err.log | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); print('   %-40s %.0f steps/s  %.2f us/step  frac=%.3f sm=%s' % ('$*', j['value'], 1e6/j['value'], j['roofline']['frac'], j['clocks']['sm_mhz']))
except Exception as e:
    print('   failed: $*', e, open('/tmp/err.log').read()[-300:])"; }
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" = main ]; then unset PSMF_B200_LIB; else export PSMF_B200_LIB=$PWD/scratch/libs/lib$lib.so; fi
  echo "== $lib (rep $rep)"
  run --rows 125024 --T 4000 --steps 6 --warmup 3
  if [ $rep = 1 ]; then run --rows 250016 --T 4000 --steps 4 --warmup 2; run --rows 500000 --T 2000 --steps 4 --warmup 2; fi
  run --steps 20 --warmup 5
done
done
