#!/bin/bash
# same-box A/B incl. DRAM bytes of one launch (100 filter steps at d = 1M): scratch/exp4.sh lib1 lib2 ...
run() { python bench.py --no-e2e --no-cpu --parity-steps 0 "$@" 2>/tmp/err.log | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); print('   %-50s %.5g /s  %.3f us/step  frac=%.3f sm=%s' % ('$*', j['value'], 1e6/j['value'], j['roofline']['frac'], j['clocks']['sm_mhz']))
except Exception as e:
    print('   failed: $*', e, open('/tmp/err.log').read()[-300:])"; }
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" = main ]; then unset PSMF_B200_LIB; else export PSMF_B200_LIB=$PWD/scratch/libs/lib$lib.so; fi
  echo "== $lib (rep $rep)"
  run --steps 20 --warmup 5
  run --steps 20 --warmup 5 --mask-encoding bytes
  if [ $rep = 1 ]; then
    run --rows 125024 --T 4000 --steps 6 --warmup 3
    run --rows 500000 --T 2000 --steps 4 --warmup 2
    run --dtype f32 --steps 10 --warmup 3
    timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:psmf_stream -s 2 -c 1 \
        python bench.py --T 400 --window 100 --steps 2 --warmup 3 --no-e2e --no-cpu --parity-steps 0 2>&1 | grep -E "dram__bytes|gpu__time" | sed 's/^/   ncu /'
  fi
done
done
