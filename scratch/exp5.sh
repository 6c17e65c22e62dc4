#!/bin/bash
# early rank-1 phase (resident regime): GPU suite + sanitizers on the product library, then same-box A/B against libe0 (off)
mkdir -p gpurun_out
export PSMF_SPIN_TIMEOUT_MS=900000
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02g_gputests.log; tail -3 gpurun_out/r02g_gputests.log
for tool in racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 30 python scratch/sanitize.py stream_resident stream > gpurun_out/r02g_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -2 gpurun_out/r02g_sanitizer_$tool.log
done
unset PSMF_SPIN_TIMEOUT_MS
run() { python bench.py --no-e2e --no-cpu "$@" 2>/tmp/err.log | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); print('   %-50s %.5g /s  %.3f us/step  frac=%.3f parity=%s sm=%s' % ('$*', j['value'], 1e6/j['value'], j['roofline']['frac'], (j.get('parity') or {}).get('ok'), j['clocks']['sm_mhz']))
except Exception as e:
    print('   failed: $*', e, open('/tmp/err.log').read()[-300:])"; }
for rep in 1 2; do
for lib in e0 main; do
  if [ "$lib" = main ]; then unset PSMF_B200_LIB; else export PSMF_B200_LIB=$PWD/scratch/libs/lib$lib.so; fi
  echo "== $lib (rep $rep)"
  run --rows 125024 --T 4000 --steps 6 --warmup 3
  if [ $rep = 1 ]; then run --rows 62528 --T 4000 --steps 6 --warmup 3; run --rows 250016 --T 4000 --steps 4 --warmup 2 --parity-steps 0; run --steps 10 --warmup 3 --parity-steps 0; fi
done
done
