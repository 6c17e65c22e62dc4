#!/bin/bash
# quick untraced timing: steps/s and us/step for the 1-GPU workload and the 8-GPU-shard-sized problem
run() { python bench.py --d $1 --T 1000 --window 250 --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('d=%d  %.0f steps/s  %.2f us/step  frac=%.3f  %s' % (j['config']['d'], j['value'], 1e6/j['value'], j['roofline']['frac'], j['launch']))"; }
for npw in $NPWS; do echo "NPW=$npw"; PSMF_NPW=$npw run 1000000; PSMF_NPW=$npw run 125024; done
