#!/bin/bash
run() { python bench.py --d $1 --T $2 --window $3 --steps $4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('d=%d T=$2 window=$3 steps=$4 %.0f steps/s  %.2f us/step  frac=%.3f clocks=%s' % (j['config']['d'], j['value'], 1e6/j['value'], j['roofline']['frac'], j['clocks']))"; }
run 1000000 2000 100 16
run 1000000 2000 250 8
run 1000000 2000 500 4
run 1000000 2000 2000 2
PSMF_NPW=14 run 1000000 2000 500 4
PSMF_NPW=9 run 1000000 2000 500 4
