#!/bin/bash
run() { python bench.py --d $1 --T $2 --window $3 --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('d=%d T=$2 window=$3  %.0f steps/s  %.2f us/step  frac=%.3f  %s' % (j['config']['d'], j['value'], 1e6/j['value'], j['roofline']['frac'], j['launch']))"; }
run 125024 4000 500
run 250016 4000 500
run 500000 2000 500
run 1000000 2000 500
