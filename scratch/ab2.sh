#!/bin/bash
# sustained (power-capped) same-box A/B: default bench shape, 20 x 500 steps back to back
fmt='import json,sys; j=json.loads(sys.stdin.read()); print("%s %.0f steps/s  %.2f us/step  frac=%.3f clocks=%s" % (sys.argv[1], j["value"], 1e6/j["value"], j["roofline"]["frac"], j["clocks"]))'
(cd scratch/old961 && python bench.py --T 4000 --steps 20 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "$fmt" old)
for npw in 11 12 13 14; do
PSMF_NPW=$npw python bench.py --T 4000 --steps 20 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "$fmt" new-npw$npw
done
(cd scratch/old961 && python bench.py --T 4000 --steps 20 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "$fmt" old)
