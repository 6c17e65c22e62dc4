#!/bin/bash
# same-box A/B: old barrier-synchronised pipelined kernel (commit 961a94e) vs the current tree
fmt='import json,sys; j=json.loads(sys.stdin.read()); print("%s d=%d window=%s %.0f steps/s  %.2f us/step  frac=%.3f" % (sys.argv[1], j["config"]["d"], sys.argv[2], j["value"], 1e6/j["value"], j["roofline"]["frac"]))'
for w in 250 1000; do
  (cd scratch/old961 && python bench.py --d 1000000 --T 2000 --window $w --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "$fmt" old $w)
  python bench.py --d 1000000 --T 2000 --window $w --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "$fmt" new $w
  PSMF_NPW=14 python bench.py --d 1000000 --T 2000 --window $w --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "$fmt" new-npw14 $w
done
(cd scratch/old961 && python bench.py --d 125024 --T 4000 --window 500 --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "$fmt" old 500)
python bench.py --d 125024 --T 4000 --window 500 --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "$fmt" new 500
