#!/bin/bash
# dress rehearsal of the driver's round-end sequence on one GPU: GPU tests, smoke(), reference arm, default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_gputests.log; tail -3 gpurun_out/r02_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_L_reference.json 2> gpurun_out/r02_bench_L_reference.err; echo "reference arm rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_L.json 2> gpurun_out/r02_bench_L.err; echo "L rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_L", "r02_bench_L_reference"):
    j = json.loads([l for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1])
    print(f, "value %.5g" % j["value"], "e2e %.5g" % j["e2e"]["value"], "roofline", {k: v for k, v in (j.get("roofline") or {}).items() if k in ("frac", "achieved", "copy_sustained_gbs_now")}, "parity", (j.get("parity") or {}).get("ok"), j.get("clocks"))
PY
