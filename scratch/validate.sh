#!/bin/bash
# full validation of HEAD on one GPU: GPU test suite, smoke(), the driver's default bench command, workload B, resident shard
mkdir -p gpurun_out
export PSMF_SPIN_TIMEOUT_MS=900000
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02c_gputests.log; tail -3 gpurun_out/r02c_gputests.log
unset PSMF_SPIN_TIMEOUT_MS
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02c_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02c_bench_L.json 2> gpurun_out/r02c_bench_L.err; echo "L rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --mask-encoding nan > gpurun_out/r02c_bench_L_nan.json 2> gpurun_out/r02c_bench_L_nan.err; echo "Lnan rc=$?"
timeout 600 python bench.py --workload B --series 512 --steps 10 --warmup 3 > gpurun_out/r02c_bench_B512.json 2> gpurun_out/r02c_bench_B512.err; echo "B rc=$?"
bash scratch/quick3.sh 2>&1 | tee gpurun_out/r02c_quick3.log
python - <<'PY'
import json
for f in ("r02c_bench_L", "r02c_bench_L_nan", "r02c_bench_B512"):
    try:
        j = json.loads([l for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1])
        print(f, "value %.4g" % j["value"], "frac %.3f" % j["roofline"]["frac"], "e2e %.4g" % j["e2e"]["value"], "parity", j["parity"]["ok"], "cpu", j["cpu_baseline"]["value"], j["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
