#!/bin/bash
# full validation of HEAD on one GPU: GPU test suite, smoke(), the driver's default bench command, the other bench lines
mkdir -p gpurun_out
export PSMF_SPIN_TIMEOUT_MS=900000
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_gputests.log; tail -3 gpurun_out/r02_gputests.log
unset PSMF_SPIN_TIMEOUT_MS
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_L.json 2> gpurun_out/r02_bench_L.err; echo "L rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --impl reference > gpurun_out/r02_bench_L_reference.json 2> gpurun_out/r02_bench_L_reference.err; echo "L reference arm rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --mask-encoding bytes > gpurun_out/r02_bench_L_bytes.json 2> gpurun_out/r02_bench_L_bytes.err; echo "Lbytes rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --mask segments > gpurun_out/r02_bench_L_segments.json 2> gpurun_out/r02_bench_L_segments.err; echo "Lseg rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --dtype f32 > gpurun_out/r02_bench_L_f32.json 2> gpurun_out/r02_bench_L_f32.err; echo "Lf32 rc=$?"
timeout 600 python bench.py --workload B --series 512 --steps 10 --warmup 3 > gpurun_out/r02_bench_B512.json 2> gpurun_out/r02_bench_B512.err; echo "B rc=$?"
bash scratch/quick3.sh 2>&1 | tee gpurun_out/r02_quick3.log
python scratch/synthetic_timing.py b200 20 2>&1 | tail -2 | tee gpurun_out/r02_synthetic_timing.log
python scratch/impute_timing.py b200 2>&1 | tail -6 | tee gpurun_out/r02_impute_timing.log
python - <<'PY'
import json
for f in ("r02_bench_L", "r02_bench_L_reference", "r02_bench_L_bytes", "r02_bench_L_segments", "r02_bench_L_f32", "r02_bench_B512"):
    try:
        j = json.loads([l for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1])
        print(f, "value %.4g" % j["value"], "frac", (j.get("roofline") or {}).get("frac"), "e2e %.4g" % j["e2e"]["value"], "parity", (j.get("parity") or {}).get("ok"),
              "cpu", (j.get("cpu_baseline") or {}).get("value"), j.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
