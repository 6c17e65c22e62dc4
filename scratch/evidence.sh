#!/bin/bash
# round-2 evidence on one GPU: full GPU test suite, sanitizer logs, segment-mask / NaN-encoded bench lines
export PSMF_SPIN_TIMEOUT_MS=900000
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_gputests.log; tail -3 gpurun_out/r02_gputests.log
unset PSMF_SPIN_TIMEOUT_MS
python bench.py --mask segments --steps 10 > gpurun_out/r02_bench_L_segments.json 2> gpurun_out/r02_bench_L_segments.err; echo "segments rc=$?"
python bench.py --mask-encoding nan --steps 10 > gpurun_out/r02_bench_L_nan.json 2> gpurun_out/r02_bench_L_nan.err; echo "nan rc=$?"
python bench.py --dtype f32 --steps 10 > gpurun_out/r02_bench_L_f32.json 2> gpurun_out/r02_bench_L_f32.err; echo "f32 rc=$?"
export PSMF_SPIN_TIMEOUT_MS=900000
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python scratch/sanitize.py stream direct batch stream_resident > gpurun_out/r02_sanitizer_$tool.log 2>&1; echo "$tool rc=$? $(grep -c 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r02_sanitizer_$tool.log)"; tail -3 gpurun_out/r02_sanitizer_$tool.log
done
