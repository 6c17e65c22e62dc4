#!/bin/bash
# round-2 evidence on one GPU: validation (tests, smoke, bench lines), ncu captures, sanitizer logs
bash scratch/validate.sh
bash scratch/profile.sh
export PSMF_SPIN_TIMEOUT_MS=900000
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python scratch/sanitize.py stream direct batch stream_resident > gpurun_out/r02_sanitizer_$tool.log 2>&1; echo "$tool rc=$? $(grep -c 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r02_sanitizer_$tool.log)"; tail -3 gpurun_out/r02_sanitizer_$tool.log
done
