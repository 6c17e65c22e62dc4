#!/bin/bash
# ncu captures for profiles/: launch list of the bench command + full-set capture of the dominant kernel
# (streaming regime d = 1M, and the shared-memory-resident regime of an 8-GPU shard, d = 125k)
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv \
    python bench.py --steps 2 --warmup 1 --T 2000 --no-e2e --no-cpu > gpurun_out/r01b_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:psmf_ -s 2 -c 1 -f -o gpurun_out/r01b_stream_full \
    python bench.py --T 400 --window 100 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r01b_full.log 2>&1
ncu -i gpurun_out/r01b_stream_full.ncu-rep --page raw --csv > gpurun_out/r01b_stream_full_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:psmf_ -s 2 -c 1 -f -o gpurun_out/r01b_resident_full \
    python bench.py --d 125024 --T 800 --window 200 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r01b_resident.log 2>&1
ncu -i gpurun_out/r01b_resident_full.ncu-rep --page raw --csv > gpurun_out/r01b_resident_full_raw.csv 2>/dev/null
ls -la gpurun_out/
