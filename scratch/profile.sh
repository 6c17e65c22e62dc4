#!/bin/bash
# round-2 ncu captures for profiles/ (one GPU): launch list of the driver's bench command (warm-up + timed region only),
# full-set captures of the pipelined kernel at the shard sizes of 1 / 2 / 4 / 8 GPUs (per-launch DRAM traffic for
# roofline.traffic) and of the resident batch kernel.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_B.csv \
    python bench.py --workload B --series 512 --T 1500 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_launches_B.log 2>&1; echo "launch list B rc=$?"
for rows in 1000000 500000 250000 125024; do
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:psmf_stream -s 2 -c 1 -f \
      -o gpurun_out/r02_stream_full_$rows python bench.py --rows $rows --T 400 --window 100 --steps 2 --warmup 3 --no-e2e --no-cpu \
      --parity-steps 0 > gpurun_out/r02_full_$rows.log 2>&1; echo "full $rows rc=$?"
  ncu -i gpurun_out/r02_stream_full_$rows.ncu-rep --page raw --csv > gpurun_out/r02_stream_full_${rows}_raw.csv 2>/dev/null
done
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:psmf_batch -s 2 -c 1 -f \
    -o gpurun_out/r02_batch_full python bench.py --workload B --series 512 --T 500 --window 250 --steps 2 --warmup 3 --no-e2e --no-cpu \
    --parity-steps 0 > gpurun_out/r02_batch_full.log 2>&1; echo "batch full rc=$?"
ncu -i gpurun_out/r02_batch_full.ncu-rep --page raw --csv > gpurun_out/r02_batch_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_stream_full_500000.ncu-rep gpurun_out/r02_stream_full_250000.ncu-rep     # keep the 64 MiB return budget
ls -la gpurun_out/ | tail -20
