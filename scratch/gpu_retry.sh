#!/bin/bash
# scratch/gpu_retry.sh <logfile> <timeout> [--gpus N] -- '<command>': retry while the pod answers "busy" (exit 3)
LOG=$1; shift
for i in $(seq 1 40); do
  bash scratch/gpu.sh "$@" > $LOG 2>&1; rc=$?
  if grep -q "status=transient" $LOG; then sleep 120; continue; fi
  exit $rc
done
