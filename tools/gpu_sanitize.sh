#!/bin/bash
# compute-sanitizer memcheck + racecheck of every kernel (small batches, few substeps): logs -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for CASE in fast general c3 c5; do
  for TOOL in memcheck racecheck; do
    timeout 900 $CS --tool $TOOL --print-limit 20 python tools/sanitize_case.py $CASE 4 > gpurun_out/sanitize_${CASE}_${TOOL}.log 2>&1
    echo "exit $?" >> gpurun_out/sanitize_${CASE}_${TOOL}.log
    tail -n 4 gpurun_out/sanitize_${CASE}_${TOOL}.log
  done
done
