#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
tools/ubench/fma_peak > gpurun_out/r2_fma_peak.json 2>&1
cat gpurun_out/r2_fma_peak.json
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_base.json 2> gpurun_out/r2_base.err
HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_prof.so python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_base_prof.json 2>> gpurun_out/r2_base.err
cut -c1-400 gpurun_out/r2_base.json
tools/gpu_sanitize.sh
