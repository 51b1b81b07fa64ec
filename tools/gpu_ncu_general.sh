#!/bin/bash
# one full ncu capture of the phase-locked general kernel on the C3 (gripper) workload
mkdir -p gpurun_out
TAG=${1:-c3lock}
python bench.py --steps 3 --warmup 3 --no-cpu --only-configs c3 > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hsrb_step_lock_kernel -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu --only-configs c3 > gpurun_out/ncu_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_$TAG.log | cut -c1-200
