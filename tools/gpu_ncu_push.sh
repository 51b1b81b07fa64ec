#!/bin/bash
# one full ncu capture of the fast-path action kernel (after the same command ran clean without ncu)
mkdir -p gpurun_out
L=${1:-8}
python bench.py --steps 1 --warmup 3 --no-cpu --lanes $L > gpurun_out/plain_push.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hsrb_push_kernel -s 3 -c 1 -o gpurun_out/prof_push_l$L \
    python bench.py --steps 1 --warmup 3 --no-cpu --lanes $L > gpurun_out/ncu_push.log 2>&1
tail -3 gpurun_out/plain_push.log gpurun_out/ncu_push.log | cut -c1-300
