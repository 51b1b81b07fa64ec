#!/bin/bash
# one full ncu capture of the warp-per-environment action kernel (after the same command ran clean without ncu)
# usage: tools/gpu_ncu_wpe.sh tag [env assignments...]
mkdir -p gpurun_out
TAG=${1:-wpe}; shift
env "$@" python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --kernel wpe > gpurun_out/plain_$TAG.log 2>&1 &&
env "$@" ncu --set full --clock-control none --import-source on -k regex:hsrb_wpe_kernel -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --kernel wpe > gpurun_out/ncu_$TAG.log 2>&1
tail -n 2 gpurun_out/plain_$TAG.log | cut -c1-300
