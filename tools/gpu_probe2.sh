#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu --action-scale 0 > gpurun_out/p2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hsrb_push_kernel -s 3 -c 1 -o gpurun_out/prof_push_zero \
    python bench.py --steps 1 --warmup 3 --no-cpu --action-scale 0 > gpurun_out/p2_ncu.log 2>&1
tail -2 gpurun_out/p2_ncu.log | cut -c1-300
