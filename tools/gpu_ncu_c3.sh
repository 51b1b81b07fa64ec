#!/bin/bash
# one full ncu capture of the general kernel on C3 (after the same command ran clean without ncu)
mkdir -p gpurun_out
python tools/bench_configs.py c3 > gpurun_out/c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hsrb_step_kernel -s 3 -c 1 -o gpurun_out/prof_c3_general \
    python tools/bench_configs.py c3 > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_ncu.log | cut -c1-200
