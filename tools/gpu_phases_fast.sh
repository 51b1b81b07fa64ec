#!/bin/bash
mkdir -p gpurun_out
HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_prof.so python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/phases_fast.json 2> gpurun_out/bench.err
python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hsrb_fast_kernel -s 3 -c 1 -o gpurun_out/prof_fast \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/phases_fast.json'))
print(d['value'], d['substeps_per_s'], d.get('phase_share'), d.get('phase_cycles_per_substep_lane0'))
PY
tail -3 gpurun_out/bench.err gpurun_out/ncu_full.log
