#!/bin/bash
mkdir -p gpurun_out
HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_prof.so python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/p1_phases.json 2> gpurun_out/p1.err
HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_prof.so python bench.py --steps 5 --warmup 3 --no-cpu --action-scale 0 > gpurun_out/p1_phases_zero.json 2>> gpurun_out/p1.err
python bench.py --steps 10 --warmup 3 --no-cpu --action-scale 0 > gpurun_out/p1_zero.json 2>> gpurun_out/p1.err
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/p1_std.json 2>> gpurun_out/p1.err
python - <<'PY'
import json
for f in ['p1_phases','p1_phases_zero','p1_zero','p1_std']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['substeps_per_s']/1e6, d['mean_substeps_per_action'], d.get('phase_share'), d.get('phase_cycles_per_substep_lane0'), d['fp32']['mean_algorithmic_flops_per_substep'])
PY
tail -3 gpurun_out/p1.err
