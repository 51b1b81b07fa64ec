#!/bin/bash
# general-kernel configs (C3 gripper, C5 clutter): old free-running one-warp blocks vs the phase-locked kernel
mkdir -p gpurun_out
for MODE in 0 1; do
HSRB_GENERAL_LOCK=$MODE python bench.py --steps 3 --warmup 3 --no-cpu --only-configs c3,c5 > gpurun_out/gen_$MODE.json 2> gpurun_out/gen_$MODE.err; tail -2 gpurun_out/gen_$MODE.err
python - <<PY
import json
d=json.load(open('gpurun_out/gen_$MODE.json'))
for k,v in d['configs'].items():
    if k.startswith('c4'): continue
    print('lock=$MODE', k, '%.2fM substeps/s'%(v['substeps_per_s']/1e6), 'bad', v['bad_states'], 'lanes', v['lanes_per_env'], 'envs/SM', v['resident_envs_per_sm'], 'tpb', v['threads_per_block'], 'contacts', round(v['mean_contacts_per_substep'],2), 'iters', round(v['mean_newton_iters_per_substep'],2))
PY
done
python -m pytest tests -q -m gpu -k "golden or one_substep or goal_list or cupboard" 2>&1 | tail -4
