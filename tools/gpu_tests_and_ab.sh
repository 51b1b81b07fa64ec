#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
for K in fast wpe; do
python bench.py --steps 5 --warmup 3 --no-cpu --kernel $K > gpurun_out/ab_$K.json 2> gpurun_out/ab_$K.err; tail -2 gpurun_out/ab_$K.err
python - <<PY
import json
d=json.load(open('gpurun_out/ab_$K.json'))
print('$K', '%.2fM substeps/s'%(d['substeps_per_s']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'bad', d['bad_states'], 'succ', d['success_per_action'], d['config']['threads_per_block'], d['config']['grid'])
PY
done
