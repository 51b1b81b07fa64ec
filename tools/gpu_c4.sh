#!/bin/bash
# C4 share (131072 envs on one GPU) on both fast kernels
mkdir -p gpurun_out
for K in "$@"; do
python bench.py --steps 2 --warmup 3 --no-cpu --kernel $K --envs-per-gpu 131072 > gpurun_out/c4_$K.json 2> gpurun_out/c4_$K.err; tail -2 gpurun_out/c4_$K.err
python - <<PY
import json
d=json.load(open('gpurun_out/c4_$K.json'))
print('$K', '%.2fM substeps/s'%(d['substeps_per_s']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'bad', d['bad_states'], d['config']['threads_per_block'], d['config']['grid'])
PY
done
