#!/bin/bash
mkdir -p gpurun_out
run() { # tag kernel envs steps
python bench.py --steps ${4:-5} --warmup 3 --no-cpu --no-configs --kernel $2 --envs-per-gpu $3 > gpurun_out/$1.json 2> gpurun_out/$1.err; tail -2 gpurun_out/$1.err
python - <<PY
import json
d=json.load(open('gpurun_out/$1.json'))
print('$1', '%.2fM substeps/s'%(d['substeps_per_s']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'bad', d['bad_states'], d['config']['threads_per_block'], d['config']['grid'])
PY
}
run w4_fast fast 4096
run w4_wpe wpe 4096
HSRB_WPE_LOCK=1 run w4_wpelock wpe 4096
run w4_fast_c4 fast 131072 2
run w4_wpe_c4 wpe 131072 2
HSRB_WPE_LOCK=1 run w4_wpelock_c4 wpe 131072 2
