#!/bin/bash
mkdir -p gpurun_out
run() { # tag kernel envs lib
python bench.py --steps ${5:-5} --warmup 3 --no-cpu --kernel $2 --envs-per-gpu $3 > gpurun_out/$1.json 2> gpurun_out/$1.err; tail -2 gpurun_out/$1.err
python - <<PY
import json
d=json.load(open('gpurun_out/$1.json'))
print('$1', '%.2fM substeps/s'%(d['substeps_per_s']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'bad', d['bad_states'], d['config']['threads_per_block'], d['config']['grid'])
PY
}
run w3_wpe wpe 4096
run w3_wpe_c4 wpe 131072 x 2
HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_w16.so run w3_w16_c4 wpe 131072 x 2
HSRB_WPE_WPB=14 run w3_wpb14_c4 wpe 131072 x 2
HSRB_WPE_WPB=20 run w3_wpb20_c4 wpe 131072 x 2
