#!/bin/bash
mkdir -p gpurun_out
for cfg in "32 1184" "16 2368" "16 4096" "8 1184" "8 4096"; do
  set -- $cfg
  python bench.py --steps 6 --warmup 3 --no-cpu --lanes $1 --envs-per-gpu $2 > gpurun_out/p3_l$1_n$2.json 2>> gpurun_out/p3.err
  HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_prof.so python bench.py --steps 4 --warmup 3 --no-cpu --lanes $1 --envs-per-gpu $2 > gpurun_out/p3_ph_l$1_n$2.json 2>> gpurun_out/p3.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/p3_*.json')):
    d=json.load(open(f)); c=d['config']
    print(f, c['lanes_per_env'], c['envs_per_gpu'], c['threads_per_block'], c['grid'], '%.2fM'%(d['substeps_per_s']/1e6), 'ms/step %.1f'%d['ms_per_step'], d.get('phase_share'), d.get('phase_cycles_per_substep_lane0'))
PY
tail -3 gpurun_out/p3.err
