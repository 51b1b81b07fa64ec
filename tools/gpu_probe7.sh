#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/p7_*
R1=$PWD/hsr_env_b200/csrc/libhsrb_r1.so
HSRB_LIB=$R1 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { name=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/p7_$name.json 2>> gpurun_out/p7.err; }
run base A=1
run r1 HSRB_LIB=$R1
run r1_wpb4 HSRB_LIB=$R1 HSRB_PUSH_WPB=4
run r1_wpb8 HSRB_LIB=$R1 HSRB_PUSH_WPB=8
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/p7_*.json')):
    try:
        d=json.load(open(f)); c=d['config']
        print(f, c['smem_per_env'], c['threads_per_block'], c['grid'], c['resident_envs_per_sm'], '%.2fM'%(d['substeps_per_s']/1e6), 'bad', d['bad_states'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/p7.err
