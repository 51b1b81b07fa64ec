#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/p5_$name.json 2>> gpurun_out/p5.err
}
run base A=1
run mc6_wpb7 HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_mc6.so
run mc6_wpb4 HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_mc6.so HSRB_PUSH_WPB=4
run mc6_wpb3 HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_mc6.so HSRB_PUSH_WPB=3
run mc6_wpb2 HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_mc6.so HSRB_PUSH_WPB=2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/p5_*.json')):
    try:
        d=json.load(open(f)); c=d['config']
        print(f, c['threads_per_block'], c['grid'], c['resident_envs_per_sm'], '%.2fM'%(d['substeps_per_s']/1e6), 'bad', d['bad_states'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/p5.err
