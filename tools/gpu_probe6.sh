#!/bin/bash
mkdir -p gpurun_out
for pad in 0 1 2 3 4 6 8 10; do
  HSRB_PUSH_PAD=$pad python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/p6_pad$pad.json 2>> gpurun_out/p6.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/p6_*.json')):
    try:
        d=json.load(open(f)); c=d['config']
        print(f, c['smem_per_env'], c['threads_per_block'], c['grid'], '%.2fM'%(d['substeps_per_s']/1e6), 'bad', d['bad_states'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/p6.err
