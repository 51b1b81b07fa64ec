#!/bin/bash
mkdir -p gpurun_out
for OPT in 0 0x20 0x40; do for rep in 1 2; do
HSRB_OPTS=$OPT HSRB_WPE_LOCK=1 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --kernel wpe > gpurun_out/tm_${OPT}_$rep.json 2>> gpurun_out/tm.err
done
HSRB_OPTS=$OPT HSRB_WPE_LOCK=1 python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --kernel wpe --envs-per-gpu 131072 > gpurun_out/tm_${OPT}_c4.json 2>> gpurun_out/tm.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/tm_*.json')):
    try:
        d=json.load(open(f))
        print(f, '%.2fM substeps/s'%(d['substeps_per_s']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'bad', d['bad_states'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/tm.err
