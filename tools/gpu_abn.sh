#!/bin/bash
# A/B/n of library variants: tools/gpu_abn.sh tag1 tag2 ...   (libhsrb.so vs libhsrb_<tag>.so), two runs each
mkdir -p gpurun_out
for rep in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/abn_base_$rep.json 2>> gpurun_out/abn.err
for TAG in "$@"; do
HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_$TAG.so python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/abn_${TAG}_$rep.json 2>> gpurun_out/abn.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/abn_*.json')):
    try:
        d=json.load(open(f))
        print(f, '%.2fM substeps/s'%(d['substeps_per_s']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'bad', d['bad_states'], 'succ', d['success_per_action'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/abn.err
