#!/bin/bash
# A/B of library variants of the phase-locked wpe kernel: tools/gpu_ab_lock.sh tag1 tag2 ...  ("base" = libhsrb.so)
mkdir -p gpurun_out
for rep in 1 2; do
for TAG in "$@"; do
LIB=$PWD/hsr_env_b200/csrc/libhsrb_$TAG.so; [ $TAG == base ] && LIB=$PWD/hsr_env_b200/csrc/libhsrb.so
HSRB_WPE_LOCK=1 HSRB_LIB=$LIB python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --kernel wpe > gpurun_out/abl_${TAG}_$rep.json 2>> gpurun_out/abl.err
HSRB_WPE_LOCK=1 HSRB_LIB=$LIB python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --kernel wpe --envs-per-gpu 131072 > gpurun_out/abl_${TAG}_c4_$rep.json 2>> gpurun_out/abl.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/abl_*.json')):
    try:
        d=json.load(open(f))
        print(f, '%.2fM substeps/s'%(d['substeps_per_s']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'bad', d['bad_states'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/abl.err
