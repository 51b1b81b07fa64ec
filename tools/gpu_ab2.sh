#!/bin/bash
# A/B of library variants: tools/gpu_ab2.sh <tag> [bench args]  (libhsrb.so vs libhsrb_<tag>.so)
mkdir -p gpurun_out
TAG=$1; shift
python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/ab_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/ab_bench.json 2> gpurun_out/ab.err
HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_$TAG.so python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/ab_bench_$TAG.json 2>> gpurun_out/ab.err
tail -3 gpurun_out/ab_pytest.log
python - <<PY
import json
for f in ['ab_bench','ab_bench_$TAG']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, '%.2fM substeps/s'%(d['substeps_per_s']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'e2e', d['e2e']['value'], 'bad', d['bad_states'], 'succ', d['success_per_action'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/ab.err
