#!/bin/bash
# quick iteration: gpu parity tests of the fast kernel + bench (8 lanes, 4096 and 131072 envs) + optional ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
: > gpurun_out/bench.err
python bench.py --steps 10 --warmup 3 --no-cpu --lanes 8 > gpurun_out/bench_iter_l8.json 2>> gpurun_out/bench.err
python bench.py --steps 4 --warmup 3 --no-cpu --lanes 16 --envs-per-gpu 131072 > gpurun_out/bench_iter_131k_l16.json 2>> gpurun_out/bench.err
if [ "$1" = "ncu" ]; then
python bench.py --steps 1 --warmup 3 --no-cpu --lanes 8 > gpurun_out/plain_push.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hsrb_push_kernel -s 3 -c 1 -o gpurun_out/prof_push_l8 \
    python bench.py --steps 1 --warmup 3 --no-cpu --lanes 8 > gpurun_out/ncu_push.log 2>&1
fi
tail -n 4 gpurun_out/pytest_gpu.log gpurun_out/bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_iter*.json')):
    try:
        d=json.load(open(f))
        print(f, d['config']['envs_per_gpu'], round(d['value']), 'act/s', round(d['substeps_per_s']/1e6,2), 'Msub/s', round(d['ms_per_step'],2), 'ms', d['config'].get('lanes_per_env'), d['config'].get('threads_per_block'), d['config'].get('grid'), 'e2e', round(d['e2e']['value']))
    except Exception as e:
        print(f, 'ERR', e)
PY
