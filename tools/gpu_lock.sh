#!/bin/bash
# phase-lock experiment + small-batch latency probes
mkdir -p gpurun_out
: > gpurun_out/bench.err
python bench.py --steps 10 --warmup 3 --no-cpu --lanes 8 > gpurun_out/bench_push_l8.json 2>> gpurun_out/bench.err
HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_lock.so python bench.py --steps 10 --warmup 3 --no-cpu --lanes 8 > gpurun_out/bench_lock_l8.json 2>> gpurun_out/bench.err
HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_lock.so python bench.py --steps 10 --warmup 3 --no-cpu --lanes 16 > gpurun_out/bench_lock_l16.json 2>> gpurun_out/bench.err
for N in 592 1184 2368; do
  python bench.py --steps 10 --warmup 3 --no-cpu --lanes 8 --envs-per-gpu $N > gpurun_out/bench_push_n$N.json 2>> gpurun_out/bench.err
done
python bench.py --steps 10 --warmup 3 --no-cpu --lanes 32 --envs-per-gpu 1184 > gpurun_out/bench_push_l32_n1184.json 2>> gpurun_out/bench.err
python bench.py --steps 10 --warmup 3 --no-cpu --lanes 16 --envs-per-gpu 2368 > gpurun_out/bench_push_l16_n2368.json 2>> gpurun_out/bench.err
tail -n 4 gpurun_out/bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_push*.json')+glob.glob('gpurun_out/bench_lock*.json')):
    try:
        d=json.load(open(f))
        print(f, d['config']['envs_per_gpu'], round(d['value']), 'act/s', round(d['substeps_per_s']/1e6,2), 'Msub/s', round(d['ms_per_step'],2), 'ms', d['config'].get('lanes_per_env'), d['config'].get('threads_per_block'), d['config'].get('grid'), 'e2e', round(d['e2e']['value']))
    except Exception as e:
        print(f, 'ERR', e)
PY
