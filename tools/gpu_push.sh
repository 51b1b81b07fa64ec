#!/bin/bash
# gpu tests + smoke + bench of the fast (push) kernel at 8/16/32 lanes per environment, phase-clock build at the end
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
: > gpurun_out/bench.err
for L in 8 16 32; do
  python bench.py --steps 10 --warmup 3 --no-cpu --lanes $L > gpurun_out/bench_push_l$L.json 2>> gpurun_out/bench.err
done
python bench.py --steps 5 --warmup 3 --no-cpu --envs-per-gpu 131072 --lanes 8 > gpurun_out/bench_push_131k_l8.json 2>> gpurun_out/bench.err
if [ -f hsr_env_b200/csrc/libhsrb_prof.so ]; then
  for L in 8 16; do
    HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_prof.so python bench.py --steps 3 --warmup 3 --no-cpu --lanes $L > gpurun_out/phases_push_l$L.json 2>> gpurun_out/bench.err
  done
fi
tail -n 6 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_push*.json')+glob.glob('gpurun_out/phases_push*.json')):
    try:
        d=json.load(open(f))
        print(f, round(d['value']), 'act/s', round(d['substeps_per_s']/1e6,2), 'Msub/s', d['ms_per_step'], 'ms', d['config'].get('lanes_per_env'), d['config'].get('threads_per_block'), d['config'].get('grid'), 'e2e', round(d['e2e']['value']), d.get('phase_share'))
    except Exception as e:
        print(f, 'ERR', e)
PY
