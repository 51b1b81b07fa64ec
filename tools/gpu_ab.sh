#!/bin/bash
# quick A/B: gpu parity tests + bench (default config), optional extra args
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/ab_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/ab_bench.json 2> gpurun_out/ab.err
HSRB_OPTS=${ABOPTS:-1} python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/ab_bench_opts1.json 2>> gpurun_out/ab.err
tail -3 gpurun_out/ab_pytest.log
python - <<'PY'
import json
for f in ['ab_bench','ab_bench_opts1']:
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, '%.2fM substeps/s'%(d['substeps_per_s']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'e2e', d['e2e']['value'], 'flops/sub', d['fp32']['mean_algorithmic_flops_per_substep'], 'bad', d['bad_states'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/ab.err
