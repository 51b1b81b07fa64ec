#!/bin/bash
# A/B of the two fast kernels on the bench workload + one full ncu capture of the warp-per-environment kernel
mkdir -p gpurun_out
TAG=${1:-w2}
python -m pytest tests/test_gpu_parity.py -x -q -k "wpe or independent" 2>&1 | tail -3
for K in fast wpe; do
python bench.py --steps 5 --warmup 3 --no-cpu --kernel $K > gpurun_out/${TAG}_$K.json 2> gpurun_out/${TAG}_$K.err; tail -2 gpurun_out/${TAG}_$K.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_$K.json'))
print('$K', '%.2fM substeps/s'%(d['substeps_per_s']/1e6), 'ms/step %.2f'%d['ms_per_step'], 'bad', d['bad_states'], 'succ', d['success_per_action'], d['config']['threads_per_block'], d['config']['grid'], 'e2e', d['e2e']['value'])
PY
done
if [ "$2" == "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:hsrb_wpe_kernel -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu --kernel wpe > gpurun_out/ncu_$TAG.log 2>&1
fi
