#!/bin/bash
# A/B of the general lock kernel: phase barriers (opts bit 1 off = on) and pass-locked Newton (bit 2)
mkdir -p gpurun_out
for o in 0x6 0x4 0x2 0x0; do
HSRB_OPTS=$o python bench.py --steps 3 --warmup 3 --no-cpu --only-configs c3,c5 > gpurun_out/gl_$o.json 2> gpurun_out/gl_$o.err
python -c "
import json;d=json.load(open('gpurun_out/gl_$o.json'))['configs']
for k in ('c3_arm_gripper','c5_clutter'): print('opts=$o', k, round(d[k]['substeps_per_s']/1e6,3), 'M substeps/s')"
done
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
