#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for s in 0 1; do
HSRB_WPE_SORT=$s python bench.py --steps 3 --warmup 3 --no-cpu --only-configs c3,c5 > gpurun_out/gs_$s.json 2> gpurun_out/gs_$s.err
python -c "
import json;d=json.load(open('gpurun_out/gs_$s.json'))['configs']
for k,v in d.items(): print('sort=$s', k, round(v['substeps_per_s']/1e6,3), 'M substeps/s')"
done
