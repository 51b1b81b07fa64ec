#!/bin/bash
mkdir -p gpurun_out
HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_ck.so HSRB_WPE_LOCK=0 python tools/chain_clocks.py 148 10 2>&1 | tail -26
python tools/light_env_rate.py 4096 4
python tools/light_env_rate.py 131072 2
python bench.py --steps 10 --warmup 5 --no-cpu --no-configs > gpurun_out/ab2.json 2> gpurun_out/ab2.err; python -c "import json;d=json.load(open('gpurun_out/ab2.json'));print('bench', round(d['substeps_per_s']/1e6,2),'M substeps/s', round(d['ms_per_step'],2),'ms')"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
