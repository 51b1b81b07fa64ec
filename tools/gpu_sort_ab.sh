#!/bin/bash
# A/B of the work-sorted launch order of the wpe kernel (HSRB_WPE_SORT) x teams per block
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 5 --no-cpu --no-configs > gpurun_out/sort_$tag.json 2> gpurun_out/sort_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/sort_$tag.json'));print('$tag', round(d['substeps_per_s']/1e6,2),'M substeps/s', round(d['ms_per_step'],2),'ms')"; }
run s0_t2 HSRB_WPE_SORT=0 HSRB_WPE_TEAMS=2
run s1_t2 HSRB_WPE_SORT=1 HSRB_WPE_TEAMS=2
run s1_t4 HSRB_WPE_SORT=1 HSRB_WPE_TEAMS=4
run s0_t4 HSRB_WPE_SORT=0 HSRB_WPE_TEAMS=4
run s1_t1 HSRB_WPE_SORT=1 HSRB_WPE_TEAMS=1
run s1_free HSRB_WPE_SORT=1 HSRB_WPE_LOCK=0
runc4() { tag=$1; shift; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu --only-configs c4 --c4-envs 131072 > gpurun_out/sort_c4_$tag.json 2> gpurun_out/sort_c4_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/sort_c4_$tag.json'))['configs']['c4_1m_envs_sharded'];print('c4 $tag', round(d['substeps_per_s']/1e6,2),'M substeps/s')"; }
runc4 s0_t2 HSRB_WPE_SORT=0 HSRB_WPE_TEAMS=2
runc4 s1_t2 HSRB_WPE_SORT=1 HSRB_WPE_TEAMS=2
runc4 s1_t4 HSRB_WPE_SORT=1 HSRB_WPE_TEAMS=4
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
