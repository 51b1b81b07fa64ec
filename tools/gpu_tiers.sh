#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 5 --no-cpu --no-configs > gpurun_out/tier_$tag.json 2> gpurun_out/tier_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/tier_$tag.json'));print('$tag', round(d['substeps_per_s']/1e6,2),'M substeps/s', round(d['ms_per_step'],2),'ms', d['bad_states'])" || tail -3 gpurun_out/tier_$tag.err; }
run uniform HSRB_WPE_TIERS=none
run t07x4 HSRB_WPE_TIERS=0.07:4
run t10x4 HSRB_WPE_TIERS=0.10:4
run t05x2_10x6 HSRB_WPE_TIERS=0.05:2,0.10:6
run t05x4_10x7 HSRB_WPE_TIERS=0.05:4,0.10:7
run t12x7 HSRB_WPE_TIERS=0.12:7
run t03x2 HSRB_WPE_TIERS=0.03:2
run t07x4_1team HSRB_WPE_TIERS=0.07:4 HSRB_WPE_TEAMS=1
run t07x4_4team HSRB_WPE_TIERS=0.07:2 HSRB_WPE_TEAMS=4
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
