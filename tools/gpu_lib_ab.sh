#!/bin/bash
# usage: tools/gpu_lib_ab.sh libA.so libB.so ...   A/B of library builds on the headline workload (interleaved, two rounds)
for round in 1 2; do for L in "$@"; do
HSRB_LIB=$PWD/hsr_env_b200/csrc/$L python bench.py --steps 10 --warmup 5 --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$L', round(d['substeps_per_s']/1e6,2),'M substeps/s', round(d['ms_per_step'],2),'ms')"
done; done
for L in "$@"; do HSRB_LIB=$PWD/hsr_env_b200/csrc/$L python tools/light_env_rate.py 131072 2 | sed "s/^/$L /"; done
