#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for round in 1 2; do for o in 0x20020 0x20; do
HSRB_OPTS=$o python bench.py --steps 10 --warmup 5 --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('opts=$o', round(d['substeps_per_s']/1e6,2),'M substeps/s', round(d['ms_per_step'],2),'ms')"
done; done
for o in 0x20020 0x20; do HSRB_OPTS=$o python tools/light_env_rate.py 131072 2 | sed "s/^/opts=$o /"; HSRB_OPTS=$o python tools/light_env_rate.py 4096 4 | sed "s/^/opts=$o /"; done
