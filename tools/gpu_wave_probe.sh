#!/bin/bash
for n in 4096 8192 16384 32768; do python tools/light_env_rate.py $n 3 2>&1 | grep parked; done
echo "wpb=14:"; HSRB_WPE_WPB=14 python tools/light_env_rate.py 4096 3 2>&1 | grep parked
echo "wpb=7:"; HSRB_WPE_WPB=7 python tools/light_env_rate.py 4096 3 2>&1 | grep parked
echo "sort off:"; HSRB_WPE_SORT=0 python tools/light_env_rate.py 4096 3 2>&1 | grep parked
echo "teams 1:"; HSRB_WPE_TEAMS=1 python tools/light_env_rate.py 4096 3 2>&1 | grep parked
echo "free:"; HSRB_WPE_LOCK=0 python tools/light_env_rate.py 4096 3 2>&1 | grep parked
