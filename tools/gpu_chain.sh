#!/bin/bash
# chain-section latencies of the wpe kernel (free-running, one warp per SM and the full 4096-env batch)
mkdir -p gpurun_out
export HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_ck.so HSRB_WPE_LOCK=0
python tools/chain_clocks.py 148 10 > gpurun_out/chain_148.txt 2>&1
python tools/chain_clocks.py 4096 5 > gpurun_out/chain_4096.txt 2>&1
cat gpurun_out/chain_148.txt; tail -30 gpurun_out/chain_4096.txt
