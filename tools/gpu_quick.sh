#!/bin/bash
# gpu tests + smoke + bench (fast and general kernels), no profiler
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_fast.json 2> gpurun_out/bench.err
python bench.py --steps 10 --warmup 3 --no-cpu --envs-per-gpu 131072 > gpurun_out/bench_fast_131k.json 2>> gpurun_out/bench.err
python bench.py --steps 3 --warmup 3 --no-cpu --kernel general > gpurun_out/bench_general.json 2>> gpurun_out/bench.err
tail -n 8 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.err
cut -c1-900 gpurun_out/bench_fast.json; echo; cut -c1-400 gpurun_out/bench_fast_131k.json; echo; cut -c1-400 gpurun_out/bench_general.json
