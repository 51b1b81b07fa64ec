#!/bin/bash
# round-2 final pass: gpu tests, smoke, bench (both arms), ncu launch list + full capture of the wpe action kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
python bench.py --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu --no-configs > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hsrb_wpe_kernel -s 3 -c 1 -f -o gpurun_out/prof_wpe_r2f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.err gpurun_out/bench_reference.err
cut -c1-1200 gpurun_out/bench.json; echo; cut -c1-400 gpurun_out/bench_reference.json
