#!/bin/bash
# GPU tests + smoke + bench, then the phase-clock build (libhsrb_prof.so) at several lanes-per-env settings.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_auto.json 2> gpurun_out/bench.err
for L in 4 8 16 32; do
  HSRB_LIB=$PWD/hsr_env_b200/csrc/libhsrb_prof.so python bench.py --steps 3 --warmup 3 --no-cpu --lanes $L > gpurun_out/phases_l$L.json 2>> gpurun_out/bench.err
done
tail -n 5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
cat gpurun_out/bench_auto.json
