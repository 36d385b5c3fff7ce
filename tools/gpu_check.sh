#!/bin/bash
# usage (under gpurun): bash tools/gpu_check.sh [tests|bench|ncu|all]
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
if [[ $what == tests || $what == all ]]; then
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -30 gpurun_out/pytest_gpu.log
fi
if [[ $what == bench || $what == all ]]; then
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [[ $what == ncu || $what == all ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"; tail -3 gpurun_out/ncu_bench.log
fi
