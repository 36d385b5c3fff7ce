#!/bin/bash
# round 2, GPU call M (1 GPU): Bluestein lengths, selectable row-kernel variants, then the whole GPU suite once more on the final library
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_z_bluestein_gpu.py -m gpu -q 2>&1 | tail -25 ) > gpurun_out/m_pytest_bluestein.log
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > gpurun_out/m_pytest_all.log
timeout 200 python bench.py --nx 8209 --ny 16418 --steps 5 --no-e2e --no-cpu-baseline > gpurun_out/m_bench_8209x16418.json 2> gpurun_out/m_bench_8209x16418.err
ls -la gpurun_out | grep " m_"
