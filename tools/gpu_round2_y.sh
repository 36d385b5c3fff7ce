#!/bin/bash
# round 2, GPU call Y (1 GPU): last sanity run of the final defaults
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "r2c_rows or c2_16384 or golden or anchor or end_to_end or c3_kernels" 2>&1 | tail -3 ) > gpurun_out/y_pytest.log
( timeout 300 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2 ) > gpurun_out/y_smoke.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/y_bench_16384.json 2> gpurun_out/y_bench_16384.err
cat gpurun_out/y_pytest.log gpurun_out/y_smoke.log
