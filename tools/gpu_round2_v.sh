#!/bin/bash
# round 2, GPU call V (1 GPU): the whole GPU suite on the final library, smoke(), the default bench line, 32768^2
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/v_pytest_all.log
( timeout 300 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2 ) > gpurun_out/v_smoke.log
timeout 600 python bench.py > gpurun_out/v_bench_default.json 2> gpurun_out/v_bench_default.err
timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/v_bench_32768.json 2> gpurun_out/v_bench_32768.err
ls -la gpurun_out | grep " v_"
