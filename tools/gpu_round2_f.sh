#!/bin/bash
# round 2, GPU call F (1 GPU): ny = 32768 row kernel with parked even bins + 256-bit stores
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "r2c_rows or c3_kernels or sweep" 2>&1 | tail -5 ) > gpurun_out/f_pytest.log
timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/f_bench_32768.json 2> gpurun_out/f_bench_32768.err
HPXFFT_B200_ROWS_LONG=1 timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/f_bench_32768_long1.json 2> gpurun_out/f_bench_32768_long1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rows_long2' -s 3 -c 1 -f -o gpurun_out/f_ncu_rows32768 \
    python bench.py --nx 32768 --ny 32768 --steps 1 --warmup 3 $B --no-parity > gpurun_out/f_ncu_rows32768.log 2>&1
ls -la gpurun_out | grep " f_"
