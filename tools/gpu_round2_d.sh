#!/bin/bash
# round 2, GPU call D (1 GPU): v2 row kernel (ny = 16384), long-row kernel with L2 eviction hints
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/d_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/d_bench_16384.json 2> gpurun_out/d_bench_16384.err
HPXFFT_B200_ROWS_V1=1 timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/d_bench_16384_rowsv1.json 2> gpurun_out/d_bench_16384_rowsv1.err
timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/d_bench_32768.json 2> gpurun_out/d_bench_32768.err
timeout 300 python bench.py --nx 4096 --ny 65536 --steps 10 $B > gpurun_out/d_bench_4096x65536.json 2> gpurun_out/d_bench_4096x65536.err
timeout 300 python bench.py --nx 2048 --ny 131072 --steps 10 $B > gpurun_out/d_bench_2048x131072.json 2> gpurun_out/d_bench_2048x131072.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rows_long_kernel|rows_r2c_v2' -s 3 -c 1 -f -o gpurun_out/d_ncu_rows32768 \
    python bench.py --nx 32768 --ny 32768 --steps 1 --warmup 3 $B --no-parity > gpurun_out/d_ncu_rows32768.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rows_r2c_v2' -s 3 -c 1 -f -o gpurun_out/d_ncu_rows16384 \
    python bench.py --steps 1 --warmup 3 $B --no-parity > gpurun_out/d_ncu_rows16384.log 2>&1
ls -la gpurun_out | grep " d_"
