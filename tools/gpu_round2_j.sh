#!/bin/bash
# round 2, GPU call J (1 GPU): L2 prefetch of the next row / next level-A tile, on vs off
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "r2c_rows or c2c_cols or c2_16384 or c3_kernels" 2>&1 | tail -4 ) > gpurun_out/j_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/j_bench_16384.json 2> gpurun_out/j_bench_16384.err
HPXFFT_B200_COLPF=0 timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/j_bench_16384_nocolpf.json 2> gpurun_out/j_bench_16384_nocolpf.err
timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/j_bench_32768.json 2> gpurun_out/j_bench_32768.err
HPXFFT_B200_COLPF=0 timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/j_bench_32768_nocolpf.json 2> gpurun_out/j_bench_32768_nocolpf.err
timeout 300 python bench.py --nx 8192 --ny 8192 --steps 20 $B > gpurun_out/j_bench_8192.json 2> gpurun_out/j_bench_8192.err
ls -la gpurun_out | grep " j_"
