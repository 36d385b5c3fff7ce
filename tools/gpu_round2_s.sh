#!/bin/bash
# round 2, GPU call S (1 GPU): new default long-row kernels (dit2 + prefetch, ditc<4|8>) -- row parity incl. general addressing, A/B bench
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "r2c_rows or decimation_in_time or c3_kernels or small_pow2_sweep" 2>&1 | tail -6 ) > gpurun_out/s_pytest.log
timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/s_bench_32768.json 2> gpurun_out/s_bench_32768.err
timeout 300 python bench.py --nx 4096 --ny 65536 --steps 10 $B > gpurun_out/s_bench_4096x65536.json 2> gpurun_out/s_bench_4096x65536.err
timeout 300 python bench.py --nx 2048 --ny 131072 --steps 10 $B > gpurun_out/s_bench_2048x131072.json 2> gpurun_out/s_bench_2048x131072.err
HPXFFT_B200_ROWS_GENERAL=1 timeout 300 python bench.py --nx 2048 --ny 131072 --steps 10 $B > gpurun_out/s_bench_2048x131072_general.json 2> gpurun_out/s_bench_2048x131072_general.err
ls -la gpurun_out | grep " s_"
