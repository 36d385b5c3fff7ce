#!/bin/bash
# round 2, GPU call C (1 GPU): long-row kernel with batched assembly, column pre-stage (split 2) on/off, fence cost
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "r2c_rows or c2c_cols or sweep or c3_kernels" 2>&1 | tail -30 ) > gpurun_out/c_pytest.log
timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/c_bench_32768.json 2> gpurun_out/c_bench_32768.err
HPXFFT_B200_COLSPLIT=0 timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/c_bench_32768_nosplit.json 2> gpurun_out/c_bench_32768_nosplit.err
HPXFFT_B200_LIB=$PWD/hpx-fft_b200/libdiag_nofence.so timeout 300 python bench.py --steps 10 $B --no-parity > gpurun_out/c_diag_nofence_16384.json 2> gpurun_out/c_diag_nofence_16384.err
for bps in 3 4; do HPXFFT_B200_FUSED_BPS=$bps timeout 300 python bench.py --steps 10 $B > gpurun_out/c_bench_16384_bps$bps.json 2> gpurun_out/c_bench_16384_bps$bps.err; done
for lag in 3 4; do HPXFFT_B200_LAG=$lag timeout 300 python bench.py --steps 10 $B > gpurun_out/c_bench_16384_lag$lag.json 2> gpurun_out/c_bench_16384_lag$lag.err; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rows_long_kernel|cols_fused_kernel' -s 4 -c 2 -f -o gpurun_out/c_ncu_32768 \
    python bench.py --nx 32768 --ny 32768 --steps 1 --warmup 3 $B --no-parity > gpurun_out/c_ncu_32768.log 2>&1
ls -la gpurun_out | grep " c_"
