#!/bin/bash
# round 2, GPU call Q (1 GPU): bulk L2 prefetch of the next row (HPXFFT_B200_ROWS_PF=1) in the ny = 16384 and ny = 32768 row kernels
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( HPXFFT_B200_ROWS_PF=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decimation_in_time or r2c_rows or c2_16384" 2>&1 | tail -4 ) > gpurun_out/q_pytest.log
for pf in 0 1; do
  HPXFFT_B200_ROWS_PF=$pf timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/q_bench_16384_pf$pf.json 2> gpurun_out/q_bench_16384_pf$pf.err
  HPXFFT_B200_ROWS_PF=$pf HPXFFT_B200_ROWS_LONG=3 timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/q_bench_32768_pf$pf.json 2> gpurun_out/q_bench_32768_pf$pf.err
done
ls -la gpurun_out | grep " q_"
