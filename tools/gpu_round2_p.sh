#!/bin/bash
# round 2, GPU call P (1 GPU): decimation-in-time row kernel for ny = 32768 (HPXFFT_B200_ROWS_LONG=3) -- parity, then A/B bench
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decimation_in_time or r2c_rows" 2>&1 | tail -8 ) > gpurun_out/p_pytest.log
for v in 2 3; do
  HPXFFT_B200_ROWS_LONG=$v timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/p_bench_32768_v$v.json 2> gpurun_out/p_bench_32768_v$v.err
done
HPXFFT_B200_ROWS_LONG=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rows_dit2 -c 1 -o gpurun_out/p_rows_dit2 python bench.py --nx 32768 --ny 32768 --steps 1 --warmup 1 $B --no-parity > gpurun_out/p_ncu.log 2>&1
ls -la gpurun_out | grep " p_"
