#!/bin/bash
# round 2, GPU call R (1 GPU): decimation-in-time row kernels over C sample classes (HPXFFT_B200_ROWS_LONG=5), ny = 32768 / 65536 / 131072
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decimation_in_time" 2>&1 | tail -8 ) > gpurun_out/r_pytest.log
HPXFFT_B200_ROWS_LONG=5 timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/r_bench_32768_v5.json 2> gpurun_out/r_bench_32768_v5.err
for v in 2 5; do
  HPXFFT_B200_ROWS_LONG=$v timeout 300 python bench.py --nx 4096 --ny 65536 --steps 10 $B > gpurun_out/r_bench_4096x65536_v$v.json 2> gpurun_out/r_bench_4096x65536_v$v.err
  HPXFFT_B200_ROWS_LONG=$v timeout 300 python bench.py --nx 2048 --ny 131072 --steps 10 $B > gpurun_out/r_bench_2048x131072_v$v.json 2> gpurun_out/r_bench_2048x131072_v$v.err
done
ls -la gpurun_out | grep " r_"
