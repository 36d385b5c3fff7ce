#!/bin/bash
# round 2, GPU call W (1 GPU): general output addressing (P = 1) of the ny = 32768 row kernels, for the record; launch list of the 32768^2 bench
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline --no-parity"
for v in 0 2 5; do
  HPXFFT_B200_ROWS_GENERAL=1 HPXFFT_B200_ROWS_LONG=$v timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/w_bench_32768_general_v$v.json 2> gpurun_out/w_bench_32768_general_v$v.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/w_launches_32768.csv python bench.py --nx 32768 --ny 32768 --steps 2 --warmup 3 $B > gpurun_out/w_ncu_bench.log 2>&1
ls -la gpurun_out | grep " w_"
