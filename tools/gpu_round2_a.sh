#!/bin/bash
# round 2, GPU call A (1 GPU): parity suite, bench lines, diagnostic variants, granule microbench, ncu of the 32768^2 kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt
B="--no-e2e --no-cpu-baseline"
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/a_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_16384.json 2> gpurun_out/a_bench_16384.err
timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/a_bench_32768.json 2> gpurun_out/a_bench_32768.err
for v in nomath wrap; do
  HPXFFT_B200_LIB=$PWD/hpx-fft_b200/libdiag_$v.so timeout 200 python bench.py --steps 10 $B --no-parity > gpurun_out/a_diag_${v}_16384.json 2> gpurun_out/a_diag_${v}_16384.err
  HPXFFT_B200_LIB=$PWD/hpx-fft_b200/libdiag_$v.so timeout 200 python bench.py --nx 32768 --ny 32768 --steps 6 $B --no-parity > gpurun_out/a_diag_${v}_32768.json 2> gpurun_out/a_diag_${v}_32768.err
done
HPXFFT_B200_LIB=$PWD/hpx-fft_b200/libdiag_cw32.so timeout 200 python bench.py --steps 10 $B > gpurun_out/a_diag_cw32_16384.json 2> gpurun_out/a_diag_cw32_16384.err
HPXFFT_B200_LIB=$PWD/hpx-fft_b200/libdiag_cw32.so timeout 200 python bench.py --nx 32768 --ny 32768 --steps 6 $B > gpurun_out/a_diag_cw32_32768.json 2> gpurun_out/a_diag_cw32_32768.err
for v in contig noprefetch; do
  HPXFFT_B200_LIB=$PWD/hpx-fft_b200/libdiag_$v.so timeout 200 python bench.py --nx 32768 --ny 32768 --steps 6 $B --no-parity > gpurun_out/a_diag_${v}_32768.json 2> gpurun_out/a_diag_${v}_32768.err
done
timeout 300 tools/microbench/granule > gpurun_out/a_granule.csv 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rows_r2c_kernel|cols_fused_kernel' -s 4 -c 2 -f -o gpurun_out/a_ncu_32768 \
    python bench.py --nx 32768 --ny 32768 --steps 1 --warmup 3 $B --no-parity > gpurun_out/a_ncu_32768.log 2>&1
ls -la gpurun_out
