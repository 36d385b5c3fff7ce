#!/bin/bash
# round 2, GPU call E (1 GPU): v2 row kernel with XOR swizzle; column-kernel skeleton diagnostics
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "r2c_rows or c2_16384 or sweep" 2>&1 | tail -5 ) > gpurun_out/e_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/e_bench_16384.json 2> gpurun_out/e_bench_16384.err
for v in skel nostorev noloadi noio; do
  HPXFFT_B200_LIB=$PWD/hpx-fft_b200/libdiag_$v.so HPXFFT_B200_ROWS_V1=1 timeout 200 python bench.py --steps 10 $B --no-parity > gpurun_out/e_diag_${v}_16384.json 2> gpurun_out/e_diag_${v}_16384.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rows_r2c_v2' -s 3 -c 1 -f -o gpurun_out/e_ncu_rows16384 \
    python bench.py --steps 1 --warmup 3 $B --no-parity > gpurun_out/e_ncu_rows16384.log 2>&1
ls -la gpurun_out | grep " e_"
