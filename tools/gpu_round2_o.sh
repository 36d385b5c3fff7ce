#!/bin/bash
# round 2, GPU call O (1 GPU): super-group scheduling of the fused column kernel (HPXFFT_B200_COL_SG = 1 | 2 | 4)
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c2c_cols or c2_16384 or sweep" 2>&1 | tail -5 ) > gpurun_out/o_pytest.log
for sg in 1 2 4; do
  HPXFFT_B200_COL_SG=$sg timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/o_bench_16384_sg$sg.json 2> gpurun_out/o_bench_16384_sg$sg.err
done
HPXFFT_B200_COL_SG=2 HPXFFT_B200_LAG=3 timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/o_bench_16384_sg2_lag3.json 2> gpurun_out/o_bench_16384_sg2_lag3.err
HPXFFT_B200_COL_SG=4 HPXFFT_B200_LAG=2 timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/o_bench_16384_sg4_lag2.json 2> gpurun_out/o_bench_16384_sg4_lag2.err
for sg in 2 4; do
  HPXFFT_B200_COL_SG=$sg timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/o_bench_32768_sg$sg.json 2> gpurun_out/o_bench_32768_sg$sg.err
done
ls -la gpurun_out | grep " o_"
