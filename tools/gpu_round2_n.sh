#!/bin/bash
# round 2, GPU call N (1 GPU): output-store cache policy (.cs evict-first vs .cg normal priority)
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/n_bench_16384_cs.json 2> gpurun_out/n_bench_16384_cs.err
HPXFFT_B200_LIB=$PWD/hpx-fft_b200/libdiag_storecg.so timeout 300 python bench.py --steps 20 --warmup 5 $B > gpurun_out/n_bench_16384_cg.json 2> gpurun_out/n_bench_16384_cg.err
HPXFFT_B200_LIB=$PWD/hpx-fft_b200/libdiag_storecg.so timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/n_bench_32768_cg.json 2> gpurun_out/n_bench_32768_cg.err
ls -la gpurun_out | grep " n_"
