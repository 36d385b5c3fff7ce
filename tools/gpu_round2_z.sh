#!/bin/bash
# round 2, GPU call Z (1 GPU): ncu --set full of rows_ditc_kernel<8> (2048 x 131072)
mkdir -p gpurun_out
timeout 55 ncu --set full --clock-control none --import-source on -k regex:rows_ditc -c 1 -o gpurun_out/z_rows_ditc8 python bench.py --nx 2048 --ny 131072 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/z_ncu.log 2>&1
ls -la gpurun_out | grep " z_"
