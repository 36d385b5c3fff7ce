#!/bin/bash
# round 2, GPU call U (2 GPUs): 32768^2 bench with the default (fused) transport after the transport-aware row-kernel dispatch
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/u_bench_n2.json 2> gpurun_out/u_bench_n2.err
HPXFFT_B200_A2A=ce timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/u_bench_n2_ce.json 2> gpurun_out/u_bench_n2_ce.err
ls -la gpurun_out | grep " u_"
