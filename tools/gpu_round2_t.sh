#!/bin/bash
# round 2, GPU call T (2 GPUs): the new default long-row kernels with two destination ranks (all transports) + 32768^2 sampled + bench
mkdir -p gpurun_out
( HPXFFT_B200_DIST_CASES=rows timeout 900 python -m pytest "tests/test_gpu_distributed.py::test_distributed_equals_shared[2]" -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/t_pytest_dist2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/t_bench_n2.json 2> gpurun_out/t_bench_n2.err
ls -la gpurun_out | grep " t_"
