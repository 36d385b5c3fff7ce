#!/bin/bash
# round 2, GPU call H (2 GPUs): full 1-GPU test suite, then world-2 distributed parity and bench lines of the default (fused) and ce transports
mkdir -p gpurun_out
( CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/h_pytest_1gpu.log
( HPXFFT_B200_DIST_CASES=fast timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/h_pytest_dist2.log
run() { local name=$1; shift
  ( env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 $EXTRA > gpurun_out/h_bench_n2_$name.json 2> gpurun_out/h_bench_n2_$name.err ); }
EXTRA="" run default X=1
EXTRA="--no-e2e --no-anchor" run ce HPXFFT_B200_A2A=ce
ls -la gpurun_out | grep " h_"
