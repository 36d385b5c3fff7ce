#!/bin/bash
# round 2, final 8-GPU call: the default all_to_all path (fused transport) at N = 8 (with e2e and the 1-GPU anchor) and N = 4
mkdir -p gpurun_out
run() { local n=$1 name=$2; shift 2
  ( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n $EXTRA > gpurun_out/l_bench_$name.json 2> gpurun_out/l_bench_$name.err ); tail -c 200 gpurun_out/l_bench_$name.err; }
EXTRA="--steps 20 --warmup 5" run 8 n8_default
EXTRA="--steps 10 --warmup 3 --no-e2e" run 4 n4_default
ls -la gpurun_out | grep " l_"
