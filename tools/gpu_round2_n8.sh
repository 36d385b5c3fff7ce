#!/bin/bash
# round 2, 8-GPU call: parity at 8 ranks, bench lines for every transport, C5 (131072^2), message / strong sweeps
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n8_topo.txt 2>&1
( HPXFFT_B200_DIST_CASES=fast timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -k "equals_shared and $N" 2>&1 | tail -15 ) > gpurun_out/n8_pytest.log
run() { # name, timeout, env..., (EXTRA holds bench args)
  local name=$1 to=$2; shift 2
  ( env "$@" timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N $EXTRA > gpurun_out/n8_bench_$name.json 2> gpurun_out/n8_bench_$name.err )
  tail -c 300 gpurun_out/n8_bench_$name.err | tail -2
}
EXTRA="--steps 20 --warmup 5 --comm all_to_all" run ce 400 HPXFFT_B200_A2A=ce
EXTRA="--steps 10 --warmup 3 --comm p2p --no-e2e --no-anchor" run fused 300 X=1
EXTRA="--steps 10 --warmup 3 --comm all_to_all --no-e2e --no-anchor" run nccl 300 HPXFFT_B200_A2A=nccl
EXTRA="--steps 10 --warmup 3 --comm scatter --no-e2e --no-anchor" run scatter 300 X=1
EXTRA="--steps 10 --warmup 3 --comm all_to_all --no-e2e --no-anchor" run ce_chunks2 300 HPXFFT_B200_A2A=ce HPXFFT_B200_CHUNKS=2
EXTRA="--steps 10 --warmup 3 --comm all_to_all --no-e2e --no-anchor" run ce_chunks8 300 HPXFFT_B200_A2A=ce HPXFFT_B200_CHUNKS=8
EXTRA="--config c5 --steps 3 --warmup 3 --comm all_to_all" run c5_ce 600 HPXFFT_B200_A2A=ce
EXTRA="--config c5 --steps 3 --warmup 3 --comm p2p" run c5_fused 600 X=1
( cd gpurun_out && timeout 300 python ../benchmark/sweeps.py strong --loop 2 --out n8_sweeps > n8_sweep_strong.log 2>&1 )
( cd gpurun_out && timeout 300 python ../benchmark/sweeps.py message --loop 2 --out n8_sweeps > n8_sweep_message.log 2>&1 )
ls -la gpurun_out | grep " n8_"
