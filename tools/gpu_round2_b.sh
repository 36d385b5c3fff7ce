#!/bin/bash
# round 2, GPU call B (2 GPUs): distributed parity (all transports), bench lines per transport
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/b_gpus.txt
nvidia-smi topo -m > gpurun_out/b_topo.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "r2c_rows or sweep or c3_kernels or golden" 2>&1 | tail -30 ) > gpurun_out/b_pytest_rows.log
timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/b_bench_n1_32768_rowslong.json 2> gpurun_out/b_bench_n1_32768_rowslong.err
HPXFFT_B200_ROWS_OLD=1 timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/b_bench_n1_32768_rowsold.json 2> gpurun_out/b_bench_n1_32768_rowsold.err
if [ "$FULLDIST" = 1 ]; then ( timeout 1500 python -m pytest tests/test_gpu_distributed.py tests/test_cpp_dropin.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/b_pytest.log; fi
run() { # name, env..., args
  local name=$1; shift
  ( env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 $EXTRA > gpurun_out/b_bench_n2_$name.json 2> gpurun_out/b_bench_n2_$name.err )
}
EXTRA="--comm all_to_all" run ce HPXFFT_B200_A2A=ce
EXTRA="--comm all_to_all --no-e2e --no-anchor" run ce_chunks1 HPXFFT_B200_A2A=ce HPXFFT_B200_CHUNKS=1
EXTRA="--comm all_to_all --no-e2e --no-anchor" run ce_chunks8 HPXFFT_B200_A2A=ce HPXFFT_B200_CHUNKS=8
EXTRA="--comm all_to_all --no-e2e --no-anchor" run nccl HPXFFT_B200_A2A=nccl
EXTRA="--comm p2p --no-e2e --no-anchor" run fused X=1
EXTRA="--comm scatter --no-e2e --no-anchor" run scatter X=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/b_ref_n2.json 2> gpurun_out/b_ref_n2.err
ls -la gpurun_out | grep " b_"
