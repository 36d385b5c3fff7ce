#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_a2a_ab.sh N   -- A/B of all_to_all exchange variants
N=$1
run() {
  tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 5 --warmup 3 --comm all_to_all --e2e-steps 2 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
ls=[l for l in open("gpurun_out/ab_$tag.json").read().splitlines() if l.startswith("{")]
if ls:
    d=json.loads(ls[-1]); p=d["phases_ms"]
    print("$tag N=$N ms=%.3f rows=%.3f comm1=%.3f cols=%.3f comm2=%.3f unpack=%.3f e2e=%.1f"%(d["ms_per_step"],p["rows_kernel"],p["first_comm"],p["cols_kernel"],p["second_comm"],p["second_trans"],d["e2e"]["ms_per_step"]))
else:
    print("$tag FAILED", open("gpurun_out/ab_$tag.err").read()[-600:])
PY
}
run grouped X=1
run stepwise HPXFFT_B200_A2A_STEPWISE=1
run grouped_ch32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run stepwise_ch32 HPXFFT_B200_A2A_STEPWISE=1 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
