#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_multi.sh N [tests]
N=$1
mkdir -p gpurun_out
if [[ "$2" == "tests" ]]; then
  timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_dist$N.log 2>&1
  tail -4 gpurun_out/pytest_dist$N.log | cut -c1-300
fi
for run in all_to_all scatter p2p; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 5 --warmup 3 --comm $run --e2e-steps 1 > gpurun_out/bench_n${N}_$run.json 2> gpurun_out/bench_n${N}_$run.err
  python - <<PY
import json
txt=open("gpurun_out/bench_n${N}_$run.json").read().strip().splitlines()
ls=[l for l in txt if l.startswith("{")]
if ls:
    d=json.loads(ls[-1]); print("$run N=$N", "ms=%.3f"%d["ms_per_step"], "GF=%.0f"%d["value"], {k:round(v,3) for k,v in d["phases_ms"].items()}, "e2e ms=%.1f"%d["e2e"]["ms_per_step"])
else:
    print("$run N=$N FAILED"); print(open("gpurun_out/bench_n${N}_$run.err").read()[-1500:])
PY
done
