#!/bin/bash
# round 2, GPU call X (1 GPU): ny = 16384 row kernel with the pencil refill issued in groups between the steps of the tail (HPXFFT_B200_ROWS_ILV=1)
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( HPXFFT_B200_ROWS_ILV=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "r2c_rows or c2_16384 or golden or anchor or x_dependence or end_to_end" 2>&1 | tail -4 ) > gpurun_out/x_pytest.log
for v in 0 1 0 1; do
  HPXFFT_B200_ROWS_ILV=$v timeout 300 python bench.py --steps 30 --warmup 5 $B > gpurun_out/x_bench_16384_ilv$v.json 2> gpurun_out/x_bench_16384_ilv$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/x_bench_16384_ilv$v.json").read().strip().splitlines()[-1])
print("ilv$v", d["ms_per_step"], d["phases_ms"]["rows_kernel"], d["phases_ms"]["cols_kernel"], d["parity"]["rel_l2"])
PY
done > gpurun_out/x_summary.txt 2>&1
cat gpurun_out/x_summary.txt
