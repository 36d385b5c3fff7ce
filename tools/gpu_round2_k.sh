#!/bin/bash
# round 2, GPU call K (1 GPU): final library -- full GPU test suite, final bench lines
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/k_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/k_bench_16384.json 2> gpurun_out/k_bench_16384.err
timeout 300 python bench.py --nx 32768 --ny 32768 --steps 10 $B > gpurun_out/k_bench_32768.json 2> gpurun_out/k_bench_32768.err
for n in 15872 12288 7680 1536; do
  timeout 200 python bench.py --nx $n --ny $n --steps 10 $B > gpurun_out/k_bench_${n}.json 2> gpurun_out/k_bench_${n}.err
done
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/k_smoke.log 2>&1
ls -la gpurun_out | grep " k_"
