#!/bin/bash
# round 2, GPU call G (1 GPU): full GPU test suite, final N=1 bench line, mixed-radix timings, sanitizer, ncu captures of the final kernels
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline"
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/g_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench_16384.json 2> gpurun_out/g_bench_16384.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/g_ref_16384.json 2> gpurun_out/g_ref_16384.err
for n in 15872 12288 7680 8192 1536 2048; do
  timeout 200 python bench.py --nx $n --ny $n --steps 10 $B > gpurun_out/g_bench_${n}.json 2> gpurun_out/g_bench_${n}.err
done
timeout 200 python bench.py --config c1 --steps 20 $B > gpurun_out/g_bench_c1.json 2> gpurun_out/g_bench_c1.err
# sanitizer: memcheck + racecheck on <= 512^2 through the C++ example CLI
( cd gpurun_out && for tool in memcheck racecheck; do
    for sz in "512 512" "64 2048" "1024 64" "96 24"; do set -- $sz
      echo "== $tool nx=$1 ny=$2"; timeout 300 compute-sanitizer --tool $tool --error-exitcode 7 ../tests/cpp/hpxfft_shared_loop --nx=$1 --ny=$2 --plan=estimate 2>&1 | tail -4; echo "rc=$?"
    done; done ) > gpurun_out/g_sanitizer.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rows_r2c_v2|cols_fused_kernel' -s 6 -c 2 -f -o gpurun_out/g_ncu_16384 \
    python bench.py --steps 1 --warmup 3 $B --no-parity > gpurun_out/g_ncu_16384.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rows_long2|cols_fused_kernel' -s 6 -c 2 -f -o gpurun_out/g_ncu_32768 \
    python bench.py --nx 32768 --ny 32768 --steps 1 --warmup 3 $B --no-parity > gpurun_out/g_ncu_32768.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/g_launches_16384.csv python bench.py --steps 3 --warmup 3 $B --no-parity > gpurun_out/g_launches.log 2>&1
ls -la gpurun_out | grep " g_"
