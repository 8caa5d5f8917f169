#!/bin/bash
# Round 2, call 2: two-epilogue-warpgroup igemm + fused audio head: head tests first, whole suite, bench (c2 + c5), warm profile.
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_head_gpu.py -x -q > gpurun_out/r2c2_head.log 2>&1
echo "head rc=$?"; tail -15 gpurun_out/r2c2_head.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c2_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2c2_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
echo "bench rc=$?"; cat gpurun_out/r2c2_bench.json; tail -3 gpurun_out/r2c2_bench.err
timeout 300 python bench.py --config c5 > gpurun_out/r2c2_c5.json 2> gpurun_out/r2c2_c5.err
cat gpurun_out/r2c2_c5.json; tail -3 gpurun_out/r2c2_c5.err
timeout 300 python tools/warm_profile.py > gpurun_out/r2c2_warm.txt 2>&1
head -12 gpurun_out/r2c2_warm.txt
