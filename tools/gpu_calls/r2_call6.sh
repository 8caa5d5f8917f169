#!/bin/bash
# Round 2, call 6: device-resident step control (graph replay under the shipped dropouts): tests, suite, bench.
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_lrw_gpu.py -x -q -k "dropout or skip or graph" > gpurun_out/r2c6_k.log 2>&1
echo "focus rc=$?"; tail -30 gpurun_out/r2c6_k.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c6_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2c6_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
echo "bench rc=$?"; cat gpurun_out/r2c6_bench.json; tail -3 gpurun_out/r2c6_bench.err
