#!/bin/bash
# Round 2, call 4: BatchNorm-backward reductions fused into the dgrad epilogues: kernel test, whole suite, bench, trace.
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "fused_batchnorm_backward or halo or dgrad" > gpurun_out/r2c4_k.log 2>&1
echo "kernel rc=$?"; tail -25 gpurun_out/r2c4_k.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c4_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2c4_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
echo "bench rc=$?"; cat gpurun_out/r2c4_bench.json; tail -3 gpurun_out/r2c4_bench.err
SVSR_BN_BWD_FUSED=0 timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c4_bench_unfused.json 2>/dev/null
cat gpurun_out/r2c4_bench_unfused.json
timeout 300 python tools/step_trace.py r2c4 1 2>&1 | tail -1
