#!/bin/bash
# Round 2, call 5: tcgen05 attention core for the x-transformers encoder: kernel tests, whole suite, bench A/B, trace.
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/r2c5_k.log 2>&1
echo "kernel rc=$?"; tail -30 gpurun_out/r2c5_k.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c5_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2c5_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2c5_bench.json; tail -3 gpurun_out/r2c5_bench.err
SVSR_ATTN_TC=0 timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c5_bench_cudacore.json 2>/dev/null
cut -c1-400 gpurun_out/r2c5_bench_cudacore.json
timeout 300 python tools/step_trace.py r2c5 1 2>&1 | tail -1
