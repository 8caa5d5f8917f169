#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/ab4_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee gpurun_out/ab4_rc.txt
tail -3 gpurun_out/ab4_tests_all.log
timeout 300 python tools/encoder_kernel_bench.py > gpurun_out/ab4_encoder_kernels.txt 2>&1
cat gpurun_out/ab4_encoder_kernels.txt
timeout 300 python bench.py > gpurun_out/ab4_bench.json 2> gpurun_out/ab4_bench.err
cut -c1-400 gpurun_out/ab4_bench.json
