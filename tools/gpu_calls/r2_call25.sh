#!/bin/bash
# Round 2, call 25: CTA-pair (cta_group::2, M = 256) mode of the generic implicit GEMM: GEMM / conv parity tests with
# SVSR_IGEMM_2CTA=1, then the microbench against the single-CTA kernel and cuBLAS (every step under its own short timeout)
mkdir -p gpurun_out
SVSR_IGEMM_2CTA=1 timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "test_gemm" > gpurun_out/r2c25_gemm.log 2>&1
echo "gemm tests (pair) rc=$?"; tail -12 gpurun_out/r2c25_gemm.log
SVSR_IGEMM_2CTA=1 timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "conv or gemm" > gpurun_out/r2c25_conv.log 2>&1
echo "conv tests (pair) rc=$?"; tail -8 gpurun_out/r2c25_conv.log
echo "== pair"; SVSR_IGEMM_2CTA=1 timeout 120 python tools/gemm_bench.py 2>&1 | cut -c1-150 | tee gpurun_out/r2c25_bench_pair.txt
