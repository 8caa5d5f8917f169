#!/bin/bash
# Round 2, call 22: N-tile choice of the generic implicit GEMM on the encoder / Conformer shapes (forced 64 / 128 / 256 vs default)
mkdir -p gpurun_out
for bn in default 64 128 256; do
  echo "== SVSR_IGEMM_BN=$bn"
  if [ $bn = default ]; then timeout 300 python tools/gemm_bench.py; else SVSR_IGEMM_BN=$bn timeout 300 python tools/gemm_bench.py; fi
done 2>&1 | grep -v "^$" | tee gpurun_out/r2c22_gemm_bn.txt | cut -c1-140
