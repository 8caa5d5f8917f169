#!/bin/bash
# Round 2, call 23: where does the generic implicit GEMM lose time on the Conformer FFN shape? ncu --set full with source-level stalls
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'igemm_kernel' --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2c23_gemm_ffn \
  python tools/gemm_one.py 2400 3072 768 > gpurun_out/r2c23_ncu.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r2c23_gemm_ffn.ncu-rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'igemm_kernel' --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2c23_gemm_k3072 \
  python tools/gemm_one.py 2400 768 3072 > gpurun_out/r2c23_ncu2.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r2c23_gemm_k3072.ncu-rep
