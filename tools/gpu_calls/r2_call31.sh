#!/bin/bash
# Round 2, call 31: ncu --set full of ONE 8192^3 launch of the implicit GEMM as CTA pairs (cta_group::2) and single-CTA; the fused
# q|k|v attention forward inside a bench step
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none -k regex:'igemm_kernel' --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2c31_gemm_pair \
  python tools/gemm_one.py 8192 8192 8192 > gpurun_out/r2c31_a.log 2>&1
echo "pair rc=$?"
SVSR_IGEMM_2CTA=0 timeout 200 ncu --set full --clock-control none -k regex:'igemm_kernel' --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2c31_gemm_single \
  python tools/gemm_one.py 8192 8192 8192 > gpurun_out/r2c31_b.log 2>&1
echo "single rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attention_qkv' --launch-skip 36 --launch-count 1 -f -o gpurun_out/r2c31_attn_qkv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > gpurun_out/r2c31_c.log 2>&1
echo "attn rc=$?"; ls -la gpurun_out/r2c31_*.ncu-rep
