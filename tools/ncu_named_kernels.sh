#!/bin/bash
# One `ncu --set full` capture of the kernels north_star names (conv stem: temporal implicit GEMM + BN/GELU/pool; attention)
# on the LRW bench step (B=64) and the LRS C3 step; summaries are extracted from the .ncu-rep with tools/ncu_extract.py.
set -x
OUT=gpurun_out
timeout 500 ncu --set full --clock-control none -k regex:"stem_|attention_fwd|attention_bwd|rmsnorm_bwd" --launch-skip 60 --launch-count 12 \
  -o $OUT/lrw_named_full python tools/step_profile.py 2 1 > $OUT/lrw_named_full.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:"igemm_kernel" --launch-skip 296 --launch-count 12 \
  -o $OUT/lrw_igemm_fwd_full python tools/step_profile.py 2 1 > $OUT/lrw_igemm_fwd_full.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:"attention_core" --launch-skip 48 --launch-count 5 \
  -o $OUT/lrs_attn_full python tools/bench_lrs.py --steps 1 --warmup 1 > $OUT/lrs_attn_full.log 2>&1
ls -la $OUT/*.ncu-rep
