#!/bin/bash
# Round 2, call 30: fused q|k|v projection + rotary + softmax + PV forward kernel: kernel test, LRW suite, C2 A/B
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "fused_qkv or attention" > gpurun_out/r2c30_k.log 2>&1
echo "kernel tests rc=$?"; tail -12 gpurun_out/r2c30_k.log
timeout 900 python -m pytest tests/test_lrw_gpu.py tests/test_optim_gpu.py -m gpu -q > gpurun_out/r2c30_lrw.log 2>&1
echo "lrw tests rc=$?"; tail -5 gpurun_out/r2c30_lrw.log
for on in 1 0; do
  SVSR_ATTN_QKV_FUSED=$on timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c30_c2_$on.json 2> gpurun_out/r2c30_c2_$on.err
  echo "c2 fused=$on rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c30_c2_$on.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['gpu_launches'],d['shipped_dropouts']['ms_per_step'])"
done
