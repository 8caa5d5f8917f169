#!/bin/bash
# Round 2, call 27: CTA pairs on by default in the implicit GEMM and now in the weight-gradient kernel: kernel tests, whole suite, C2 / C3 A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv or gemm or wgrad" > gpurun_out/r2c27_k.log 2>&1
echo "kernel tests rc=$?"; tail -8 gpurun_out/r2c27_k.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c27_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2c27_tests.log
for on in 1 0; do
  SVSR_IGEMM_2CTA=$on timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c27_c2_$on.json 2> gpurun_out/r2c27_c2_$on.err
  echo "c2 pair=$on rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c27_c2_$on.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['roofline']['avg_launch_us'],d['roofline']['share_of_step'])"
  SVSR_IGEMM_2CTA=$on timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c27_c3_$on.json 2> gpurun_out/r2c27_c3_$on.err
  echo "c3 pair=$on rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c27_c3_$on.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'])"
done
