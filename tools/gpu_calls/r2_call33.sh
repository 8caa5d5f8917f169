#!/bin/bash
# Round 2, call 33: CTA pairs also under the fused cross-entropy epilogue of the audio head: head tests, LRW + LRS suites, C5 / C2 bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -q > gpurun_out/r2c33_head.log 2>&1
echo "head tests rc=$?"; tail -4 gpurun_out/r2c33_head.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c33_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2c33_tests.log
for on in 1 0; do
SVSR_IGEMM_2CTA=$on timeout 300 python bench.py --config c5 --steps 20 --warmup 5 > gpurun_out/r2c33_c5_$on.json 2> gpurun_out/r2c33_c5_$on.err
echo "c5 pair=$on rc=$?"; cut -c1-420 gpurun_out/r2c33_c5_$on.json
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c33_c2.json 2> gpurun_out/r2c33_c2.err
echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c33_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['roofline']['frac'])"
