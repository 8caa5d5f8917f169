#!/bin/bash
# Round 2, call 37: ReLU mask of the trunk's block outputs applied in the producing dgrad epilogue (BatchNorm-backward passes
# read one tensor less): LRW suite + kernel suite, A/B C2, two-stream hazard test
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c37_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2c37_tests.log
for on in 1 0; do
  SVSR_RELU_MASK_IN_DGRAD=$on timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c37_c2_$on.json 2> gpurun_out/r2c37_c2_$on.err
  echo "c2 mask_in_dgrad=$on rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c37_c2_$on.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['config'].get('loss_total'))"
done
