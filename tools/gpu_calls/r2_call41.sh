#!/bin/bash
# Round 2, call 41: stem without a patch tensor, second cut (all producer loads in flight before the first store, eight
# forward producer warps): kernel + model tests, per-launch times, C2 A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "patch_tensor or patch_free or uncovered_frames" > gpurun_out/r2c41_new.log 2>&1
echo "new tests rc=$?"; tail -4 gpurun_out/r2c41_new.log
timeout 200 python tools/stem_bench.py direct > gpurun_out/r2c41_stem_bench.txt 2>&1; echo "bench rc=$?"; cat gpurun_out/r2c41_stem_bench.txt
for on in 0 1; do
  SVSR_STEM_DIRECT=$on timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c41_c2_$on.json 2> gpurun_out/r2c41_c2_$on.err
  echo "c2 direct=$on rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c41_c2_$on.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['gpu_launches'],d['config'].get('loss_total'))"
done
