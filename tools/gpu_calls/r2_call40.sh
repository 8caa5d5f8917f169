#!/bin/bash
# Round 2, call 40: parity mode of the word-boundary / HuggingFace variants; the stem without a patch tensor
# (stem_direct.cu): new tests first, whole suite with the default path and with SVSR_STEM_DIRECT=1, C2 / C3 A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "parity_mode or patch_tensor or patch_free or uncovered_frames" > gpurun_out/r2c40_new.log 2>&1
echo "new tests rc=$?"; tail -25 gpurun_out/r2c40_new.log
timeout 200 python tools/parity_report.py > gpurun_out/r2c40_parity.txt 2>&1; echo "parity rc=$?"; tail -12 gpurun_out/r2c40_parity.txt
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2c40_tests.log 2>&1
echo "suite rc=$?"; tail -6 gpurun_out/r2c40_tests.log
SVSR_STEM_DIRECT=1 timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2c40_tests_direct.log 2>&1
echo "suite(direct) rc=$?"; tail -6 gpurun_out/r2c40_tests_direct.log
for on in 0 1; do
  SVSR_STEM_DIRECT=$on timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c40_c2_$on.json 2> gpurun_out/r2c40_c2_$on.err
  echo "c2 direct=$on rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c40_c2_$on.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['gpu_launches'],d['config'].get('loss_total'))"
done
for on in 0 1; do
  SVSR_STEM_DIRECT=$on timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c40_c3_$on.json 2> gpurun_out/r2c40_c3_$on.err
  echo "c3 direct=$on rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c40_c3_$on.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'])"
done
