#!/bin/bash
# Second GPU call: whole GPU suite, A/B of the overlapped repack, then the final evidence script.
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/ab2_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee gpurun_out/ab2_rc.txt
tail -15 gpurun_out/ab2_tests_all.log
SVSR_PACK_OVERLAP=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/ab2_nooverlap.json 2> gpurun_out/ab2_nooverlap.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/ab2_default.json 2> gpurun_out/ab2_default.err
for f in gpurun_out/ab2_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])" 2>&1 | tail -1)"; done | tee gpurun_out/ab2_summary.txt
bash tools/gpu_calls/r1_final.sh
