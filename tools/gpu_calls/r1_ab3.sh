#!/bin/bash
# Third GPU call: encoder-kernel changes -- parity tests, warm microbench (new build), step bench.
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/ab3_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee gpurun_out/ab3_rc.txt
tail -4 gpurun_out/ab3_tests_all.log
timeout 300 python tools/encoder_kernel_bench.py > gpurun_out/ab3_encoder_kernels.txt 2>&1
cat gpurun_out/ab3_encoder_kernels.txt
timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/ab3_default.json 2> gpurun_out/ab3_default.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/ab3_default_b.json 2> gpurun_out/ab3_default_b.err
for f in gpurun_out/ab3_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])" 2>&1 | tail -1)"; done | tee gpurun_out/ab3_summary.txt
