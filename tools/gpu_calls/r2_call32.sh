#!/bin/bash
# Round 2, call 32: bn_bwd_finalize folded into the apply launch: whole GPU suite, C2 / C3 bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c32_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2c32_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c32_c2.json 2> gpurun_out/r2c32_c2.err
echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c32_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['gpu_launches'],d['roofline']['frac'])"
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c32_c3.json 2> gpurun_out/r2c32_c3.err
echo "c3 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c32_c3.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['gpu_launches'])"
