#!/bin/bash
# Round 2, call 19: device-resident LRS step seed (graph replay under the shipped dropouts): LRS + kernel suites, C3 line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_lrs_gpu.py tests/test_kernels_gpu.py -m gpu -q > gpurun_out/r2c19_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r2c19_tests.log
timeout 900 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c19_c3.json 2> gpurun_out/r2c19_c3.err
echo "c3 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c19_c3.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['shipped_dropouts'])"; tail -3 gpurun_out/r2c19_c3.err
