#!/bin/bash
# Round 2, call 35: stem backward apply as 2x2 input-pixel quads (four window loads per four pixels): kernel tests, whole suite, per-kernel time, C2 / C3
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "stem" > gpurun_out/r2c35_k.log 2>&1
echo "stem kernel tests rc=$?"; tail -5 gpurun_out/r2c35_k.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c35_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2c35_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'stem_bwd_apply' --launch-skip 3 --launch-count 2 --csv --log-file gpurun_out/r2c35_pool.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > /dev/null 2>&1
grep -o 'stem_bwd_apply[^"]*".*' gpurun_out/r2c35_pool.csv | awk -F'","' '{print $1, $NF}' | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c35_c2.json 2> gpurun_out/r2c35_c2.err
echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c35_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'])"
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c35_c3.json 2> gpurun_out/r2c35_c3.err
echo "c3 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c35_c3.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'])"
