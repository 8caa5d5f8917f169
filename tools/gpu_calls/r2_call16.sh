#!/bin/bash
# Round 2, call 16: vectorised bias-gradient column sums + unrolled depthwise-conv weight gradient: whole GPU suite, C2/C3/C4 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c16_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2c16_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c16_c2.json 2> gpurun_out/r2c16_c2.err
echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c16_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['roofline']['frac'],d.get('shipped_dropouts',{}).get('ms_per_step'))"
for cfg in c3 c4; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c16_$cfg.json 2> gpurun_out/r2c16_$cfg.err
  echo "$cfg rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c16_$cfg.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['config']['launch_mode'],d['gpu_launches'])"; tail -2 gpurun_out/r2c16_$cfg.err
done
