#!/bin/bash
# Round 2, call 26: CTA-pair mode end to end: whole GPU suite with SVSR_IGEMM_2CTA=1, then C2 / C3 bench A/B
mkdir -p gpurun_out
SVSR_IGEMM_2CTA=1 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c26_tests.log 2>&1
echo "tests (pair) rc=$?"; tail -6 gpurun_out/r2c26_tests.log
for on in 1 0; do
  SVSR_IGEMM_2CTA=$on timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c26_c2_$on.json 2> gpurun_out/r2c26_c2_$on.err
  echo "c2 pair=$on rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c26_c2_$on.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['roofline']['avg_launch_us'])"
  SVSR_IGEMM_2CTA=$on timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c26_c3_$on.json 2> gpurun_out/r2c26_c3_$on.err
  echo "c3 pair=$on rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c26_c3_$on.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'])"
done
