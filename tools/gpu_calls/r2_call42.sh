#!/bin/bash
# Round 2, call 42: stem without a patch tensor as the default? C2 A/B twice (40 steps each, alternating), C3 / C4 A/B
mkdir -p gpurun_out
for rep in a b; do for on in 0 1; do
  SVSR_STEM_DIRECT=$on timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c42_c2_${on}$rep.json 2> gpurun_out/r2c42_c2_${on}$rep.err
  echo "c2 direct=$on rep=$rep rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c42_c2_${on}$rep.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['shipped_dropouts']['ms_per_step'] if 'shipped_dropouts' in d else None)"
done; done
for cfg in c3 c4; do for on in 0 1; do
  SVSR_STEM_DIRECT=$on timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c42_${cfg}_$on.json 2> gpurun_out/r2c42_${cfg}_$on.err
  echo "$cfg direct=$on rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c42_${cfg}_$on.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'])"
done; done
