#!/bin/bash
# Round 2, call 29: the driver's commands on the final tree (CTA pairs on): default bench, reference arm, C3 / C4 lines, smoke()
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2c29_c2.json 2> gpurun_out/r2c29_c2.err
echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c29_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['whole_step_frac'],d['cpu_baseline']['value'],d['gpu_baseline']['value'],d['shipped_dropouts']['ms_per_step'],d['also']['c5_audio_head_stress']['ms_per_step'])"
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/r2c29_ref.json 2> gpurun_out/r2c29_ref.err
echo "ref rc=$?"; cut -c1-160 gpurun_out/r2c29_ref.json
for cfg in c3 c4; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/r2c29_$cfg.json 2> gpurun_out/r2c29_$cfg.err
  echo "$cfg rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c29_$cfg.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['cpu_baseline']['value'],d['shipped_dropouts']['ms_per_step'],d['roofline']['frac'])"
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c29_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r2c29_smoke.log
