#!/bin/bash
# Round 2, call 38: final tree -- ncu time + DRAM bytes of one kernel-by-kernel C2 step (feeds roofline.traffic), then the
# driver's commands (default bench, reference arm, C3 / C4 lines)
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 2500 --csv \
  --log-file gpurun_out/r2c38_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > gpurun_out/r2c38_traffic.log 2>&1
echo "traffic rc=$?"; wc -l gpurun_out/r2c38_traffic.csv
timeout 900 python bench.py > gpurun_out/r2c38_c2.json 2> gpurun_out/r2c38_c2.err
echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c38_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['whole_step_frac'],d['cpu_baseline']['value'],d['gpu_baseline']['value'],d['shipped_dropouts']['ms_per_step'],d['also']['c5_audio_head_stress']['ms_per_step'],d['gpu_launches'])"
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/r2c38_ref.json 2> gpurun_out/r2c38_ref.err
echo "ref rc=$?"; cut -c1-160 gpurun_out/r2c38_ref.json
for cfg in c3 c4; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/r2c38_$cfg.json 2> gpurun_out/r2c38_$cfg.err
  echo "$cfg rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c38_$cfg.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['cpu_baseline']['value'],d['shipped_dropouts']['ms_per_step'],d['roofline']['frac'])"
done
