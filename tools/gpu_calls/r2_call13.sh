#!/bin/bash
# Round 2, call 13: attention_rel_tc v2 + tail-tile warp skipping: LRS suite, microbench, ncu times, C3/C4 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lrs_gpu.py -m gpu -q > gpurun_out/r2c13_lrs.log 2>&1
echo "lrs tests rc=$?"; tail -3 gpurun_out/r2c13_lrs.log
timeout 300 python tools/attn_rel_bench.py 20 | tee gpurun_out/r2c13_attn_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attn_rel' --csv \
  --log-file gpurun_out/r2c13_attn_launches.csv python tools/attn_rel_bench.py 1 > /dev/null 2>&1
grep -o 'attn_rel[a-z_]*kernel[^"]*".*' gpurun_out/r2c13_attn_launches.csv | awk -F'","' '{print $1, $NF}' | head -20
for cfg in c3 c4; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c13_$cfg.json 2> gpurun_out/r2c13_$cfg.err
  echo "$cfg rc=$?"; cut -c1-330 gpurun_out/r2c13_$cfg.json; tail -2 gpurun_out/r2c13_$cfg.err
done
