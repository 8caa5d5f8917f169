#!/bin/bash
# Last GPU call of the round: whole GPU suite on the final tree, DRAM-traffic pass for roofline.traffic, smoke().
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/last_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee gpurun_out/last_rc.txt
tail -4 gpurun_out/last_tests_all.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1700 --csv \
  --log-file gpurun_out/lrw_traffic_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --priority 0 > gpurun_out/last_traffic.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/last_rc.txt
