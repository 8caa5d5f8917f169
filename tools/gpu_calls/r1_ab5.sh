#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/ab6_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee gpurun_out/ab6_rc.txt
tail -3 gpurun_out/ab6_tests_all.log
timeout 300 python bench.py > gpurun_out/ab6_bench.json 2> gpurun_out/ab6_bench.err
cut -c1-330 gpurun_out/ab6_bench.json
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pack_all -c 4 --csv --log-file gpurun_out/ab6_pack.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph 0 --priority 0 > gpurun_out/ab6_pack.log 2>&1
grep pack_all gpurun_out/ab6_pack.csv | tail -3
