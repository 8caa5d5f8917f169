#!/bin/bash
# Round 2, call 21 (redo of call 20 with bounded captures): time + DRAM bytes per launch of one kernel-by-kernel C2 step,
# --set full of 2 x-transformers attention launches and 16 implicit-GEMM launches of the forward trunk
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 2600 --csv \
  --log-file gpurun_out/r2c21_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > gpurun_out/r2c21_traffic.log 2>&1
echo "traffic rc=$?"; wc -l gpurun_out/r2c21_traffic.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attention_tc' --launch-skip 48 --launch-count 2 -f -o gpurun_out/r2c21_attn_tc_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > gpurun_out/r2c21_ncu1.log 2>&1
echo "ncu attn rc=$?"
timeout 400 ncu --set full --clock-control none -k regex:'igemm_kernel' --launch-skip 447 --launch-count 16 -f -o gpurun_out/r2c21_igemm_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > gpurun_out/r2c21_ncu2.log 2>&1
echo "ncu igemm rc=$?"; ls -la gpurun_out/r2c21_*.ncu-rep; du -sh gpurun_out
