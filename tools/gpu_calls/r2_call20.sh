#!/bin/bash
# Round 2, call 20: ncu evidence of the round-2 tree: time + DRAM bytes per launch of one kernel-by-kernel C2 step,
# --set full of the x-transformers tcgen05 attention kernels, the fused-head GEMM and one wide conv GEMM
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 2600 --csv \
  --log-file gpurun_out/r2c20_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > gpurun_out/r2c20_traffic.log 2>&1
echo "traffic rc=$?"; wc -l gpurun_out/r2c20_traffic.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attention_tc' --launch-skip 48 --launch-count 2 -f -o gpurun_out/r2c20_attn_tc_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > gpurun_out/r2c20_ncu1.log 2>&1
echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'igemm_kernel' --launch-skip 300 --launch-count 150 -f -o gpurun_out/r2c20_igemm_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > gpurun_out/r2c20_ncu2.log 2>&1
echo "ncu igemm rc=$?"; ls -la gpurun_out/r2c20_*.ncu-rep
