#!/bin/bash
# Round 2, call 15: LRS graph tests (staged + unstaged), ncu launch list of one kernel-by-kernel C3 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lrs_gpu.py -m gpu -q -k "graph_replayed" > gpurun_out/r2c15_graph.log 2>&1
echo "graph tests rc=$?"; tail -4 gpurun_out/r2c15_graph.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3790 -c 1400 --csv --log-file gpurun_out/r2c15_c3_launches.csv \
  python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --graph 0 > gpurun_out/r2c15_c3_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2c15_c3_launches.csv
