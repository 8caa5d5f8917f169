#!/bin/bash
# Round 2, call 14: CUDA-graph replay of the LRS sentence-level step: parity test, C3/C4 bench graph vs kernel by kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lrs_gpu.py -m gpu -q -k "graph_replayed" > gpurun_out/r2c14_graph.log 2>&1
echo "graph test rc=$?"; tail -15 gpurun_out/r2c14_graph.log
for cfg in c3 c4; do for gr in 1 0; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --graph $gr --no-cpu-baseline > gpurun_out/r2c14_${cfg}_g$gr.json 2> gpurun_out/r2c14_${cfg}_g$gr.err
  echo "$cfg graph=$gr rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c14_${cfg}_g$gr.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['config']['launch_mode'],d['gpu_launches'],d['config']['loss'])"; tail -2 gpurun_out/r2c14_${cfg}_g$gr.err
done; done
