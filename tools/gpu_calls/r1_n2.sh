#!/bin/bash
mkdir -p gpurun_out
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2_graph.json 2> gpurun_out/n2_graph.err
echo "rc=$?"
timeout 400 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --graph 0 --priority 0 > gpurun_out/n2_eager.json 2> gpurun_out/n2_eager.err
echo "rc=$?"
grep -h "capture failed" gpurun_out/n2_*.err
tail -c 600 gpurun_out/n2_graph.json; echo; tail -c 300 gpurun_out/n2_graph.err
for f in gpurun_out/n2_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['config'].get('launch_mode'))" 2>&1 | tail -1)"; done | tee gpurun_out/n2_summary.txt
