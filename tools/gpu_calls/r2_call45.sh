#!/bin/bash
# Round 2, call 45 (4 GPUs): torchrun N=4 on the final tree (patch-free stem): LRW c2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29645 \
  bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2c45_n4_c2.json 2> gpurun_out/r2c45_n4_c2.err
echo "n4 c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c45_n4_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['n_gpus'],d.get('e2e'),d['config'].get('launch_mode'))"; tail -2 gpurun_out/r2c45_n4_c2.err
