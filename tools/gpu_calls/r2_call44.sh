#!/bin/bash
# Round 2, call 44 (2 GPUs): torchrun N=2 on the final tree (patch-free stem): LRW c2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29644 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c44_n2_c2.json 2> gpurun_out/r2c44_n2_c2.err
echo "n2 c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c44_n2_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['n_gpus'],d.get('e2e'),d['config'].get('launch_mode'))"; tail -2 gpurun_out/r2c44_n2_c2.err
