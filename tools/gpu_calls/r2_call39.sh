#!/bin/bash
# Round 2, call 39 (8 GPUs): torchrun N=8 on the final tree: LRW c2 and LRS c3
mkdir -p gpurun_out
for cfg in c2 c3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29639 \
    bench.py --gpus 8 --steps 10 --warmup 3 --config $cfg > gpurun_out/r2c39_n8_$cfg.json 2> gpurun_out/r2c39_n8_$cfg.err
  echo "n8 $cfg rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c39_n8_$cfg.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['n_gpus'],d.get('e2e'),d['config'].get('launch_mode'))"; tail -2 gpurun_out/r2c39_n8_$cfg.err
done
