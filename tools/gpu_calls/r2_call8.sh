#!/bin/bash
# Round 2, call 8 (8 GPUs): torchrun N=8: LRW c2 and LRS c3 / c4 with the staged all-reduces
mkdir -p gpurun_out
set -x
for cfg in c2 c3 c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 8 --steps 10 --warmup 3 --config $cfg > gpurun_out/r2c8_n8_$cfg.json 2> gpurun_out/r2c8_n8_$cfg.err
  echo "n8 $cfg rc=$?"; cut -c1-260 gpurun_out/r2c8_n8_$cfg.json; tail -2 gpurun_out/r2c8_n8_$cfg.err
done
