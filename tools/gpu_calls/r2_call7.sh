#!/bin/bash
# Round 2, call 7 (2 GPUs): whole suite on GPU 0, then torchrun N=2: LRW c2 (three-stage graph + staged all-reduce) and LRS c3 (staged)
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c7_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r2c7_tests.log
for cfg in c2 c3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 10 --warmup 3 --config $cfg > gpurun_out/r2c7_n2_$cfg.json 2> gpurun_out/r2c7_n2_$cfg.err
  echo "n2 $cfg rc=$?"; cut -c1-700 gpurun_out/r2c7_n2_$cfg.json; tail -3 gpurun_out/r2c7_n2_$cfg.err
done
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/r2c7_n1_c3.json 2> gpurun_out/r2c7_n1_c3.err
cut -c1-700 gpurun_out/r2c7_n1_c3.json
