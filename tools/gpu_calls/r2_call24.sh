#!/bin/bash
# Round 2, call 24 (2 GPUs): whole GPU suite + smoke() on GPU 0, then torchrun N=2: C2 and C3 (graph replay + staged all-reduce + shipped dropouts)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c24_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2c24_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c24_smoke.log 2>&1
echo "smoke rc=$?"; tail -4 gpurun_out/r2c24_smoke.log
for cfg in c2 c3; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29624 \
    bench.py --gpus 2 --steps 10 --warmup 3 --config $cfg > gpurun_out/r2c24_n2_$cfg.json 2> gpurun_out/r2c24_n2_$cfg.err
  echo "n2 $cfg rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c24_n2_$cfg.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['n_gpus'],d.get('e2e'),d.get('shipped_dropouts'))"; tail -2 gpurun_out/r2c24_n2_$cfg.err
done
