#!/bin/bash
# Round 2, call 9 (8 GPUs): how much of the N=8 loss is NCCL's CTAs taking SMs from the persistent 148-CTA kernels? sweep
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus 8 --steps 20 --warmup 5 --config c2 > gpurun_out/r2c9_$name.json 2> gpurun_out/r2c9_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c9_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "ms", round(d["value"]), "clips/s", "e2e", round(d["e2e"]["ms_per_step"],3))
except Exception as e:
    print("$name failed", e)
PY
}
run default A=1
run ctas8 NCCL_MAX_CTAS=8
run ctas4 NCCL_MAX_CTAS=4
run ctas2 NCCL_MAX_CTAS=2
run nvls NCCL_ALGO=NVLS
run nvls_ctas4 NCCL_ALGO=NVLS NCCL_MAX_CTAS=4
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 8 --steps 3 --warmup 3 --config c2 2>&1 | grep -iE "nvls|channels|Algo|Ring|Tree" | head -20 > gpurun_out/r2c9_nccl_info.txt
head -12 gpurun_out/r2c9_nccl_info.txt
