#!/bin/bash
# Round 2, call 1: whole GPU suite on the tree with the optimizer/engine-cache/advice fixes, both bench arms plus the new
# torch.cuda eager comparator, warm in-situ kernel profile as the baseline for this round's kernel work.
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_tests.log 2>&1
echo "tests rc=$?"
tail -5 gpurun_out/r2c1_tests.log
timeout 600 python bench.py > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench rc=$?"
cat gpurun_out/r2c1_bench.json
tail -3 gpurun_out/r2c1_bench.err
timeout 300 python bench.py --impl eager --steps 20 --warmup 5 > gpurun_out/r2c1_eager.json 2> gpurun_out/r2c1_eager.err
cat gpurun_out/r2c1_eager.json; tail -3 gpurun_out/r2c1_eager.err
timeout 300 python tools/warm_profile.py > gpurun_out/r2c1_warm.txt 2>&1
head -50 gpurun_out/r2c1_warm.txt
