#!/bin/bash
# Round 2, call 10: tcgen05 rel-pos / decoder attention core (attention_rel_tc.cu): kernel parity tests, LRS suite, C3 bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lrs_gpu.py -m gpu -q -x -k "attention" > gpurun_out/r2c10_attn.log 2>&1
echo "attention tests rc=$?"; tail -25 gpurun_out/r2c10_attn.log
timeout 900 python -m pytest tests/test_lrs_gpu.py -m gpu -q > gpurun_out/r2c10_lrs.log 2>&1
echo "lrs tests rc=$?"; tail -8 gpurun_out/r2c10_lrs.log
for tc in 1 0; do
  SVSR_ATTN_TC=$tc timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c10_c3_tc$tc.json 2> gpurun_out/r2c10_c3_tc$tc.err
  echo "c3 tc=$tc rc=$?"; cut -c1-400 gpurun_out/r2c10_c3_tc$tc.json; tail -2 gpurun_out/r2c10_c3_tc$tc.err
done
