#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_lrs_gpu.py -q -x > gpurun_out/lrs_tests.log 2>&1
echo "lrs tests rc=$?" | tee gpurun_out/lrs_rc.txt
tail -3 gpurun_out/lrs_tests.log
timeout 300 python tools/bench_lrs.py --T 150 --B 16 > gpurun_out/lrs2_c3.json 2> gpurun_out/lrs2_c3.err
timeout 300 python tools/bench_lrs.py --T 250 --B 8 > gpurun_out/lrs2_c4.json 2> gpurun_out/lrs2_c4.err
cut -c1-260 gpurun_out/lrs2_c3.json; echo; cut -c1-260 gpurun_out/lrs2_c4.json
