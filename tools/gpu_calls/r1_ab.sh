#!/bin/bash
# One GPU call: new-kernel parity tests, stem kernel microbench, A/B of the step launch modes, then the whole GPU suite.
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "stem_temporal" > gpurun_out/ab_tests_kernels.log 2>&1
echo "kernel tests rc=$?" | tee gpurun_out/ab_rc.txt
timeout 600 python -m pytest tests/test_lrw_gpu.py -x -q -k "graph" > gpurun_out/ab_tests_graph.log 2>&1
echo "graph tests rc=$?" | tee -a gpurun_out/ab_rc.txt
timeout 300 python tools/stem_bench.py > gpurun_out/stem_bench.txt 2>&1
echo "stem bench rc=$?" | tee -a gpurun_out/ab_rc.txt
SVSR_STEM_HALO=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --graph 0 --priority 0 > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline --graph 0 --priority 0 > gpurun_out/ab_stem.json 2> gpurun_out/ab_stem.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline --graph 0 --priority 1 > gpurun_out/ab_stem_prio.json 2> gpurun_out/ab_stem_prio.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline --graph 1 --priority 0 > gpurun_out/ab_stem_graph.json 2> gpurun_out/ab_stem_graph.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline --graph 1 --priority 1 > gpurun_out/ab_stem_graph_prio.json 2> gpurun_out/ab_stem_graph_prio.err
for f in gpurun_out/ab_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])" 2>&1 | tail -1)"; done | tee gpurun_out/ab_summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a gpurun_out/ab_rc.txt
tail -3 gpurun_out/ab_tests_all.log
