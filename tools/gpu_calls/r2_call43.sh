#!/bin/bash
# Round 2, call 43: final tree (stem without a patch tensor by default) -- ncu time + DRAM bytes of one kernel-by-kernel
# C2 step (feeds roofline.traffic), ncu --set full of the two patch-free stem kernels, whole GPU suite, smoke(), the
# driver's default bench command
mkdir -p gpurun_out
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 2500 --csv \
  --log-file gpurun_out/r2c43_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > gpurun_out/r2c43_traffic.log 2>&1
echo "traffic rc=$?"; wc -l gpurun_out/r2c43_traffic.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_direct -c 2 -f -o gpurun_out/r2c43_stem_direct \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --graph 0 --priority 0 > gpurun_out/r2c43_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r2c43_stem_direct.ncu-rep
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2c43_tests.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/r2c43_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c43_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r2c43_smoke.log
timeout 900 python bench.py > gpurun_out/r2c43_c2.json 2> gpurun_out/r2c43_c2.err
echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c43_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['whole_step_frac'],d['cpu_baseline']['value'],d['gpu_baseline']['value'],d['shipped_dropouts']['ms_per_step'],d['also']['c5_audio_head_stress']['ms_per_step'],d['gpu_launches'])"
