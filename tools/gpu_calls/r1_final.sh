#!/bin/bash
# Final evidence of the round: both bench arms, the ncu launch list of the kernel-by-kernel step, one ncu --set full
# capture of the two temporal-halo stem kernels, smoke().
mkdir -p gpurun_out
set -x
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 300 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/final_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --priority 0 > gpurun_out/final_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"t5_c64" --launch-skip 24 --launch-count 2 \
  -o gpurun_out/stem_halo_full python tools/stem_bench.py > gpurun_out/stem_halo_full.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$?"
cat gpurun_out/final_bench.json
ls -la gpurun_out/*.ncu-rep
timeout 300 python tools/bench_lrs.py --T 150 --B 16 > gpurun_out/final_lrs_c3.json 2> gpurun_out/final_lrs_c3.err
timeout 300 python tools/bench_lrs.py --T 250 --B 8 > gpurun_out/final_lrs_c4.json 2> gpurun_out/final_lrs_c4.err
tail -c 400 gpurun_out/final_lrs_c3.json
