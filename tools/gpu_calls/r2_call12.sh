#!/bin/bash
# Round 2, call 12: attention tests (both paths), LRS suite, per-kernel times + --set full of the tcgen05 attention kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lrs_gpu.py -m gpu -q > gpurun_out/r2c12_lrs.log 2>&1
echo "lrs tests rc=$?"; tail -5 gpurun_out/r2c12_lrs.log
SVSR_ATTN_TC=0 timeout 600 python -m pytest tests/test_lrs_gpu.py -m gpu -q -k attention > gpurun_out/r2c12_attn_cc.log 2>&1
echo "attention tests (CUDA-core path) rc=$?"; tail -3 gpurun_out/r2c12_attn_cc.log
timeout 300 python tools/attn_rel_bench.py 20 | tee gpurun_out/r2c12_attn_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attn_rel|attention_core' --csv \
  --log-file gpurun_out/r2c12_attn_launches.csv python tools/attn_rel_bench.py 1 > /dev/null 2>&1
python - <<'PY'
import csv, re
rows = [l for l in open("gpurun_out/r2c12_attn_launches.csv") if not l.startswith("==")]
for r in csv.DictReader(rows):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("svsr::", "").replace("<unnamed>::", "").replace("void ", "")
        print(f'{name:45s} grid {r["Grid Size"]:16s} {r["Metric Value"]:>10s} {r["Metric Unit"]}')
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_rel' -c 4 -f -o gpurun_out/r2c12_attn_full \
  python tools/attn_rel_bench.py 1 > gpurun_out/r2c12_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r2c12_attn_full.ncu-rep
for tc in 1; do
  SVSR_ATTN_TC=$tc timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c12_c3_tc$tc.json 2> gpurun_out/r2c12_c3_tc$tc.err
  echo "c3 tc=$tc rc=$?"; cut -c1-330 gpurun_out/r2c12_c3_tc$tc.json; tail -2 gpurun_out/r2c12_c3_tc$tc.err
done
