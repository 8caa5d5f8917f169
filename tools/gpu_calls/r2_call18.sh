#!/bin/bash
# Round 2, call 18: the driver's commands on the current tree (default bench, reference arm), C3 line with e2e / cpu_baseline / shipped_dropouts
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2c18_c2.json 2> gpurun_out/r2c18_c2.err
echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c18_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline']['value'],d['gpu_baseline']['value'],list(d['also']))"
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/r2c18_ref.json 2> gpurun_out/r2c18_ref.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/r2c18_ref.json
timeout 900 python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/r2c18_c3.json 2> gpurun_out/r2c18_c3.err
echo "c3 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2c18_c3.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['cpu_baseline'],d['shipped_dropouts'])"; tail -3 gpurun_out/r2c18_c3.err
timeout 600 python bench.py --impl reference --config c3 --steps 6 --warmup 1 > gpurun_out/r2c18_ref_c3.json 2> gpurun_out/r2c18_ref_c3.err
echo "ref c3 rc=$?"; cut -c1-200 gpurun_out/r2c18_ref_c3.json
