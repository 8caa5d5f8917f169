"""The measurement tooling is part of the evidence chain (bench.py reads roofline.traffic from the file
tools/ncu_traffic.py writes): run the summarisers over the committed ncu passes and check what DESIGN.md quotes."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run(tool, *args):
    r = subprocess.run([sys.executable, str(ROOT / "tools" / tool), *map(str, args)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_traffic_summary_uses_the_last_complete_step(tmp_path):
    out = tmp_path / "traffic.json"
    run("ncu_traffic.py", ROOT / "profiles" / "r2_traffic_final.csv", out)
    fam = json.loads(out.read_text())["families"]
    # one step = 137 launches of the implicit-GEMM family (generic + both halo kernels; round 2 added the backward launch of
    # the fused audio head and moved the 12 q|k|v projections into the fused attention kernel) and 70 of the weight-gradient one
    assert fam["igemm_kernel"]["launches"] == 137 and fam["wgrad_kernel"]["launches"] == 70
    committed = json.loads((ROOT / "profiles" / "r2_traffic.json").read_text())["families"]["igemm_kernel"]
    assert abs(fam["igemm_kernel"]["dram_bytes_per_launch"] - committed["dram_bytes_per_launch"]) < 1.0
    # AdamW streams 28 bytes per parameter: the pass sits at the HBM roofline
    assert fam["adamw_seg_kernel"]["dram_gbs"] > 5500


def test_launch_summary_and_layer_roofline_agree_on_the_step():
    text = run("launch_summary.py", ROOT / "profiles" / "r2_traffic_final.csv")
    assert "441 launches" in text.splitlines()[0]
    table = run("layer_roofline.py", ROOT / "profiles" / "r2_traffic_final.csv")
    rows = [l for l in table.splitlines() if l.startswith("| ") and "launch |" not in l]
    assert len(rows) == 1 + 19 + 6 + 1  # stem, 19 trunk convs, first and last encoder layer (to_out, ff1, ff2), total
    stem = next(l for l in rows if l.startswith("| stem"))
    assert "conv_stem_direct_kernel" in stem  # the patch-free stem forward (csrc/stem_direct.cu)
    total = [c.strip() for c in rows[-1].split("|")]
    assert 400 < float(total[6]) < 700  # forward GEMM launches: ~500 TFLOP/s under ncu
