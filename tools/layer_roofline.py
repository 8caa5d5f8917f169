"""Per-launch achieved TFLOP/s of the forward implicit-GEMM launches of one LRW bench step (B = 64 clips, 1856 frames),
from an `ncu --metrics gpu__time_duration.sum` launch list (cold-cache, serialised, ~1.75 GHz under ncu): the launch order of
the forward pass is fixed (stem, 4 + 5 + 5 + 5 trunk convs, then per encoder layer qkv / out / ff1 / ff2, then the heads),
so each launch's algorithmic FLOPs follow from the layer geometry. Usage: layer_roofline.py profiles/r1_launches_final.csv"""
import csv
import re
import sys

NOMINAL, SUSTAINED = 2250.0, 1384.7  # dense bf16 TFLOP/s: nominal, MEASURED_PEAKS.json sustained


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    recs = []
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("svsr::", "").replace("<unnamed>::", "")
        recs.append((name, float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"]))
    idx = [i for i, r in enumerate(recs) if "stem_patch" in r[0] or "conv_stem_direct" in r[0]]  # first stem launch of a step
    segs = [recs[a:b] for a, b in zip(idx, idx[1:] + [len(recs)])]
    full = max(len(s) for s in segs)
    return [s for s in segs if len(s) == full][-1]


def main():
    step = load(sys.argv[1])
    fam = ("igemm_kernel", "conv3x3_c64_halo_kernel", "conv_t5_c64_halo_kernel", "conv_stem_direct_kernel")
    gemms = [r for r in step if r[0].startswith(fam)]
    N = 64 * 29
    layers = [("stem temporal conv 5 taps 64->64 (K 245 of 320)", 2.0 * N * 44 * 44 * 64 * 245)]
    def conv(name, hw, cin, cout, k):
        layers.append((name, 2.0 * N * hw * hw * cout * cin * k * k))
    for b in range(2):
        conv(f"layer1.{b}.conv1 3x3 64->64 @22", 22, 64, 64, 3), conv(f"layer1.{b}.conv2 3x3 64->64 @22", 22, 64, 64, 3)
    for li, (hw, cin, cout) in enumerate([(11, 64, 128), (6, 128, 256), (3, 256, 512)], start=2):
        conv(f"layer{li}.0.conv1 3x3/s2 {cin}->{cout} @{hw}", hw, cin, cout, 3)
        conv(f"layer{li}.0.conv2 3x3 {cout}->{cout} @{hw}", hw, cout, cout, 3)
        conv(f"layer{li}.0.downsample 1x1/s2 {cin}->{cout} @{hw}", hw, cin, cout, 1)
        conv(f"layer{li}.1.conv1 3x3 {cout}->{cout} @{hw}", hw, cout, cout, 3)
        conv(f"layer{li}.1.conv2 3x3 {cout}->{cout} @{hw}", hw, cout, cout, 3)
    M = 64 * 30
    # round 2: the q | k | v projection runs inside the fused attention kernel (attention_qkv_tc_fwd_kernel) when it appears
    # in the launch list, so an encoder layer contributes three implicit-GEMM launches instead of four
    fused_qkv = any("attention_qkv" in r[0] for r in step)
    for i in range(12):
        if not fused_qkv:
            layers += [(f"encoder.{i} to_qkv [1920,512]x[1536,512]", 2.0 * M * 1536 * 512)]
        layers += [(f"encoder.{i} to_out [1920,512]x[512,512]", 2.0 * M * 512 * 512),
                   (f"encoder.{i} ff GEGLU proj [1920,512]x[4096,512]", 2.0 * M * 4096 * 512),
                   (f"encoder.{i} ff out [1920,2048]x[512,2048]", 2.0 * M * 512 * 2048)]
    print("| launch | kernel | grid | us | GFLOP | TFLOP/s | % nominal 2250 | % sustained 1385 |")
    print("|---|---|---|---|---|---|---|---|")
    tot_f = tot_t = 0.0
    for (name, fl), (k, us, grid) in zip(layers, gemms):
        if name.startswith("encoder.") and not name.startswith(("encoder.0 ", "encoder.11 ")):
            tot_f += fl; tot_t += us
            continue
        tf = fl / us * 1e-6
        tot_f += fl; tot_t += us
        print(f"| {name} | {k} | {grid} | {us:.1f} | {fl / 1e9:.1f} | {tf:.0f} | {100 * tf / NOMINAL:.0f} | {100 * tf / SUSTAINED:.0f} |")
    tf = tot_f / tot_t * 1e-6
    print(f"| **forward GEMM launches, all {len(layers)}** | | | {tot_t:.0f} | {tot_f / 1e9:.0f} | {tf:.0f} | {100 * tf / NOMINAL:.0f} | {100 * tf / SUSTAINED:.0f} |")


if __name__ == "__main__":
    main()
