"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the LAST training
step in the file (from the last stem launch -- stem_patch, or conv_stem_direct on the patch-free path -- on) and, with --grids, per-grid-size detail."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    recs = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("svsr::", "").replace("<unnamed>::", "").replace("void ", "")
        recs.append((name, v, row["Grid Size"]))
    return recs


def main():
    recs = load(sys.argv[1])
    idx = [i for i, r in enumerate(recs) if "stem_patch" in r[0] or "conv_stem_direct" in r[0]]  # first stem launch of a step
    # the last COMPLETE step: a capture cut by `ncu -c N` ends inside a step, which shows as a shorter last segment
    segs = [recs[a:b] for a, b in zip(idx, idx[1:] + [len(recs)])]
    full = max(len(sg) for sg in segs)
    step = [sg for sg in segs if len(sg) == full][-1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, us, _ in step:
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"last complete step: {tot/1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches (ncu-serialised, cold cache)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]/1e3:8.3f} ms {100*v[1]/tot:5.1f}%  n={v[0]:4d}  {k[:80]}")
    if "--grids" in sys.argv:
        for name in sorted(agg, key=lambda k: -agg[k][1])[:8]:
            c, t = collections.Counter(), collections.defaultdict(float)
            for n, us, g in step:
                if n == name:
                    c[g] += 1
                    t[g] += us
            print(name)
            for g, n in sorted(c.items(), key=lambda kv: -t[kv[0]]):
                print(f"    grid {g:20s} n={n:3d} total {t[g]/1e3:7.3f} ms  avg {t[g]/n:8.1f} us")


if __name__ == "__main__":
    main()
