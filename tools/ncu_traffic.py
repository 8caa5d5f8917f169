"""Per-kernel-family DRAM traffic of the LAST training step in an ncu csv captured with
`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`: writes a small JSON
(profiles/<round>_traffic.json) that bench.py reports as `roofline.traffic` (bytes per launch of the dominant family)."""
import collections
import csv
import json
import re
import sys


def main():
    path, out = sys.argv[1], sys.argv[2]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per_id = collections.OrderedDict()
    for r in rows:
        kid = r["ID"]
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("svsr::", "").replace("<unnamed>::", "").replace("void ", "")
        d = per_id.setdefault(kid, {"name": name})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        if r["Metric Name"].startswith("dram__bytes"):
            mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
            d[r["Metric Name"]] = v * mult
        elif r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
    ks = list(per_id.values())
    idx = [i for i, k in enumerate(ks) if "stem_patch" in k["name"] or "conv_stem_direct" in k["name"]]  # first stem launch of a step
    # the last COMPLETE step: a capture cut by `ncu -c N` ends inside a step, which shows as a shorter last segment
    segs = [ks[a:b] for a, b in zip(idx, idx[1:] + [len(ks)])] if idx else [ks]
    full = max(len(sg) for sg in segs)
    step = [sg for sg in segs if len(sg) == full][-1]
    fam = collections.defaultdict(lambda: {"launches": 0, "bytes": 0.0, "us": 0.0})
    for k in step:
        f = re.sub(r"<.*", "", k["name"]).strip()
        if f in ("conv3x3_c64_halo_kernel", "conv_t5_c64_halo_kernel", "conv_stem_direct_kernel"):  # bench.py's PROF_IGEMM family
            f = "igemm_kernel"
        if f in ("wgrad3x3_c64_halo_kernel", "wgrad_t5_c64_halo_kernel", "wgrad_stem_direct_kernel"):  # bench.py's PROF_WGRAD family
            f = "wgrad_kernel"
        fam[f]["launches"] += 1
        fam[f]["bytes"] += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
        fam[f]["us"] += k.get("us", 0.0)
    res = {f: {"launches": v["launches"], "dram_bytes_per_launch": v["bytes"] / v["launches"],
               "dram_bytes_per_step": v["bytes"], "us_per_step": v["us"],
               "dram_gbs": v["bytes"] / max(v["us"], 1e-9) / 1e3} for f, v in fam.items()}
    res["_total"] = {"dram_bytes_per_step": sum(v["bytes"] for v in fam.values()), "us_per_step": sum(v["us"] for v in fam.values())}
    json.dump({"source": path, "note": "one ncu pass over the last step (serialised, cold cache)", "families": res}, open(out, "w"), indent=1)
    for f, v in sorted(res.items(), key=lambda kv: -kv[1].get("us_per_step", 0)):
        if f != "_total":
            print(f"{f:32s} n={v['launches']:4d} {v['dram_bytes_per_step']/1e6:9.1f} MB {v['us_per_step']/1e3:8.3f} ms {v['dram_gbs']:8.0f} GB/s")
    print("total", res["_total"])


if __name__ == "__main__":
    main()
