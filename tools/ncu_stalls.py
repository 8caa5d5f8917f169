"""Warp-stall breakdown and hottest SASS lines of one kernel of an .ncu-rep (`--set full --import-source on`).
usage: python tools/ncu_stalls.py REP KERNEL_REGEX [TOP]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = rows[1], rows[2:]
    i_s, i_src = hdr.index("# Samples"), hdr.index("Source")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[i_s]) for r in data)
    print(f"{rows[0][1]}: {tot} samples over {len(data)} SASS instructions")
    agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stall}
    print("  " + "  ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][i_s]))[:top_n]):
        r = data[i]
        st = {hdr[c]: int(r[c]) for c in stall if int(r[c]) > 0}
        print(f"  {i:5d} {int(r[i_s]):5d}  {r[i_src].strip()[:80]:80s} {max(st, key=st.get)[6:] if st else ''}")


if __name__ == "__main__":
    main()
