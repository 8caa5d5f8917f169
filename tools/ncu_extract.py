"""Extract a compact per-launch table from an .ncu-rep (`ncu -i rep --page raw --csv`): time, DRAM bytes and achieved
GB/s, tensor-pipe / SM / L2 / DRAM utilisation -- the numbers DESIGN.md and profiles/*.md quote."""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fma.sum",
        "smsp__inst_executed.sum"]


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {w: hdr.index(w) for w in WANT if w in hdr}
    print("| kernel | grid | time us | DRAM MB (r+w) | DRAM GB/s | dram % | tensor % | sm % | L2 % | L1 % | warps % | regs | dyn smem KB |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in data:
        g = lambda k: r[idx[k]] if k in idx else ""
        u = lambda k: units[idx[k]] if k in idx else ""
        t = float(g("gpu__time_duration.sum").replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u("gpu__time_duration.sum"), 1)
        mb = (to_bytes(g("dram__bytes_read.sum"), u("dram__bytes_read.sum")) + to_bytes(g("dram__bytes_write.sum"), u("dram__bytes_write.sum"))) / 1e6
        name = g("Kernel Name").replace("svsr::", "").replace("<unnamed>::", "").replace("void ", "")[:46]
        f = lambda k: f"{float(g(k).replace(',', '')):.1f}" if g(k) not in ("", "n/a") else "-"
        smem = g("launch__shared_mem_per_block_dynamic")
        print(f"| {name} | {g('Grid Size')} | {t:.1f} | {mb:.1f} | {mb / t * 1e3:.0f} | {f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | "
              f"{f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | {f('sm__throughput.avg.pct_of_peak_sustained_elapsed')} | "
              f"{f('lts__throughput.avg.pct_of_peak_sustained_elapsed')} | {f('l1tex__throughput.avg.pct_of_peak_sustained_elapsed')} | "
              f"{f('sm__warps_active.avg.pct_of_peak_sustained_active')} | {g('launch__registers_per_thread')} | {smem} |")


if __name__ == "__main__":
    main()
