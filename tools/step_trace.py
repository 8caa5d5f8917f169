"""Timeline of ONE warm native LRW training step (graph replay, high-priority main stream -- the bench's launch mode):
every kernel with its stream, start, duration, grid. Writes gpurun_out/<tag>_trace.csv (developer tool; profiler numbers
are never bench values). Usage: python tools/step_trace.py [tag] [graph=1|0]"""
import json
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import B_PER_GPU, S, T, lrw_config  # noqa: E402
from syncvsr_b200.lightning import TransformerLightningModule  # noqa: E402
from syncvsr_b200.train import DataParallelStep, FusedAdamW  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "step"
use_graph = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
B = B_PER_GPU
torch.manual_seed(0)
m = TransformerLightningModule(lrw_config()).train()
step = DataParallelStep(m, FusedAdamW.from_config(m), graph=use_graph, high_priority=True)
g = torch.Generator(device="cuda").manual_seed(1)
batch = (torch.randn(B, 1, T, S, S, device="cuda", generator=g), torch.randint(0, 320, (B, T * 4, 2), device="cuda", generator=g),
         torch.randint(0, 500, (B,), device="cuda", generator=g), torch.zeros(B, 1, device="cuda"))
for _ in range(6):
    step(*batch)
torch.cuda.synchronize()
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step(*batch)
    torch.cuda.synchronize()
trace = out / f"{tag}_trace.json"
prof.export_chrome_trace(str(trace))
ev = [e for e in json.loads(trace.read_text())["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# keep the middle step: from the 2nd zero-fill / first kernel after the 1st adamw to the next adamw
names = [e["name"] for e in ev]
ad = [i for i, n in enumerate(names) if "adamw" in n]
seg = ev[ad[0] + 1: ad[1] + 1] if len(ad) >= 2 else ev
t0 = seg[0]["ts"]
with open(out / f"{tag}_trace.csv", "w") as f:
    f.write("start_us,dur_us,stream,grid,block,regs,smem,name\n")
    for e in seg:
        a = e.get("args", {})
        name = e["name"].replace("svsr::", "").replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        f.write(f"{e['ts'] - t0:.2f},{e['dur']:.2f},{a.get('stream', '')},\"{a.get('grid', '')}\",\"{a.get('block', '')}\","
                f"{a.get('registers per thread', '')},{a.get('shared memory', '')},{name[:60]}\n")
trace.unlink()
print(f"step span {seg[-1]['ts'] + seg[-1]['dur'] - t0:.1f} us over {len(seg)} activities -> {tag}_trace.csv")
