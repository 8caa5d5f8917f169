"""In-situ (warm cache, real overlap) kernel breakdown of the native LRW training step with CUPTI through
torch.profiler, plus the busy fraction of the main stream and a CUDA-graph replay of forward+backward as an upper bound
on what removing launch gaps would buy. Developer tool; numbers here are never bench values."""
import collections
import ctypes as C
import re
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import B_PER_GPU, S, T, lrw_config  # noqa: E402
from syncvsr_b200._lib import check, lib  # noqa: E402
from syncvsr_b200.lightning import TransformerLightningModule  # noqa: E402
from syncvsr_b200.train import DataParallelStep, FusedAdamW  # noqa: E402

B = B_PER_GPU
torch.manual_seed(0)
m = TransformerLightningModule(lrw_config()).train()
step = DataParallelStep(m, FusedAdamW.from_config(m))
g = torch.Generator(device="cuda").manual_seed(1)
batch = (torch.randn(B, 1, T, S, S, device="cuda", generator=g), torch.randint(0, 320, (B, T * 4, 2), device="cuda", generator=g),
         torch.randint(0, 500, (B,), device="cuda", generator=g), torch.zeros(B, 1, device="cuda"))
for _ in range(5):
    step(*batch)
torch.cuda.synchronize()
N = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    step(*batch)
e1.record()
torch.cuda.synchronize()
print(f"eager step: {e0.elapsed_time(e1) / N:.3f} ms")

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step(*batch)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
spans = collections.defaultdict(list)
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and ev.device_time > 0:
        name = ev.name.replace("(anonymous namespace)::", "").replace("svsr::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        agg[name][0] += 1
        agg[name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"sum of kernel time: {tot / N / 1e3:.3f} ms/step over {sum(v[0] for v in agg.values()) // N} kernels")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1] / N / 1e3:8.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0] // N:4d}  avg {v[1] / v[0]:7.1f} us  {k[:70]}")

# ---- CUDA graph of forward + backward (no optimizer): replay vs eager ----
L = lib()


def fwd_bwd():
    m.flat_grads.zero_()
    with torch.no_grad():
        m(*batch)
    check(L.svsr_lrw_backward(m._h, C.c_void_p(0), m._stream()), "bwd")


for _ in range(3):
    fwd_bwd()
torch.cuda.synchronize()
e0.record()
for _ in range(N):
    fwd_bwd()
e1.record()
torch.cuda.synchronize()
print(f"eager fwd+bwd: {e0.elapsed_time(e1) / N:.3f} ms")
try:
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        fwd_bwd()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            fwd_bwd()
    torch.cuda.synchronize()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(N):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"graph fwd+bwd: {e0.elapsed_time(e1) / N:.3f} ms")
except Exception as ex:  # capture may be refused (e.g. an unsupported call inside the step)
    print("graph capture failed:", repr(ex)[:300])
