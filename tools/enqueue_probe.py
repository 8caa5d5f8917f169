"""How long does the host take to ENQUEUE one training step (no sync) vs how long the GPU takes to run it?"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import B_PER_GPU, S, T, lrw_config  # noqa: E402
from syncvsr_b200.lightning import TransformerLightningModule  # noqa: E402
from syncvsr_b200.train import DataParallelStep, FusedAdamW  # noqa: E402

m = TransformerLightningModule(lrw_config()).train()
step = DataParallelStep(m, FusedAdamW.from_config(m))
g = torch.Generator(device="cuda").manual_seed(1)
B = B_PER_GPU
batch = (torch.randn(B, 1, T, S, S, device="cuda", generator=g), torch.randint(0, 320, (B, T * 4, 2), device="cuda", generator=g),
         torch.randint(0, 500, (B,), device="cuda", generator=g), torch.zeros(B, 1, device="cuda"))
for _ in range(3):
    step(*batch)
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
# (a) enqueue time with an idle GPU queue kept short: sync each step
enq = []
for _ in range(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(*batch)
    enq.append(time.perf_counter() - t0)
torch.cuda.synchronize()
e0.record()
t0 = time.perf_counter()
for _ in range(n):
    step(*batch)
t_host = time.perf_counter() - t0
e1.record()
torch.cuda.synchronize()
print(f"host enqueue per step (GPU idle at start): {1e3*sum(enq)/n:.2f} ms ; back-to-back host loop {1e3*t_host/n:.2f} ms ; "
      f"GPU events {e0.elapsed_time(e1)/n:.2f} ms/step")
