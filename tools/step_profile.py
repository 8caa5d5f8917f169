"""Runs WARM warm-up steps and then STEPS steps of the native LRW training step (B=64, 12 layers) -- the command that is
wrapped in ncu for the per-launch time list and the --set full capture (see profiles/)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import B_PER_GPU, S, T, lrw_config  # noqa: E402
from syncvsr_b200.lightning import TransformerLightningModule  # noqa: E402
from syncvsr_b200.train import DataParallelStep, FusedAdamW  # noqa: E402

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = int(sys.argv[3]) if len(sys.argv) > 3 else B_PER_GPU
torch.manual_seed(0)
m = TransformerLightningModule(lrw_config()).train()
step = DataParallelStep(m, FusedAdamW.from_config(m))
g = torch.Generator(device="cuda").manual_seed(1)
batch = (torch.randn(B, 1, T, S, S, device="cuda", generator=g), torch.randint(0, 320, (B, T * 4, 2), device="cuda", generator=g),
         torch.randint(0, 500, (B,), device="cuda", generator=g), torch.zeros(B, 1, device="cuda"))
for _ in range(warm):
    step(*batch)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
for _ in range(steps):
    out = step(*batch)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("loss", float(out["loss_total"]))
