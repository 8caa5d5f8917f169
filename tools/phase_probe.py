"""Time the phases of one training step separately (forward / backward / optimizer+repack), B=64."""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import B_PER_GPU, S, T, lrw_config  # noqa: E402
from syncvsr_b200._lib import check, lib  # noqa: E402
from syncvsr_b200.lightning import TransformerLightningModule  # noqa: E402
from syncvsr_b200.train import FusedAdamW  # noqa: E402

m = TransformerLightningModule(lrw_config()).train()
opt = FusedAdamW.from_config(m)
g = torch.Generator(device="cuda").manual_seed(1)
B = B_PER_GPU
batch = (torch.randn(B, 1, T, S, S, device="cuda", generator=g), torch.randint(0, 320, (B, T * 4, 2), device="cuda", generator=g),
         torch.randint(0, 500, (B,), device="cuda", generator=g), torch.zeros(B, 1, device="cuda"))


def ev():
    return torch.cuda.Event(enable_timing=True)


acc = {"pack": 0.0, "fwd": 0.0, "bwd0": 0.0, "bwd1": 0.0, "opt": 0.0}
n = 12
for it in range(n + 3):
    e = [ev() for _ in range(6)]
    torch.cuda.synchronize()
    m.flat_grads.zero_()
    e[0].record()
    m._ensure(batch[0])  # repack
    e[1].record()
    with torch.no_grad():
        m(*batch)
    e[2].record()
    check(lib().svsr_lrw_backward_stage(m._h, C.c_void_p(0), C.c_int(0), m._stream()), "b0")
    e[3].record()
    check(lib().svsr_lrw_backward_stage(m._h, C.c_void_p(0), C.c_int(1), m._stream()), "b1")
    e[4].record()
    opt.step()
    e[5].record()
    torch.cuda.synchronize()
    if it >= 3:
        for k, (a, b) in zip(acc, zip(e[:-1], e[1:])):
            acc[k] += a.elapsed_time(b)
print("phase ms/step:", {k: round(v / n, 3) for k, v in acc.items()}, "sum", round(sum(acc.values()) / n, 3))
