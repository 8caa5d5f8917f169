"""Generates tests/golden/cutmix.pt from the UNMODIFIED reference CutMix (LRW/video/src/augment.py) run in the build
container: seeded CPU RNG, small synthetic batch; the outputs of the reference's sequential in-place mixing are stored
whole (they are small)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_loader as rl  # noqa: E402

OUT = Path(__file__).resolve().parent


def make_inputs(B=8, T=29, S=8, A=4, G=2, V=320, num_labels=500, seed=11):
    g = torch.Generator().manual_seed(seed)
    videos = torch.randn(B, 1, T, S, S, generator=g)
    tokens = torch.randint(0, V, (B, T * A, G), generator=g)
    labels = torch.randint(0, num_labels, (B,), generator=g)
    wm = (torch.rand(B, T, generator=g) > 0.5).float()
    return videos, tokens, labels, wm


if __name__ == "__main__":
    rl.load_reference_lrw()
    import augment as ref_aug  # the reference's module (sys.path set by the loader)

    cases = {}
    for name, seed in (("a", 123), ("b", 7), ("c", 2024)):
        videos, tokens, labels, wm = make_inputs(seed=seed + 1)
        torch.manual_seed(seed)
        v, a, l, w = ref_aug.CutMix(500, None).eval()(videos.clone(), tokens.clone(), labels.clone(), wm.clone())
        cases[name] = dict(seed=seed, videos=v.clone(), tokens=a.clone(), labels=l.float().clone(), word_mask=w.clone())
    torch.save(cases, OUT / "cutmix.pt")
    print({k: (float(v["videos"].sum()), int(v["tokens"].sum()), float(v["labels"].max())) for k, v in cases.items()})
