"""CutMix (LRW/video/src/augment.py:12-118): the host plan (same RNG draws, sequential in-place swaps replayed on index
tables) is checked on CPU against fixtures produced by the reference's own CutMix, and live against the reference module
when /root/reference is mounted; the device gather is checked bit-exactly on the GPU."""
import importlib.util
import sys
from pathlib import Path

import pytest
import torch

from oracle import ref_loader as rl

GOLD = Path(__file__).resolve().parent / "golden"
spec = importlib.util.spec_from_file_location("make_golden_cutmix", GOLD / "make_golden_cutmix.py")
mgc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mgc)


def _apply_plan_cpu(plan, videos, tokens, labels, wm, num_labels=500):
    """Plain-torch statement of what svsr_cutmix_gather computes (test oracle for the kernel)."""
    B, _, T = videos.shape[:3]
    Ta = tokens.shape[1]
    v = videos[plan.vsrc.long(), 0, torch.arange(T).unsqueeze(0).expand(B, T)].unsqueeze(1)
    a = tokens[plan.asrc.long(), torch.arange(Ta).unsqueeze(0).expand(B, Ta)]
    own = torch.nn.functional.one_hot(labels, num_labels).float()
    tar = torch.nn.functional.one_hot(labels[plan.tgt.long()], num_labels).float()
    r = plan.rate.unsqueeze(1)
    q = (1.0 - plan.rate.double()).float().unsqueeze(1)
    m = plan.mixed.bool().unsqueeze(1)
    soft = torch.where(m, q * own + r * tar, own)
    w = torch.where(m, q * wm + r * wm[plan.tgt.long()], wm)
    return v, a, soft, w


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_plan_reproduces_reference_cutmix_golden(name):
    from syncvsr_b200.augment import cutmix_plan

    fx = torch.load(GOLD / "cutmix.pt")[name]
    videos, tokens, labels, wm = mgc.make_inputs(seed=fx["seed"] + 1)
    torch.manual_seed(fx["seed"])
    plan = cutmix_plan(videos.shape[0], videos.shape[2], tokens.shape[1])
    v, a, soft, w = _apply_plan_cpu(plan, videos, tokens, labels, wm)
    assert torch.equal(v, fx["videos"]) and torch.equal(a, fx["tokens"])  # pure data movement: bit exact
    assert torch.equal(soft, fx["labels"]) and torch.equal(w, fx["word_mask"])
    assert int(plan.mixed.sum()) > 0 and (plan.vsrc != torch.arange(videos.shape[0], dtype=torch.int32).unsqueeze(1)).any()


@pytest.mark.skipif(not rl.reference_available(), reason="reference tree not mounted (GPU box)")
def test_plan_matches_live_reference_module_and_rng_stream():
    from syncvsr_b200.augment import cutmix_plan

    rl.load_reference_lrw()
    import augment as ref_aug

    for seed in (1, 2, 3, 4):
        videos, tokens, labels, wm = mgc.make_inputs(B=16, seed=100 + seed)
        torch.manual_seed(seed)
        rv, ra, rl_, rw = ref_aug.CutMix(500, None).eval()(videos.clone(), tokens.clone(), labels.clone(), wm.clone())
        after_ref = torch.rand(1)
        torch.manual_seed(seed)
        plan = cutmix_plan(16, 29, tokens.shape[1])
        after_mine = torch.rand(1)
        v, a, soft, w = _apply_plan_cpu(plan, videos, tokens, labels, wm)
        assert torch.equal(v, rv) and torch.equal(a, ra) and torch.equal(soft, rl_.float()) and torch.equal(w, rw)
        assert torch.equal(after_ref, after_mine)  # exactly the same number of RNG draws


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_device_gather_matches_reference_golden(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from syncvsr_b200.augment import CutMix

    fx = torch.load(GOLD / "cutmix.pt")[name]
    videos, tokens, labels, wm = mgc.make_inputs(seed=fx["seed"] + 1)
    torch.manual_seed(fx["seed"])
    v, a, soft, w = CutMix(500, None)(videos.cuda(), tokens.cuda(), labels.cuda(), wm.cuda())
    assert torch.equal(v.cpu(), fx["videos"]) and torch.equal(a.cpu(), fx["tokens"])
    assert torch.equal(soft.cpu(), fx["labels"]) and torch.equal(w.cpu(), fx["word_mask"])


@pytest.mark.gpu
def test_cutmix_feeds_the_native_step_with_soft_labels():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import test_lrw_gpu as tl
    from syncvsr_b200.augment import CutMix
    from syncvsr_b200.lightning import TransformerLightningModule

    m = TransformerLightningModule(tl.make_cfg(depth=1, use_wb=True)).train()
    g = torch.Generator(device="cuda").manual_seed(3)
    B = 4
    videos = torch.randn(B, 1, 29, 88, 88, device="cuda", generator=g)
    tokens = torch.randint(0, 320, (B, 116, 2), device="cuda", generator=g)
    labels = torch.randint(0, 500, (B,), device="cuda", generator=g)
    wm = (torch.rand(B, 29, device="cuda", generator=g) > 0.5).float()
    torch.manual_seed(5)
    batch = CutMix(500, None)(videos, tokens, labels, wm)
    assert batch[2].shape == (B, 500) and torch.allclose(batch[2].sum(1), torch.ones(B, device="cuda"), atol=1e-6)
    out = m(*batch)
    out["loss_total"].backward()
    assert torch.isfinite(out["loss_total"]) and torch.isfinite(m.flat_grads).all()
