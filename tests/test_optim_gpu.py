"""Parity of the fused clip + AdamW step (csrc/optim.cu, `svsr_adamw_step[_segmented]`) with what the reference runs:
torch.optim.AdamW over two parameter groups (decay on ndim >= 2 only, LRW/video/src/lightning.py:216-221) after
`clip_grad_norm_(gradient_clip_val)` (Lightning Trainer, train.py:32), including DDP's gradient averaging (grad_div) and
the `p.grad is None` semantics of sublayers dropped by x-transformers' layer_dropout (lightning.py:95-105)."""
import ctypes as C

import pytest
import torch

from oracle import lrw_oracle as O

pytestmark = pytest.mark.gpu


def _lib():
    from syncvsr_b200._lib import check, lib

    return lib(), check


def _torch_reference(p0, grads, n_decay, lr, betas, eps, wd, max_norm, grad_div, ranges=None, active=None):
    """torch.optim.AdamW + clip_grad_norm_ over the same numbers. `ranges`: [(begin, end)] parameter tensors; `active[s][i]`
    False = tensor i has grad None on step s."""
    n = p0.numel()
    ranges = ranges or [(0, n_decay), (n_decay, n)]
    params = [torch.nn.Parameter(p0[a:b].clone()) for a, b in ranges if b > a]
    ranges = [(a, b) for a, b in ranges if b > a]
    dec = [p for p, (a, b) in zip(params, ranges) if b <= n_decay]
    nod = [p for p, (a, b) in zip(params, ranges) if a >= n_decay]
    assert len(dec) + len(nod) == len(params)
    opt = torch.optim.AdamW([{"params": dec}, {"params": nod, "weight_decay": 0.0}], lr=lr, betas=betas, eps=eps,
                            weight_decay=wd)
    for s, g in enumerate(grads):
        for i, (p, (a, b)) in enumerate(zip(params, ranges)):
            on = active is None or active[s][i]
            p.grad = (g[a:b] / grad_div).clone() if on else None
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], max_norm)
        opt.step()
    out = p0.clone()
    for p, (a, b) in zip(params, ranges):
        out[a:b] = p.data
    return out


@pytest.mark.parametrize("grad_div,max_norm", [(1.0, 1.0), (8.0, 5.0), (2.0, 1e9)])
def test_adamw_step_matches_torch_adamw_and_clip(grad_div, max_norm):
    L, check = _lib()
    torch.manual_seed(0)
    n, n_decay = 40_000 * 4, 30_000 * 4
    p0 = torch.randn(n, device="cuda")
    grads = [torch.randn(n, device="cuda") * (0.05 * (s + 1)) for s in range(3)]
    lr, betas, eps, wd = 1e-3, (0.9, 0.999), 1e-6, 0.01
    ref = _torch_reference(p0, grads, n_decay, lr, betas, eps, wd, max_norm, grad_div)
    for seg in (False, True):
        p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
        for s, g in enumerate(grads):
            if seg:
                begin = (C.c_int64 * 3)(0, n_decay // 2, n)
                steps = (C.c_int32 * 2)(s + 1, s + 1)
                check(L.svsr_adamw_step_segmented(
                    C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(m.data_ptr()),
                    C.c_void_p(v.data_ptr()), C.c_int64(n_decay), C.c_int64(n), C.c_float(lr), C.c_float(betas[0]),
                    C.c_float(betas[1]), C.c_float(eps), C.c_float(wd), begin, steps, C.c_int(2), C.c_float(max_norm),
                    C.c_float(grad_div), C.c_void_p(scratch.data_ptr()), C.c_void_p(0)), "adamw_seg")
            else:
                check(L.svsr_adamw_step(
                    C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(m.data_ptr()),
                    C.c_void_p(v.data_ptr()), C.c_int64(n_decay), C.c_int64(n), C.c_float(lr), C.c_float(betas[0]),
                    C.c_float(betas[1]), C.c_float(eps), C.c_float(wd), C.c_int(s + 1), C.c_float(max_norm),
                    C.c_float(grad_div), C.c_void_p(scratch.data_ptr()), C.c_void_p(0)), "adamw")
            # the norm the kernel reports is clip_grad_norm_'s total norm of the averaged gradient
            assert float(scratch.view(torch.float32)[3]) == pytest.approx(float((g / grad_div).norm()), rel=1e-5)
        torch.cuda.synchronize()
        # fp32 arithmetic in a different association order: a few ulps of the update, far below lr
        assert (p - ref).abs().max().item() < 2e-6, seg
        assert ((p - ref).norm() / (ref - p0).norm()).item() < 1e-5, seg


def test_ranges_without_gradient_are_left_untouched_like_grad_none():
    """A range with step count 0 = `p.grad is None` for torch.optim.AdamW: no weight decay, no moment update, and its
    later bias corrections use ITS OWN step count."""
    L, check = _lib()
    torch.manual_seed(1)
    q = 4096
    n, n_decay = 6 * q, 4 * q
    ranges = [(i * q, (i + 1) * q) for i in range(6)]
    p0 = torch.randn(n, device="cuda")
    grads = [torch.randn(n, device="cuda") * 0.1 for _ in range(4)]
    active = [[True] * 6, [True, False, True, True, False, True], [True, False, True, True, True, True], [True] * 6]
    for s in range(4):  # a tensor without gradient contributes nothing to the clip norm either: zero its slice
        for i, (a, b) in enumerate(ranges):
            if not active[s][i]:
                grads[s][a:b] = 0
    lr, betas, eps, wd = 1e-2, (0.9, 0.98), 1e-6, 0.03
    ref = _torch_reference(p0, grads, n_decay, lr, betas, eps, wd, 1.0, 1.0, ranges, active)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
    count = [0] * 6
    begin = (C.c_int64 * 7)(*[a for a, _ in ranges], n)
    for s, g in enumerate(grads):
        for i in range(6):
            count[i] += int(active[s][i])
        steps = (C.c_int32 * 6)(*[count[i] if active[s][i] else 0 for i in range(6)])
        before = p.clone()
        check(L.svsr_adamw_step_segmented(
            C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(m.data_ptr()), C.c_void_p(v.data_ptr()),
            C.c_int64(n_decay), C.c_int64(n), C.c_float(lr), C.c_float(betas[0]), C.c_float(betas[1]), C.c_float(eps),
            C.c_float(wd), begin, steps, C.c_int(6), C.c_float(1.0), C.c_float(1.0), C.c_void_p(scratch.data_ptr()),
            C.c_void_p(0)), "adamw_seg")
        for i, (a, b) in enumerate(ranges):
            if not active[s][i]:
                assert torch.equal(p[a:b], before[a:b]), (s, i)  # bit-identical: not even decayed
    assert (p - ref).abs().max().item() < 2e-5
    assert ((p - ref).norm() / (ref - p0).norm()).item() < 1e-5


def _module(depth=2, **kw):
    import test_lrw_gpu as tl
    from syncvsr_b200.lightning import TransformerLightningModule

    meta = dict(B=2, S=88, A=4, G=2, V=320, depth=depth, seed_p=3, seed_x=77, extra_tokens=0)
    return tl._native(TransformerLightningModule, meta, **kw)


def test_fused_optimizer_skips_sublayers_dropped_by_layer_dropout(monkeypatch):
    """layer_dropout: the reference's skipped sublayers have `grad is None`, so AdamW neither decays nor moves them
    (SURVEY.md section 7). FusedAdamW reads the module's skip mask."""
    import random

    from syncvsr_b200.train import FusedAdamW

    m, P, (videos, tokens, labels, wm) = _module(depth=2, layer_dropout=0.5)
    opt = FusedAdamW(m, lr=1e-2, weight_decay=0.1)
    draws = iter([0.9, 0.1, 0.9, 0.9,   # step 1: sublayer 1 (layers.1 = first FFN) dropped
                  0.9, 0.9, 0.9, 0.9,   # step 2: nothing dropped
                  0.1, 0.9, 0.9, 0.1])  # step 3: sublayers 0 and 3 dropped
    monkeypatch.setattr(random, "random", lambda: next(draws))
    sd0 = {k: v.detach().clone() for k, v in m._param_views.items()}
    expect_skipped = [{1}, set(), {0, 3}]
    for s in range(3):
        before = {k: v.detach().clone() for k, v in m._param_views.items()}
        opt.zero_grad()
        out = m(videos.cuda(), tokens.cuda(), labels.cuda(), wm.cuda())
        out["loss_total"].backward()
        assert {i for i in range(4) if (m._last_skip >> i) & 1} == expect_skipped[s]
        opt.step()
        torch.cuda.synchronize()
        for k, v in m._param_views.items():
            sub = int(k.split(".")[2]) if k.startswith("encoder.layers.") else None
            if sub in expect_skipped[s]:
                assert torch.equal(v.detach(), before[k]), (s, k)
            else:
                assert not torch.equal(v.detach(), before[k]), (s, k)
    # per-range step counts (group 0 = always trained, group 1+i = sublayer i): 0, 1 and 3 took two steps, 2 three
    assert opt._group_steps == {0: 3, 1: 2, 2: 2, 3: 3, 4: 2}
    assert any(not torch.equal(sd0[k], v.detach()) for k, v in m._param_views.items())


def test_torch_optimizer_updates_reach_the_packed_weights():
    """configure_optimizers() returns torch.optim.AdamW, which writes the fp32 arena views in place and
    never calls mark_weights_updated(). The bf16 operand copies must still follow: after one optimizer step the module's
    forward equals the forward of a fresh module loaded from the updated state dict, and moves with the oracle."""
    from syncvsr_b200.lightning import TransformerLightningModule

    m, P, (videos, tokens, labels, wm) = _module(depth=2)
    v, t, l, w = videos.cuda(), tokens.cuda(), labels.cuda(), wm.cuda()
    (optimizer,), _ = m.configure_optimizers()
    for grp in optimizer.param_groups:
        grp["lr"] = 1e-3  # large enough that a stale weight copy shows in the loss
    losses = []
    for _ in range(3):
        optimizer.zero_grad(set_to_none=False)
        out = m(v, t, l, w)
        out["loss_total"].backward()
        optimizer.step()
        losses.append(float(out["loss_total"]))
    assert losses[2] < losses[0] - 1.0  # the trunk/encoder weights really train (stale copies would stall it)
    with torch.no_grad():
        now = float(m(v, t, l, w)["loss_total"])
    import test_lrw_gpu as tl

    fresh = TransformerLightningModule(tl.make_cfg(depth=2)).train()
    fresh.load_state_dict(m.state_dict(), strict=False)
    for (ka, a), (kb, b) in zip(sorted(m.named_buffers()), sorted(fresh.named_buffers())):
        assert ka == kb and torch.equal(a, b)
    with torch.no_grad():
        again = float(fresh(v, t, l, w)["loss_total"])
    assert now == pytest.approx(again, rel=1e-5)
    # the oracle taking the same three AdamW steps on its own (bf16-storage) gradients lands on the same loss
    Pq = {k: p.clone().requires_grad_("running" not in k) for k, p in P.items()}
    train_p = [p for k, p in Pq.items() if "running" not in k]
    oopt = torch.optim.AdamW([{"params": [p for p in train_p if p.ndim >= 2]},
                              {"params": [p for p in train_p if p.ndim < 2], "weight_decay": 0.0}],
                             lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01)
    for _ in range(3):
        oopt.zero_grad()
        o = O.lrw_forward(Pq, videos, tokens, labels, wm, depth=2, q=O.bf16_ste)
        o["loss_total"].backward()
        oopt.step()
        with torch.no_grad():
            for k, val in o["new_stats"].items():
                Pq[k].copy_(val)
    with torch.no_grad():
        o_now = float(O.lrw_forward(Pq, videos, tokens, labels, wm, depth=2, q=O.bf16_ste)["loss_total"])
    assert now == pytest.approx(o_now, rel=5e-2)
    assert abs(now - o_now) < 0.3 * abs(losses[0] - o_now)  # both moved far from the start, and together
