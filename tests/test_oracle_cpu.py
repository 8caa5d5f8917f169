"""CPU tests that pin the oracle (oracle/lrw_oracle.py) against golden vectors produced by the reference's own
TransformerLightningModule.forward (tests/golden/make_golden.py) and, where /root/reference exists, against the
reference module itself."""
import pytest
import torch

from oracle import lrw_oracle as O
from oracle import ref_loader as rl

CASES = ["lrw_c1_vq", "lrw_c1_a2", "lrw_96_d2", "lrw_wb_d2", "lrw_hf_d2"]  # word boundary (dim 513); HF BertModel encoder


def _run_oracle(meta, need_grad=False, word_mask=None):
    wb = bool(meta.get("wb", False))
    P = O.make_params(meta["seed_p"], depth=meta["depth"], n_audio=meta["A"] * meta["G"] * meta["V"],
                      dim=513 if wb else 512)
    hf = meta.get("hf")
    if hf:
        P = O.make_hf_params(P, hf, seed=meta["seed_p"] + 100)
    if need_grad:
        for k, v in P.items():
            if "running_" not in k:
                v.requires_grad_(True)
    videos, tokens, labels, wm = O.make_inputs(meta["seed_x"], meta["B"], S=meta["S"], A=meta["A"], V=meta["V"],
                                               extra_tokens=meta["extra_tokens"])
    if wb:
        wm = word_mask
    out = O.lrw_forward(P, videos, tokens, labels, wm, depth=meta["depth"], audio_alignment=meta["A"],
                        vq_groups=meta["G"], audio_vocab_size=meta["V"], use_wb=wb, hf_bert=hf)
    return P, (videos, tokens, labels, wm), out


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name, golden_dir):
    fx = torch.load(golden_dir / f"{name}.pt")
    meta = fx["meta"]
    P, inputs, out = _run_oracle(meta, need_grad=(name != "lrw_c1_vq"), word_mask=fx.get("word_mask"))
    tol = dict(rtol=2e-4, atol=2e-4)
    for k, v in fx["metrics"].items():
        assert float(out[k]) == pytest.approx(v, rel=1e-5, abs=1e-6), k
    torch.testing.assert_close(out["last_hidden_state"][:, 0, :], fx["last_hidden_state_cls"], **tol)
    torch.testing.assert_close(out["last_hidden_state"][:, 7, :], fx["last_hidden_state_t7"], **tol)
    assert out["last_hidden_state"].double().abs().sum().item() == pytest.approx(fx["last_hidden_state_abs"], rel=1e-5)
    la = out["logits_audio"].reshape(meta["B"], 29, -1)
    torch.testing.assert_close(la[:, 3, :], fx["logits_audio_t3"], **tol)
    torch.testing.assert_close(out["logits_category"], fx["logits_category"], **tol)
    torch.testing.assert_close(out["inputs_embeds"].flatten(0, 1)[:2, :512], fx["inputs_embeds_t0"], **tol)
    if meta.get("wb"):  # the word-boundary channel is the mask itself
        assert torch.equal(out["inputs_embeds"][..., 512], fx["word_mask"])
    # integer path: bit exact
    assert torch.equal(O.audio_targets(inputs[1], 29, meta["A"]), fx["audio_targets"])
    # BN buffers
    torch.testing.assert_close(out["new_stats"]["stem3d.1.running_mean"], fx["running_mean_stem"], **tol)
    torch.testing.assert_close(out["new_stats"]["resnet.layer4.1.bn2.running_var"], fx["running_var_l4"], **tol)
    if name != "lrw_c1_vq":
        out["loss_total"].backward()
        def rel(a, b):
            return ((a - b).norm() / b.norm()).item()

        # conv weight gradients behind train-mode BN are sums with heavy cancellation (dY has zero channel mean):
        # fp32 summation order alone moves them by ~2e-3 between two CPU runs of the same math.
        assert rel(P["stem3d.0.weight"].grad, fx["grad_stem_w"]) < 1e-2
        assert rel(P["resnet.layer1.0.conv1.weight"].grad[:4], fx["grad_l1_conv1_slice"]) < 1e-2
        assert rel(P["cls_token"].grad, fx["grad_cls_token"]) < 2e-3
        assert rel(P["audio_projection.bias"].grad, fx["grad_audio_bias"]) < 2e-3
        assert rel(P["resnet.layer4.1.bn2.weight"].grad, fx["grad_l4_bn2_w"]) < 2e-3
        g0 = "encoder.encoder.layer.0.attention.output.LayerNorm.weight" if meta.get("hf") else "encoder.layers.0.0.g"
        assert rel(P[g0].grad, fx["grad_enc0_g"]) < 2e-3
        for k, n in fx["grad_norms"].items():
            if n < 1e-6:  # mathematically zero (a key bias under softmax): only rounding noise on either side
                assert P[k].grad.double().norm().item() < 1e-5, k
            else:
                assert P[k].grad.double().norm().item() == pytest.approx(n, rel=2e-3), k
        assert sorted(k for k, v in P.items() if v.requires_grad and v.grad is None) == []
        unused = ["resnet.bn1.bias", "resnet.bn1.weight", "resnet.conv1.weight", "resnet.fc.bias", "resnet.fc.weight"]
        if meta.get("hf"):  # BertModel members that forward(inputs_embeds=...).last_hidden_state never touches
            unused = ["encoder.embeddings.word_embeddings.weight", "encoder.pooler.dense.bias",
                      "encoder.pooler.dense.weight"] + unused
        assert fx["unused_params"] == unused


def test_audio_target_indexing_is_the_reference_layout():
    """Row r = ((b*T+t)*A+a)*G+g of logits.reshape(-1,V) must pair with audio_tokens[b, t*A+a, g]
    (lightning.py:147,170-171; README.md:47-53)."""
    B, T, A, G = 3, 29, 4, 2
    tokens = torch.arange(B * (T * A + 7) * G).reshape(B, T * A + 7, G)
    flat = O.audio_targets(tokens, T, A)
    assert flat.numel() == B * T * A * G
    for b, t, a, g in [(0, 0, 0, 0), (1, 5, 3, 1), (2, 28, 3, 1), (2, 0, 1, 0)]:
        r = ((b * T + t) * A + a) * G + g
        assert flat[r].item() == tokens[b, t * A + a, g].item()


def test_bf16_hook_changes_little():
    fxmeta = dict(seed_p=2, seed_x=1236, depth=2, A=4, G=2, V=320, B=2, S=88, extra_tokens=0)
    P = O.make_params(fxmeta["seed_p"], depth=2)
    videos, tokens, labels, wm = O.make_inputs(fxmeta["seed_x"], 2)
    a = O.lrw_forward(P, videos, tokens, labels, wm, depth=2)
    b = O.lrw_forward(P, videos, tokens, labels, wm, depth=2, q=O.bf16_ste)
    assert float(b["loss_total"]) == pytest.approx(float(a["loss_total"]), rel=5e-3)
    rel = (a["last_hidden_state"] - b["last_hidden_state"]).norm() / a["last_hidden_state"].norm()
    assert rel < 0.1


@pytest.mark.skipif(not rl.reference_available(), reason="reference tree not mounted (GPU box)")
def test_oracle_matches_live_reference_module():
    ref = rl.load_reference_lrw()
    m = ref.TransformerLightningModule(rl.reference_config(depth=2)).train()
    P = O.make_params(7, depth=2)
    m.load_state_dict(P, strict=False)
    videos, tokens, labels, wm = O.make_inputs(99, 2)
    r = m(videos, tokens, labels, wm)
    o = O.lrw_forward({k: v.clone() for k, v in P.items()}, videos, tokens, labels, wm, depth=2)
    for k in ("loss_total", "loss_category", "loss_audio", "accuracy_top1", "accuracy_top5"):
        assert float(o[k]) == pytest.approx(float(r[k]), rel=1e-6, abs=1e-7), k


def test_eager_baseline_module_equals_oracle():
    """bench.py's `--impl eager` / `gpu_baseline` arm (oracle/eager_module.py: the reference's module graph from stock
    torch.nn modules) computes the same function as the pinned oracle on the same state dict."""
    from oracle.eager_module import EagerLRW

    torch.manual_seed(0)
    P = O.make_params(21, depth=2)
    videos, tokens, labels, wm = O.make_inputs(22, 2)
    m = EagerLRW(depth=2).train()
    m.load_oracle_params(P)
    out = m(videos, tokens, labels, wm)
    ref = O.lrw_forward(P, videos, tokens, labels, wm, depth=2)
    for k in ("loss_total", "loss_category", "loss_audio"):
        assert float(out[k]) == pytest.approx(float(ref[k]), rel=1e-5), k
    assert torch.allclose(out["last_hidden_state"], ref["last_hidden_state"], rtol=1e-4, atol=1e-4)
    assert torch.allclose(out["logits_audio"], ref["logits_audio"], rtol=1e-4, atol=1e-4)
