"""CPU tests that pin oracle/lrs_oracle.py against golden vectors produced by the reference's own E2E.forward
(tests/golden/make_golden_lrs.py) and, where /root/reference exists, against the reference module itself."""
import math

import pytest
import torch

from oracle import lrs_oracle as O
from oracle import ref_loader as rl

CASES = ["lrs_small", "lrs_c3_w768"]


def _kwargs(c):
    return dict(adim=c["adim"], heads=c["heads"], eunits=c["eunits"], elayers=c["elayers"], dlayers=c["dlayers"],
                odim=c["odim"], n_audio=c["A"] * c["G"] * c["V"])


def _run_oracle(c, need_grad):
    P = O.make_params(c["seed_p"], **_kwargs(c))
    if need_grad:
        for k, v in P.items():
            if "running_" not in k:
                v.requires_grad_(True)
    x, lengths, tokens, label = O.make_inputs(c["seed_x"], c["B"], c["T"], S=c["S"], A=c["A"], G=c["G"], V=c["V"],
                                              odim=c["odim"], extra_tokens=c["extra_tokens"])
    out = O.lrs_forward(P, x, lengths, tokens, label, elayers=c["elayers"], dlayers=c["dlayers"], heads=c["heads"],
                        odim=c["odim"], audio_alignment=c["A"], audio_vocab_size=c["V"])
    return P, (x, lengths, tokens, label), out


@pytest.mark.parametrize("name", CASES)
def test_lrs_oracle_matches_reference_golden(name, golden_dir):
    fx = torch.load(golden_dir / f"{name}.pt")
    c = fx["meta"]
    need_grad = name == "lrs_small"
    P, inputs, out = _run_oracle(c, need_grad)
    tol = dict(rtol=2e-4, atol=2e-4)
    assert torch.equal(inputs[1], fx["lengths"]) and torch.equal(inputs[3], fx["label"])
    for k, v in fx["metrics"].items():
        assert float(out[k]) == pytest.approx(v, rel=2e-5, abs=1e-6), k
    torch.testing.assert_close(out["encoder_out"][:, 3, :], fx["encoder_out_t3"], **tol)
    torch.testing.assert_close(out["encoder_out"][:, -1, :], fx["encoder_out_last"], **tol)
    assert out["encoder_out"].double().abs().sum().item() == pytest.approx(fx["encoder_out_abs"], rel=1e-5)
    torch.testing.assert_close(out["logits_audio"].flatten(2)[:, 2, :], fx["logits_audio_t2"], **tol)
    torch.testing.assert_close(out["ctc_logits"][:, 1, :], fx["ctc_logits_t1"], **tol)
    torch.testing.assert_close(out["pred"][:, 0, :], fx["pred_l0"], **tol)
    assert torch.equal(inputs[2][:, : c["T"] * c["A"]].flatten(), fx["audio_targets"])  # integer path: bit exact
    e0 = "encoder.encoders.0"
    torch.testing.assert_close(out["new_stats"][e0 + ".conv_module.norm.running_mean"], fx["running_mean_bn1d"], **tol)
    torch.testing.assert_close(out["new_stats"][e0 + ".conv_module.norm.running_var"], fx["running_var_bn1d"], **tol)
    torch.testing.assert_close(out["new_stats"]["encoder.frontend.frontend3D.1.running_var"], fx["running_var_stem"],
                               **tol)
    if need_grad:
        out["loss"].backward()

        def rel(a, b):
            return ((a - b).norm() / b.norm()).item()

        assert rel(P["encoder.embed.0.bias"].grad, fx["grad_embed_b"]) < 2e-3
        assert rel(P[e0 + ".self_attn.pos_bias_u"].grad, fx["grad_pos_bias_u"]) < 2e-3
        assert rel(P[e0 + ".self_attn.pos_bias_v"].grad, fx["grad_pos_bias_v"]) < 2e-3
        assert rel(P[e0 + ".self_attn.linear_pos.weight"].grad[:4], fx["grad_linear_pos_slice"]) < 2e-3
        assert rel(P[e0 + ".conv_module.depthwise_conv.weight"].grad[:8], fx["grad_dw_slice"]) < 2e-3
        assert rel(P[e0 + ".conv_module.norm.weight"].grad, fx["grad_bn1d_w"]) < 2e-3
        assert rel(P[e0 + ".norm_final.weight"].grad, fx["grad_norm_final_w"]) < 2e-3
        assert rel(P["ctc.ctc_lo.bias"].grad[:64], fx["grad_ctc_b_slice"]) < 2e-3
        assert rel(P["audio_classifier.bias"].grad, fx["grad_audio_bias"]) < 2e-3
        assert rel(P["encoder.frontend.frontend3D.0.weight"].grad, fx["grad_stem_w"]) < 1e-2
        for k, n in fx["grad_norms"].items():
            if n < 1e-6:  # mathematically zero gradients (key bias under softmax, conv bias before BatchNorm)
                assert P[k].grad.double().norm().item() < 1e-5, k
            else:
                assert P[k].grad.double().norm().item() == pytest.approx(n, rel=5e-3), k


def test_rel_shift_index_map_equals_reference_view_trick():
    """attention.py:216-236 rel_shift == bd[i, j] = raw[i, j - i + T - 1]."""
    T = 7
    raw = torch.randn(2, 3, T, 2 * T - 1)
    zero_pad = torch.zeros((*raw.size()[:3], 1))
    xp = torch.cat([zero_pad, raw], dim=-1).view(2, 3, 2 * T, T)
    ref = xp[:, :, 1:].view_as(raw)[:, :, :, :T]
    idx = torch.arange(T).view(1, T) - torch.arange(T).view(T, 1) + T - 1
    mine = raw.gather(-1, idx.view(1, 1, T, T).expand(2, 3, T, T))
    assert torch.equal(ref, mine)


def test_rel_pos_emb_row_order():
    pe = O.rel_pos_emb(5, 8)[0]
    assert torch.allclose(pe[4, 0::2], torch.zeros(4)) and torch.allclose(pe[4, 1::2], torch.ones(4))  # position 0
    assert pe[0, 0] == pytest.approx(math.sin(4.0)) and pe[8, 0] == pytest.approx(math.sin(-4.0))


def test_fully_masked_query_row_and_padding_are_scored_like_the_reference():
    """audio CE has no padding mask (e2e_asr_transformer.py:198-201): changing a padded frame's tokens changes the loss."""
    c = dict(B=2, T=8, S=88, adim=256, heads=4, eunits=512, elayers=1, dlayers=1, odim=120, A=2, G=2, V=32, seed_p=3,
             seed_x=9, extra_tokens=0)
    P, (x, lengths, tokens, label), out = _run_oracle(c, False)
    assert int(lengths[1]) < c["T"]
    t2 = tokens.clone()
    t2[1, -1, 0] = (t2[1, -1, 0] + 1) % c["V"]
    out2 = O.lrs_forward(P, x, lengths, t2, label, elayers=1, dlayers=1, heads=4, odim=120, audio_alignment=2,
                         audio_vocab_size=32)
    assert float(out2["loss_audio"]) != float(out["loss_audio"])
    assert float(out2["loss_ctc"]) == float(out["loss_ctc"])


@pytest.mark.skipif(not rl.reference_lrs_available(), reason="reference tree not mounted (GPU box)")
def test_lrs_oracle_matches_live_reference_module():
    c = dict(B=2, T=10, S=88, adim=256, heads=4, eunits=512, elayers=1, dlayers=1, odim=150, A=2, G=2, V=32, seed_p=4,
             seed_x=11, extra_tokens=1)
    P, (x, lengths, tokens, label), out = _run_oracle(c, False)
    m = rl.build_reference_lrs(P, odim=150, audio_alignment=2, audio_vocab_size=32, n_audio=128, tokens=tokens, adim=256,
                               aheads=4, eunits=512, elayers=1, ddim=256, dheads=4, dunits=512, dlayers=1).train()
    loss, loss_ctc, loss_att, loss_audio, acc = m(x, lengths, torch.zeros(2, 1, 1), label)
    for k, v in (("loss", loss), ("loss_ctc", loss_ctc), ("loss_att", loss_att), ("loss_audio", loss_audio)):
        assert float(out[k]) == pytest.approx(float(v), rel=2e-6, abs=1e-6), k
    assert out["acc"] == pytest.approx(acc)
