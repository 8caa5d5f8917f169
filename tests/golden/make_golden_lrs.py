"""Generates tests/golden/lrs_*.pt from the UNMODIFIED reference `E2E` module (run in the build container only:
`python tests/golden/make_golden_lrs.py`). E2E.forward
(/root/reference/LRS/video/espnet/nets/pytorch_backend/e2e_asr_transformer.py:186-227) is executed on seeded synthetic
inputs with the seeded parameters of oracle.lrs_oracle.make_params; the audio head is enabled as described in
SURVEY.md section 8c (codec attributes set by hand, pre-made tokens). Fixtures are small: scalars, slices, norms."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import lrs_oracle as O  # noqa: E402
from oracle import ref_loader as rl  # noqa: E402

OUT = Path(__file__).resolve().parent

CASES = {
    # small widths (fast): 2 conformer blocks, 1 decoder block, ragged lengths
    "lrs_small": dict(B=3, T=12, S=88, adim=256, heads=4, eunits=512, elayers=2, dlayers=1, odim=300, A=2, G=2, V=64,
                      seed_p=0, seed_x=5, extra_tokens=3),
    # reference widths of lrs2.yaml (adim 768, 12 heads, eunits 3072, odim 5049, wav2vec2 codec 2x2x640), shallow
    "lrs_c3_w768": dict(B=2, T=16, S=88, adim=768, heads=12, eunits=3072, elayers=2, dlayers=2, odim=5049, A=2, G=2,
                        V=640, seed_p=1, seed_x=6, extra_tokens=0),
}


def model_kwargs(c):
    return dict(adim=c["adim"], heads=c["heads"], eunits=c["eunits"], elayers=c["elayers"], dlayers=c["dlayers"],
                odim=c["odim"], n_audio=c["A"] * c["G"] * c["V"])


def run_case(name: str, c: dict):
    P = O.make_params(c["seed_p"], **model_kwargs(c))
    x, lengths, tokens, label = O.make_inputs(c["seed_x"], c["B"], c["T"], S=c["S"], A=c["A"], G=c["G"], V=c["V"],
                                              odim=c["odim"], extra_tokens=c["extra_tokens"])
    m = rl.build_reference_lrs(P, odim=c["odim"], audio_alignment=c["A"], audio_vocab_size=c["V"],
                               n_audio=c["A"] * c["G"] * c["V"], tokens=tokens, adim=c["adim"], aheads=c["heads"],
                               eunits=c["eunits"], elayers=c["elayers"], ddim=c["adim"], dheads=c["heads"],
                               dunits=c["eunits"], dlayers=c["dlayers"]).train()
    cap = {}
    m.encoder.register_forward_hook(lambda mod, i, o: cap.__setitem__("encoder_out", o[0].detach()))
    m.encoder.frontend.register_forward_hook(lambda mod, i, o: cap.__setitem__("frontend", o.detach()))
    m.audio_classifier.register_forward_hook(lambda mod, i, o: cap.__setitem__("logits_audio", o.detach()))
    m.ctc.ctc_lo.register_forward_hook(lambda mod, i, o: cap.__setitem__("ctc_logits", o.detach()))
    m.decoder.register_forward_hook(lambda mod, i, o: cap.__setitem__("pred", o[0].detach()))
    loss, loss_ctc, loss_att, loss_audio, acc = m(x, lengths, torch.zeros(c["B"], 1, 1), label)
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    sd = m.state_dict()
    e0 = "encoder.encoders.0"
    fx = {
        "meta": dict(c, torch=str(torch.__version__)),
        "metrics": dict(loss=float(loss), loss_ctc=float(loss_ctc), loss_att=float(loss_att),
                        loss_audio=float(loss_audio), acc=float(acc)),
        "lengths": lengths.clone(), "label": label.clone(),
        "frontend_b0": cap["frontend"][0].clone(),
        "encoder_out_t3": cap["encoder_out"][:, 3, :].clone(),
        "encoder_out_last": cap["encoder_out"][:, -1, :].clone(),
        "encoder_out_abs": cap["encoder_out"].double().abs().sum().item(),
        "logits_audio_t2": cap["logits_audio"][:, 2, :].clone(),
        "ctc_logits_t1": cap["ctc_logits"][:, 1, :].clone(),
        "pred_l0": cap["pred"][:, 0, :].clone(),
        "audio_targets": tokens[:, : c["T"] * c["A"]].flatten().clone(),
        "grad_norms": {k: g.double().norm().item() for k, g in grads.items()},
        "grad_embed_b": grads["encoder.embed.0.bias"].clone(),
        "grad_pos_bias_u": grads[e0 + ".self_attn.pos_bias_u"].clone(),
        "grad_pos_bias_v": grads[e0 + ".self_attn.pos_bias_v"].clone(),
        "grad_linear_pos_slice": grads[e0 + ".self_attn.linear_pos.weight"][:4].clone(),
        "grad_dw_slice": grads[e0 + ".conv_module.depthwise_conv.weight"][:8].clone(),
        "grad_bn1d_w": grads[e0 + ".conv_module.norm.weight"].clone(),
        "grad_norm_final_w": grads[e0 + ".norm_final.weight"].clone(),
        "grad_ctc_b_slice": grads["ctc.ctc_lo.bias"][:64].clone(),
        "grad_dec_embed_norm": grads["decoder.embed.0.weight"].double().norm().item(),
        "grad_audio_bias": grads["audio_classifier.bias"].clone(),
        "grad_stem_w": grads["encoder.frontend.frontend3D.0.weight"].clone(),
        "running_mean_bn1d": sd[e0 + ".conv_module.norm.running_mean"].clone(),
        "running_var_bn1d": sd[e0 + ".conv_module.norm.running_var"].clone(),
        "running_var_stem": sd["encoder.frontend.frontend3D.1.running_var"].clone(),
    }
    torch.save(fx, OUT / f"{name}.pt")
    print(name, fx["metrics"], "size", (OUT / f"{name}.pt").stat().st_size)


if __name__ == "__main__":
    torch.manual_seed(0)
    for name, c in CASES.items():
        run_case(name, c)
