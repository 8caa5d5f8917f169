"""Generates tests/golden/lrw_*.pt from the UNMODIFIED reference module (run in the build container only:
`python tests/golden/make_golden.py`). The reference's TransformerLightningModule.forward
(/root/reference/LRW/video/src/lightning.py:133-191) is executed on seeded synthetic inputs with the seeded
parameters of oracle.lrw_oracle.make_params; intermediate tensors are captured with forward hooks.
Fixtures are small: scalars, a few rows/slices and checksums -- parameters are regenerated from the seed."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import lrw_oracle as O  # noqa: E402
from oracle import ref_loader as rl  # noqa: E402

OUT = Path(__file__).resolve().parent


HF_CFG = dict(hidden_size=512, num_hidden_layers=2, num_attention_heads=8, intermediate_size=1024, vocab_size=1,
              max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


def wb_mask(B: int, T: int = 29, seed: int = 7):
    """[B, T] 0/1 word-boundary indicator: a centred run of U{5..20} frames (data.py:58-64)."""
    g = torch.Generator().manual_seed(seed)
    wm = torch.zeros(B, T)
    for b in range(B):
        n = int(torch.randint(5, 21, (1,), generator=g))
        s0 = (T - n) // 2
        wm[b, s0 : s0 + n] = 1.0
    return wm


def run_case(name: str, B: int, S: int, A: int, V: int, depth: int, seed_p: int, seed_x: int, extra_tokens: int,
             wb: bool = False, hf: dict | None = None):
    ref = rl.load_reference_lrw()
    cfg = rl.reference_config(depth=depth, use_wb=wb, encoder_type="huggingface" if hf else "x-transformers")
    if hf:  # the shipped yaml carries x-transformers keys only: BertConfig(**cfg.model.bert) needs these
        for k, v in hf.items():
            cfg["model"]["bert"][k] = v
    m = ref.TransformerLightningModule(cfg).train()
    G = 2
    if (A, V) != (4, 320):  # BASELINE.json config 1 (alignment=2, vocab=320): the reference derives these from the
        m.audio_alignment, m.audio_vocab_size = A, V  # codec path string, so set the attributes it reads at run time
        m.audio_projection = torch.nn.Linear(512, A * G * V)
    P = O.make_params(seed_p, depth=depth, n_audio=A * G * V, dim=513 if wb else 512)
    if hf:
        P = O.make_hf_params(P, hf, seed=seed_p + 100)
    m.load_state_dict(P, strict=False)
    videos, tokens, labels, wm = O.make_inputs(seed_x, B, S=S, A=A, V=V, extra_tokens=extra_tokens)
    if wb:
        wm = wb_mask(B)

    cap = {}
    m.encoder.register_forward_hook(lambda mod, i, o: cap.__setitem__(
        "last_hidden_state", (o.last_hidden_state if hasattr(o, "last_hidden_state") else o).detach()))
    m.audio_projection.register_forward_hook(lambda mod, i, o: cap.__setitem__("logits_audio", o.detach()))
    m.category_classifier.register_forward_hook(lambda mod, i, o: cap.__setitem__("logits_category", o.detach()))
    m.resnet.layer4.register_forward_hook(lambda mod, i, o: cap.__setitem__("layer4", o.detach()))
    m.stem3d.register_forward_hook(lambda mod, i, o: cap.__setitem__("stem", o.detach()))
    out = m(videos, tokens, labels, wm)
    out["loss_total"].backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    sd = m.state_dict()

    fx = {
        "meta": dict(B=B, S=S, A=A, G=G, V=V, depth=depth, seed_p=seed_p, seed_x=seed_x, extra_tokens=extra_tokens,
                     wb=wb, hf=hf, torch=str(torch.__version__)),
        "word_mask": wm.clone(),
        "metrics": {k: float(v) for k, v in out.items()},
        "last_hidden_state_cls": cap["last_hidden_state"][:, 0, :].clone(),
        "last_hidden_state_t7": cap["last_hidden_state"][:, 7, :].clone(),
        "last_hidden_state_sum": cap["last_hidden_state"].double().sum().item(),
        "last_hidden_state_abs": cap["last_hidden_state"].double().abs().sum().item(),
        "logits_audio_t3": cap["logits_audio"][:, 3, :].clone(),
        "logits_audio_abs": cap["logits_audio"].double().abs().sum().item(),
        "logits_category": cap["logits_category"].clone(),
        "inputs_embeds_t0": cap["layer4"].mean((2, 3))[:2].clone(),
        "stem_slice": cap["stem"][0, :, 3, 5, :].clone(),
        "audio_targets": tokens[:, : 29 * A].flatten().clone(),
        "grad_norms": {k: g.double().norm().item() for k, g in grads.items()},
        "grad_stem_w": grads["stem3d.0.weight"].clone(),
        "grad_cls_token": grads["cls_token"].clone(),
        "grad_audio_bias": grads["audio_projection.bias"].clone(),
        "grad_l4_bn2_w": grads["resnet.layer4.1.bn2.weight"].clone(),
        "grad_enc0_g": grads["encoder.encoder.layer.0.attention.output.LayerNorm.weight" if hf else "encoder.layers.0.0.g"].clone(),
        "grad_l1_conv1_slice": grads["resnet.layer1.0.conv1.weight"][:4].clone(),
        "unused_params": sorted(k for k, p in m.named_parameters() if p.grad is None),
        "running_mean_stem": sd["stem3d.1.running_mean"].clone(),
        "running_var_l4": sd["resnet.layer4.1.bn2.running_var"].clone(),
    }
    torch.save(fx, OUT / f"{name}.pt")
    print(name, fx["metrics"], "size", (OUT / f"{name}.pt").stat().st_size)


if __name__ == "__main__":
    torch.manual_seed(0)
    # C1 reference-native codec constants (vq: alignment 4, groups 2, vocab 320), B=2, 88x88
    run_case("lrw_c1_vq", B=2, S=88, A=4, V=320, depth=12, seed_p=0, seed_x=1234, extra_tokens=0)
    # C1 as BASELINE.json states it (alignment=2, vq_groups=2, vocab=320); token tensor longer than T*A (ragged tail)
    run_case("lrw_c1_a2", B=2, S=88, A=2, V=320, depth=12, seed_p=1, seed_x=1235, extra_tokens=5)
    # reference-native 96x96 crop, shallow encoder (fast CPU case)
    run_case("lrw_96_d2", B=3, S=96, A=4, V=320, depth=2, seed_p=2, seed_x=1236, extra_tokens=3)
    # shipped word-boundary configuration (data.use_word_boundary: true -> hidden dim 513), shallow encoder
    run_case("lrw_wb_d2", B=3, S=88, A=4, V=320, depth=2, seed_p=4, seed_x=1237, extra_tokens=0, wb=True)
    # `model.bert.type: huggingface` (lightning.py:90-92): transformers.BertModel as the encoder, hidden 512
    run_case("lrw_hf_d2", B=3, S=88, A=4, V=320, depth=2, seed_p=5, seed_x=1238, extra_tokens=0, hf=HF_CFG)
