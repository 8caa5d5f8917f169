"""Developer probe: native LRW step vs the oracle, tensor by tensor (forward) and gradient by gradient."""
from __future__ import annotations

import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import lrw_oracle as O  # noqa: E402
from oracle.ref_loader import AttrDict  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cfg_dict(depth):
    return AttrDict.wrap({
        "model": {"resnet": "resnet18", "wav2vec": {"path": "./vq-wav2vec_kmeans.pt"},
                  "bert": {"type": "x-transformers", "dim": 512, "depth": depth, "heads": 8, "emb_dropout": 0.0,
                           "attn_dropout": 0.0, "layer_dropout": 0.0, "ff_dropout": 0.0, "num_labels": 500}},
        "optim": {"lambda_audio": 10.0, "optimizer": {"lr": 1e-4, "betas": [0.9, 0.999], "eps": 1e-6, "weight_decay": 0.01},
                  "scheduler": {"name": "cosine", "num_warmup_steps": 15000, "num_training_steps": 270000}},
        "train": {"label_smoothing": 0.0, "use_cutmix": False},
        "data": {"use_word_boundary": False},
    })


def main(depth=2, B=2, S=88):
    from syncvsr_b200.lightning import TransformerLightningModule

    torch.manual_seed(0)
    m = TransformerLightningModule(cfg_dict(depth)).train()
    P = O.make_params(3, depth=depth)
    missing, unexpected = m.load_state_dict(P, strict=False)
    print("missing", len(missing), [k for k in missing if "num_batches" not in k and "resnet.conv1" not in k
                                    and "resnet.bn1" not in k and "resnet.fc" not in k], "unexpected", unexpected)
    videos, tokens, labels, wm = O.make_inputs(77, B, S=S)
    t0 = time.time()
    out = m(videos.cuda(), tokens.cuda(), labels.cuda(), wm.cuda())
    torch.cuda.synchronize()
    print("native forward", time.time() - t0, {k: float(v) for k, v in out.items()})

    Pq = {k: v.clone().requires_grad_("running" not in k) for k, v in P.items()}
    cap = {}
    o = O.lrw_forward(Pq, videos, tokens, labels, wm, depth=depth, q=O.bf16_ste, cap=cap)
    o32 = O.lrw_forward(P, videos, tokens, labels, wm, depth=depth)
    print("oracle(bf16 hooks)", {k: float(o[k]) for k in ("loss_total", "loss_category", "loss_audio")})
    print("oracle(fp32)      ", {k: float(o32[k]) for k in ("loss_total", "loss_category", "loss_audio")})

    T = 29
    H0 = cap["stem_conv"].shape[-1]
    nat = m._named_tensor("stem_conv", (B, T, H0, H0, 64)).float().cpu()
    print("stem_conv rel", rel(nat, cap["stem_conv"].permute(0, 2, 3, 4, 1)))
    H1 = cap["stem_out"].shape[-1]
    nat = m._named_tensor("stem_out", (B * T, H1, H1, 64)).float().cpu()
    print("stem_out  rel", rel(nat, cap["stem_out"].permute(0, 2, 3, 1)))
    for bi in range(8):
        ref = cap[f"block{bi}.out"].permute(0, 2, 3, 1)
        nat = m._named_tensor(f"block{bi}.out", tuple(ref.shape)).float().cpu()
        print(f"block{bi}.out rel", rel(nat, ref))
    emb = m._named_tensor("inputs_embeds", (B, T + 1, 512))[:, 1:].cpu()
    print("inputs_embeds rel", rel(emb, o["inputs_embeds"].detach()), "vs fp32", rel(emb, o32["inputs_embeds"]))
    print("last_hidden rel", rel(m.last_hidden_state(), o["last_hidden_state"].detach()), "vs fp32",
          rel(m.last_hidden_state(), o32["last_hidden_state"]))
    print("logits_audio rel", rel(m.logits_audio(), o["logits_audio"].detach()), "vs fp32",
          rel(m.logits_audio(), o32["logits_audio"]))
    print("logits_cat rel", rel(m.logits_category(), o["logits_category"].detach()))

    # backward
    out["loss_total"].backward()
    torch.cuda.synchronize()
    o["loss_total"].backward()
    worst = []
    for k, p in m._param_views.items():
        r = rel(p.grad, Pq[k].grad)
        worst.append((r, k, float(Pq[k].grad.norm())))
    worst.sort(reverse=True)
    print("grad rel errors (worst 25):")
    for r, k, n in worst[:25]:
        print(f"   {r:.3e}  {k}  |g|={n:.3e}")
    print("grad rel median", sorted(w[0] for w in worst)[len(worst) // 2])
    print("all grads in arena order:")
    for k, p in m._param_views.items():
        print(f"   {rel(p.grad, Pq[k].grad):.3e}  {k}  |g|={float(Pq[k].grad.norm()):.3e} |g_native|={float(p.grad.norm()):.3e}")
    sd = m.state_dict()
    print("running_mean stem rel", rel(sd["stem3d.1.running_mean"], o["new_stats"]["stem3d.1.running_mean"]),
          "running_var l4", rel(sd["resnet.layer4.1.bn2.running_var"], o["new_stats"]["resnet.layer4.1.bn2.running_var"]))


if __name__ == "__main__":
    main(depth=int(sys.argv[1]) if len(sys.argv) > 1 else 2, B=int(sys.argv[2]) if len(sys.argv) > 2 else 2)
