"""Developer probe (GPU box): where does the native LRS step first deviate from the oracle? Prints rel-L2 errors of
the residual stream after every Conformer sub-block, the heads and the losses; plus a tap-by-tap check of the
depthwise convolution. Not part of the product path."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import lrs_oracle as O  # noqa: E402
from oracle.lrw_oracle import bf16_ste  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def dwconv_probe():
    from syncvsr_b200 import ops
    B, T, C, K = 1, 12, 64, 7
    x = torch.zeros(B, T, C, device="cuda", dtype=torch.bfloat16)
    x[0, 5, :] = 1.0
    w = torch.arange(1, K + 1, device="cuda", dtype=torch.float32).repeat(C, 1).contiguous()
    y = ops.dwconv1d_fwd(x, w, None)
    print("dwconv impulse at t=5, w=1..7 -> y[:,0] =", y[0, :, 0].float().tolist())
    ref = F.conv1d(x.float().transpose(1, 2), w.view(C, 1, K), None, padding=3, groups=C).transpose(1, 2)
    print("torch                         ref[:,0] =", ref[0, :, 0].tolist())
    ref2 = F.conv1d(x.float().transpose(1, 2).contiguous(), w.view(C, 1, K), None, padding=3, groups=C).transpose(1, 2)
    print("torch contiguous input        ref[:,0] =", ref2[0, :, 0].tolist())
    xr = torch.randn(2, 40, 256, device="cuda").to(torch.bfloat16)
    wr = torch.randn(256, 31, device="cuda")
    y = ops.dwconv1d_fwd(xr, wr, None)
    r_cuda = F.conv1d(xr.float().transpose(1, 2), wr.view(256, 1, 31), None, padding=15, groups=256).transpose(1, 2)
    r_cpu = F.conv1d(xr.float().cpu().transpose(1, 2), wr.cpu().view(256, 1, 31), None, padding=15, groups=256).transpose(1, 2)
    print("random: native vs torch-cuda", rel(y, r_cuda), " native vs torch-cpu", rel(y, r_cpu), " torch-cuda vs torch-cpu",
          rel(r_cuda, r_cpu))


def model_probe():
    import test_lrs_gpu as tg
    from syncvsr_b200.e2e import E2E
    c = torch.load(ROOT / "tests/golden/lrs_small.pt")["meta"]
    for train in (True, False):
        m, P, inputs = tg._native(E2E, c, train=train)
        x, lengths, tokens, label = inputs
        with torch.no_grad():
            out = m(x.cuda(), lengths.cuda(), tokens.cuda(), label.cuda())
        q = bf16_ste
        ns = {}
        T = c["T"]
        mask = (torch.arange(T).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(-2)
        with torch.no_grad():
            feats = O.frontend(x, P, train, ns, q)
            print(f"--- train={train}")
            print("frontend", rel(m._named_tensor("frontend", (c["B"], T, 512)), feats))
            h = O.linear(feats, "encoder.embed.0", P, q) * (c["adim"] ** 0.5)
            print("embed_out", rel(m._named_tensor("embed_out", (c["B"], T, c["adim"])), h))
            pos = O.rel_pos_emb(T, c["adim"])
            for i in range(c["elayers"]):
                pre = f"encoder.encoders.{i}"
                steps = []
                h = h + 0.5 * O.ffn(O.layer_norm(h, pre + ".norm_ff_macaron", P), pre + ".feed_forward_macaron", P, q)
                steps.append(h)
                h = h + O.rel_mha(O.layer_norm(h, pre + ".norm_mha", P), pos, mask, pre + ".self_attn", P, c["heads"], q)
                steps.append(h)
                h = h + O.conv_module(O.layer_norm(h, pre + ".norm_conv", P), pre + ".conv_module", P, train, ns, q)
                steps.append(h)
                h = h + 0.5 * O.ffn(O.layer_norm(h, pre + ".norm_ff", P), pre + ".feed_forward", P, q)
                steps.append(h)
                h = O.layer_norm(h, pre + ".norm_final", P)
                steps.append(h)
                for k, s in enumerate(steps):
                    print(f"layer{i}.x{k + 1}", rel(m._named_tensor(f"layer{i}.x{k + 1}", (c["B"], T, c["adim"])), s))
            o = tg._oracle(c, P, inputs, q=q, train=train)
            print("encoder_out", rel(m.encoder_out(), o["encoder_out"]))
            print("logits_audio", rel(m.logits_audio().flatten(), o["logits_audio"].flatten()))
            print("ctc_logits", rel(m.ctc_logits(), o["ctc_logits"]))
            print("pred", rel(m.pred()[:, : o["pred"].shape[1]], o["pred"]))
            for got, key in zip(out[:4], ("loss", "loss_ctc", "loss_att", "loss_audio")):
                print(key, float(got), float(o[key]))
            print("acc", float(out[4]), o["acc"])


if __name__ == "__main__" and "grads" not in sys.argv:
    dwconv_probe()
    model_probe()


def grad_probe():
    import test_lrs_gpu as tg
    from syncvsr_b200.e2e import E2E
    c = torch.load(ROOT / "tests/golden/lrs_small.pt")["meta"]
    m, P, inputs = tg._native(E2E, c)
    x, lengths, tokens, label = inputs
    out = m(x.cuda(), lengths.cuda(), tokens.cuda(), label.cuda())
    out[0].backward()
    for name, q in (("bf16_ste", bf16_ste), ("fp32", None)):
        Pq = {k: v.clone().requires_grad_("running" not in k) for k, v in P.items()}
        o = tg._oracle(c, Pq, inputs, q=q)
        o["loss"].backward()
        print(f"--- gradients vs oracle[{name}] (rel-L2, cosine)")
        for k, p in m._param_views.items():
            ref = Pq[k].grad
            if ref is None or float(ref.norm()) < 1e-6:
                print(f"{k:70s} zero-ref native-norm {float(p.grad.norm()):.3e}")
            else:
                print(f"{k:70s} {rel(p.grad, ref):.4f} {tg.cosine(p.grad, ref):.4f}")


if __name__ == "__main__" and "grads" in sys.argv:
    grad_probe()
