"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path (syncvsr_b200/*, bench.py's GPU arm).

CPU restatement, in plain fp32 PyTorch functional ops, of the reference LRS sentence-level hot path
`E2E.forward` (/root/reference/LRS/video/espnet/nets/pytorch_backend/e2e_asr_transformer.py:186-227) operating on a
state dict with the reference's own key names (`encoder.frontend.frontend3D.*`, `encoder.encoders.N.*`, `ctc.ctc_lo.*`,
`decoder.*`, `audio_classifier.*`). Each function cites the reference lines it follows (paths relative to
LRS/video/espnet/nets/pytorch_backend/).

Pinning: every module on this path is IN-TREE in the reference (vendored espnet), so tests/test_lrs_oracle_cpu.py pins
this file (a) against the UNMODIFIED reference `E2E` imported through oracle/ref_loader.load_reference_lrs() when
/root/reference is present and (b) against tests/golden/lrs_*.pt fixtures generated from that module by
tests/golden/make_golden_lrs.py. The audio tokens are an input (the frozen wav2vec quantiser of
e2e_asr_transformer.py:167-180 is off the gradient path and needs a network download).

`q` is an optional quantisation hook applied where the CUDA path stores a tensor in bf16 (q=None: exact fp32).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .lrw_oracle import QFn, Tensor, _q, batchnorm

FRONT = "encoder.frontend"


def swish(x: Tensor) -> Tensor:
    """transformer/convolution.py:78-83"""
    return x * torch.sigmoid(x)


# ----------------------------------------------------------------------------------------------------------------
# Conv3dResNet: backbones/conv3d_extractor.py:19-48 (frontend3D = Conv3d -> BN3d -> Swish -> MaxPool3d; trunk =
# ResNet(BasicBlock,[2,2,2,2], relu_type="swish") backbones/modules/resnet.py:45-177 incl. AdaptiveAvgPool2d(1))
# ----------------------------------------------------------------------------------------------------------------
def frontend(xs: Tensor, P, train: bool, new_stats, q: QFn = None, cap: Optional[dict] = None) -> Tensor:
    """xs [B, T, 1, H, W] -> [B, T, 512]"""
    B = xs.shape[0]
    x = xs.transpose(1, 2)  # conv3d_extractor.py:41
    x = F.conv3d(_q(q, x), _q(q, P[FRONT + ".frontend3D.0.weight"]), None, (1, 2, 2), (2, 3, 3))
    x = _q(q, x)
    x = swish(batchnorm(x, FRONT + ".frontend3D.1", P, train, new_stats))
    x = _q(q, F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1)))
    T = x.shape[2]
    h = x.transpose(1, 2).reshape(B * T, 64, x.shape[3], x.shape[4])  # threeD_to_2D_tensor :13-16
    if cap is not None:
        cap["stem_out"] = h.detach()
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for b, st in ((0, stride), (1, 1)):
            pre = f"{FRONT}.trunk.layer{li}.{b}"
            out = _q(q, F.conv2d(h, _q(q, P[pre + ".conv1.weight"]), None, st, 1))
            out = _q(q, swish(batchnorm(out, pre + ".bn1", P, train, new_stats)))
            out = _q(q, F.conv2d(out, _q(q, P[pre + ".conv2.weight"]), None, 1, 1))
            out = batchnorm(out, pre + ".bn2", P, train, new_stats)
            if (pre + ".downsample.0.weight") in P:  # resnet.py:26-42
                sc = _q(q, F.conv2d(h, _q(q, P[pre + ".downsample.0.weight"]), None, st, 0))
                sc = batchnorm(sc, pre + ".downsample.1", P, train, new_stats)
            else:
                sc = h
            h = _q(q, swish(out + sc))  # resnet.py:104-105
    return h.mean((2, 3)).view(B, T, 512)  # avgpool + view (resnet.py:175-177; conv3d_extractor.py:48)


# ----------------------------------------------------------------------------------------------------------------
# RelPositionalEncoding: transformer/embedding.py:153-217. Row k of pos_emb [1, 2T-1, d] encodes relative position
# T-1-k (row 0 = +(T-1), row T-1 = 0, last row = -(T-1)).
# ----------------------------------------------------------------------------------------------------------------
def rel_pos_emb(T: int, d: int) -> Tensor:
    pos = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)  # +(T-1) ... -(T-1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(2 * T - 1, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def abs_pos_emb(L: int, d: int) -> Tensor:
    """PositionalEncoding: transformer/embedding.py:33-88"""
    pos = torch.arange(0, L, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(L, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def layer_norm(x: Tensor, prefix: str, P) -> Tensor:
    """transformer/layer_norm.py:12-33 (eps = 1e-12)"""
    return F.layer_norm(x, (x.shape[-1],), P[prefix + ".weight"], P[prefix + ".bias"], 1e-12)


def linear(x: Tensor, prefix: str, P, q: QFn = None, bias: bool = True) -> Tensor:
    return F.linear(_q(q, x), _q(q, P[prefix + ".weight"]), P[prefix + ".bias"] if bias else None)


def ffn(x: Tensor, prefix: str, P, q: QFn = None) -> Tensor:
    """transformer/positionwise_feed_forward.py:28-30 (dropout = 0 in parity runs)"""
    return linear(_q(q, torch.relu(linear(x, prefix + ".w_1", P, q))), prefix + ".w_2", P, q)


def masked_softmax_av(scores: Tensor, v: Tensor, mask: Optional[Tensor]) -> Tensor:
    """transformer/attention.py:59-88: mask [B, 1|T1, T2] bool; fill -1e10, softmax, re-mask to 0, @ V."""
    if mask is not None:
        m = mask.unsqueeze(1).eq(0)
        scores = scores.masked_fill(m, -1e10)
        attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
    else:
        attn = torch.softmax(scores, dim=-1)
    x = torch.matmul(attn, v)  # [B, h, T1, d_k]
    return x.transpose(1, 2).reshape(x.shape[0], x.shape[2], -1)


def rel_mha(x: Tensor, pos_emb: Tensor, mask: Optional[Tensor], prefix: str, P, heads: int, q: QFn = None) -> Tensor:
    """RelPositionMultiHeadedAttention.forward: transformer/attention.py:238-278 with rel_shift (216-236) written as the
    index map bd[i, j] = raw[i, j - i + T - 1] (verified against the reference in tests/test_lrs_oracle_cpu.py)."""
    B, T, D = x.shape
    dk = D // heads
    qq = _q(q, linear(x, prefix + ".linear_q", P, q)).view(B, T, heads, dk)
    k = _q(q, linear(x, prefix + ".linear_k", P, q)).view(B, T, heads, dk).transpose(1, 2)
    v = _q(q, linear(x, prefix + ".linear_v", P, q)).view(B, T, heads, dk).transpose(1, 2)
    p = _q(q, linear(pos_emb, prefix + ".linear_pos", P, q, bias=False)).view(1, 2 * T - 1, heads, dk).transpose(1, 2)
    qu = (qq + P[prefix + ".pos_bias_u"]).transpose(1, 2)
    qv = (qq + P[prefix + ".pos_bias_v"]).transpose(1, 2)
    ac = torch.matmul(qu, k.transpose(-2, -1))
    raw = torch.matmul(qv, p.transpose(-2, -1))  # [B, h, T, 2T-1]
    idx = (torch.arange(T).view(1, T) - torch.arange(T).view(T, 1) + T - 1).to(x.device)  # [i, j] -> j - i + T - 1
    bd = raw.gather(-1, idx.view(1, 1, T, T).expand(B, heads, T, T))
    scores = (ac + bd) / math.sqrt(dk)
    ctx = masked_softmax_av(scores, v, mask)
    return linear(_q(q, ctx), prefix + ".linear_out", P, q)


def mha(xq: Tensor, xkv: Tensor, mask: Optional[Tensor], prefix: str, P, heads: int, q: QFn = None) -> Tensor:
    """MultiHeadedAttention.forward: transformer/attention.py:38-57,90-108"""
    B, T1, D = xq.shape
    T2 = xkv.shape[1]
    dk = D // heads
    qq = _q(q, linear(xq, prefix + ".linear_q", P, q)).view(B, T1, heads, dk).transpose(1, 2)
    k = _q(q, linear(xkv, prefix + ".linear_k", P, q)).view(B, T2, heads, dk).transpose(1, 2)
    v = _q(q, linear(xkv, prefix + ".linear_v", P, q)).view(B, T2, heads, dk).transpose(1, 2)
    scores = torch.matmul(qq, k.transpose(-2, -1)) / math.sqrt(dk)
    ctx = masked_softmax_av(scores, v, mask)
    return linear(_q(q, ctx), prefix + ".linear_out", P, q)


def conv_module(x: Tensor, prefix: str, P, train: bool, new_stats, q: QFn = None) -> Tensor:
    """ConvolutionModule.forward: transformer/convolution.py:56-75. x [B, T, C]; no padding mask is applied."""
    C = x.shape[-1]
    K = P[prefix + ".depthwise_conv.weight"].shape[-1]
    h = x.transpose(1, 2)
    h = F.conv1d(_q(q, h), _q(q, P[prefix + ".pointwise_cov1.weight"]), P[prefix + ".pointwise_cov1.bias"])
    h = _q(q, F.glu(_q(q, h), dim=1))
    h = F.conv1d(h, P[prefix + ".depthwise_conv.weight"], P[prefix + ".depthwise_conv.bias"], padding=(K - 1) // 2,
                 groups=C)
    h = _q(q, h)
    h = _q(q, swish(batchnorm(h, prefix + ".norm", P, train, new_stats)))
    h = F.conv1d(h, _q(q, P[prefix + ".pointwise_cov2.weight"]), P[prefix + ".pointwise_cov2.bias"])
    return h.transpose(1, 2)


def encoder_layer(x: Tensor, pos_emb: Tensor, mask, prefix: str, P, heads: int, train: bool, new_stats,
                  q: QFn = None) -> Tensor:
    """EncoderLayer.forward: transformer/encoder_layer.py:76-150 (macaron, normalize_before, no concat_after)"""
    x = x + 0.5 * ffn(layer_norm(x, prefix + ".norm_ff_macaron", P), prefix + ".feed_forward_macaron", P, q)
    x = x + rel_mha(layer_norm(x, prefix + ".norm_mha", P), pos_emb, mask, prefix + ".self_attn", P, heads, q)
    x = x + conv_module(layer_norm(x, prefix + ".norm_conv", P), prefix + ".conv_module", P, train, new_stats, q)
    x = x + 0.5 * ffn(layer_norm(x, prefix + ".norm_ff", P), prefix + ".feed_forward", P, q)
    return layer_norm(x, prefix + ".norm_final", P)


def encoder(xs: Tensor, mask: Optional[Tensor], P, elayers: int, heads: int, train: bool, new_stats, q: QFn = None,
            cap: Optional[dict] = None) -> Tensor:
    """Encoder.forward: transformer/encoder.py:257-289"""
    feats = frontend(xs, P, train, new_stats, q, cap)
    if cap is not None:
        cap["frontend"] = feats.detach()
    x = linear(feats, "encoder.embed.0", P, q)
    D = x.shape[-1]
    x = x * math.sqrt(D)  # embedding.py:212
    pos_emb = rel_pos_emb(x.shape[1], D).to(x.device)
    for i in range(elayers):
        x = encoder_layer(x, pos_emb, mask, f"encoder.encoders.{i}", P, heads, train, new_stats, q)
    return layer_norm(x, "encoder.after_norm", P)


# ----------------------------------------------------------------------------------------------------------------
# heads
# ----------------------------------------------------------------------------------------------------------------
def audio_loss(x: Tensor, audio_tokens: Tensor, P, A: int, V: int, q: QFn = None):
    """e2e_asr_transformer.py:195-201: no padding mask -- padded frames are scored too."""
    tok = audio_tokens[:, : x.shape[1] * A]
    logits = linear(x, "audio_classifier", P, q).float().unflatten(2, (-1, V))
    return F.cross_entropy(logits.flatten(0, 2), tok.flatten()), logits


def ctc_loss(x: Tensor, lengths: Tensor, label: Tensor, P, q: QFn = None):
    """CTC.forward: ctc.py:83-151 with loss_fn :64-73 (builtin CTCLoss, reduction sum, zero_infinity, / batch)"""
    ys = [y[y != -1] for y in label]
    ys_hat = linear(x.float(), "ctc.ctc_lo", P, q).transpose(0, 1)
    olens = torch.tensor([len(s) for s in ys], dtype=torch.long)
    lp = ys_hat.log_softmax(2)
    loss = F.ctc_loss(lp, torch.cat(ys), lengths.long(), olens, blank=0, reduction="sum", zero_infinity=True)
    return loss / lp.shape[1], ys_hat.transpose(0, 1)


def add_sos_eos(label: Tensor, sos: int, eos: int, ignore_id: int = -1):
    """transformer/add_sos_eos.py:12-31"""
    ys = [y[y != ignore_id] for y in label]
    L = max(len(y) for y in ys) + 1
    ys_in = label.new_full((len(ys), L), eos)
    ys_out = label.new_full((len(ys), L), ignore_id)
    for i, y in enumerate(ys):
        ys_in[i, 0] = sos
        ys_in[i, 1 : len(y) + 1] = y
        ys_out[i, : len(y)] = y
        ys_out[i, len(y)] = eos
    return ys_in, ys_out


def decoder(ys_in: Tensor, memory: Tensor, memory_mask: Tensor, P, dlayers: int, heads: int, q: QFn = None) -> Tensor:
    """Decoder.forward: transformer/decoder.py:122-151; DecoderLayer.forward decoder_layer.py:58-121;
    target_mask mask.py:41-51 (ys_in never contains the ignore id, so the mask is the causal one)."""
    B, L = ys_in.shape
    D = P["decoder.embed.0.weight"].shape[1]
    tgt_mask = (ys_in != -1).unsqueeze(-2) & torch.tril(torch.ones(L, L, dtype=torch.bool)).unsqueeze(0)
    x = F.embedding(ys_in, P["decoder.embed.0.weight"]) * math.sqrt(D) + abs_pos_emb(L, D)
    for i in range(dlayers):
        pre = f"decoder.decoders.{i}"
        y = layer_norm(x, pre + ".norm1", P)
        x = x + mha(y, y, tgt_mask, pre + ".self_attn", P, heads, q)
        x = x + mha(layer_norm(x, pre + ".norm2", P), memory, memory_mask, pre + ".src_attn", P, heads, q)
        x = x + ffn(layer_norm(x, pre + ".norm3", P), pre + ".feed_forward", P, q)
    x = layer_norm(x, "decoder.after_norm", P)
    return linear(x, "decoder.output_layer", P, q)


def label_smoothing_loss(pred: Tensor, target: Tensor, smoothing: float, ignore_id: int = -1) -> Tensor:
    """LabelSmoothingLoss.forward: transformer/label_smoothing_loss.py:41-63 (normalize_length=False: / batch)"""
    B, L, V = pred.shape
    x = pred.reshape(-1, V)
    t = target.reshape(-1)
    ignore = t == ignore_id
    true_dist = torch.full_like(x, smoothing / (V - 1))
    true_dist.scatter_(1, t.masked_fill(ignore, 0).unsqueeze(1), 1.0 - smoothing)
    kl = F.kl_div(torch.log_softmax(x, dim=1), true_dist, reduction="none")
    return kl.masked_fill(ignore.unsqueeze(1), 0).sum() / B


def th_accuracy(pred: Tensor, target: Tensor, ignore_id: int = -1) -> float:
    """nets_utils.py:303-322"""
    B, L, V = pred.shape
    hyp = pred.argmax(2)
    m = target != ignore_id
    return float((hyp[m] == target[m]).sum()) / float(m.sum())


def lrs_forward(P: Dict[str, Tensor], x: Tensor, lengths: Tensor, audio_tokens: Optional[Tensor], label: Tensor, *,
                elayers: int = 12, dlayers: int = 6, heads: int = 12, odim: int = 5049, audio_alignment: int = 2,
                audio_vocab_size: int = 640, audio_weight: float = 10.0, mtlalpha: float = 0.1,
                lsm_weight: float = 0.1, train: bool = True, q: QFn = None, cap: Optional[dict] = None):
    """E2E.forward: e2e_asr_transformer.py:186-227. x [B, T, 1, H, W], lengths [B], label [B, Lmax] padded with -1."""
    new_stats: Dict[str, Tensor] = {}
    T = x.shape[1]
    mask = (torch.arange(T).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(-2)  # make_non_pad_mask, nets_utils.py:183
    h = encoder(x, mask, P, elayers, heads, train, new_stats, q, cap)
    out = {"encoder_out": h}
    if audio_tokens is not None:
        loss_audio, logits_audio = audio_loss(h, audio_tokens, P, audio_alignment, audio_vocab_size, q)
        out["logits_audio"] = logits_audio
    else:
        loss_audio = None
    loss_ctc, ys_hat = ctc_loss(h, lengths, label, P, q)
    out["ctc_logits"] = ys_hat
    sos = eos = odim - 1
    ys_in, ys_out = add_sos_eos(label, sos, eos)
    pred = decoder(ys_in, h, mask, P, dlayers, heads, q)
    loss_att = label_smoothing_loss(pred.float(), ys_out, lsm_weight)
    loss = mtlalpha * loss_ctc + (1 - mtlalpha) * loss_att
    if loss_audio is not None:
        loss = loss + loss_audio * audio_weight
    out.update(loss=loss, loss_ctc=loss_ctc, loss_att=loss_att, loss_audio=loss_audio,
               acc=th_accuracy(pred, ys_out), pred=pred, ys_in=ys_in, ys_out=ys_out, new_stats=new_stats)
    return out


# ----------------------------------------------------------------------------------------------------------------
# deterministic synthetic parameters / inputs shared by golden generation, CPU tests and GPU parity tests
# ----------------------------------------------------------------------------------------------------------------
def make_params(seed: int = 0, adim: int = 768, heads: int = 12, eunits: int = 3072, elayers: int = 12,
                dlayers: int = 6, odim: int = 5049, n_audio: int = 2560, kernel: int = 31) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    P: Dict[str, Tensor] = {}

    def bn(prefix, c):
        P[prefix + ".weight"] = 1.0 + 0.1 * randn(c)
        P[prefix + ".bias"] = 0.1 * randn(c)
        P[prefix + ".running_mean"] = torch.zeros(c)
        P[prefix + ".running_var"] = torch.ones(c)

    def lin(prefix, n, k, bias=True, std=None):
        P[prefix + ".weight"] = randn(n, k, std=std if std is not None else (3 * k) ** -0.5)
        if bias:
            P[prefix + ".bias"] = 0.02 * randn(n)

    def ln(prefix, c):
        P[prefix + ".weight"] = 1.0 + 0.05 * randn(c)
        P[prefix + ".bias"] = 0.02 * randn(c)

    cin = 64
    for li, c in ((1, 64), (2, 128), (3, 256), (4, 512)):
        for b in (0, 1):
            pre = f"{FRONT}.trunk.layer{li}.{b}"
            P[pre + ".conv1.weight"] = randn(c, cin if b == 0 else c, 3, 3, std=math.sqrt(2.0 / (c * 9)))
            bn(pre + ".bn1", c)
            P[pre + ".conv2.weight"] = randn(c, c, 3, 3, std=math.sqrt(2.0 / (c * 9)))
            bn(pre + ".bn2", c)
            if b == 0 and li > 1:
                P[pre + ".downsample.0.weight"] = randn(c, cin, 1, 1, std=math.sqrt(2.0 / c))
                bn(pre + ".downsample.1", c)
        cin = c
    P[FRONT + ".frontend3D.0.weight"] = randn(64, 1, 5, 7, 7, std=math.sqrt(2.0 / (64 * 245)) * 4)
    bn(FRONT + ".frontend3D.1", 64)
    lin("encoder.embed.0", adim, 512)
    dk = adim // heads
    for i in range(elayers):
        pre = f"encoder.encoders.{i}"
        P[pre + ".self_attn.pos_bias_u"] = randn(heads, dk, std=0.1)
        P[pre + ".self_attn.pos_bias_v"] = randn(heads, dk, std=0.1)
        for c in ("q", "k", "v", "out"):
            lin(f"{pre}.self_attn.linear_{c}", adim, adim)
        lin(pre + ".self_attn.linear_pos", adim, adim, bias=False)
        lin(pre + ".feed_forward.w_1", eunits, adim)
        lin(pre + ".feed_forward.w_2", adim, eunits)
        P[pre + ".conv_module.pointwise_cov1.weight"] = randn(2 * adim, adim, 1, std=(3 * adim) ** -0.5)
        P[pre + ".conv_module.pointwise_cov1.bias"] = 0.02 * randn(2 * adim)
        P[pre + ".conv_module.depthwise_conv.weight"] = randn(adim, 1, kernel, std=kernel ** -0.5)
        P[pre + ".conv_module.depthwise_conv.bias"] = 0.02 * randn(adim)
        bn(pre + ".conv_module.norm", adim)
        P[pre + ".conv_module.pointwise_cov2.weight"] = randn(adim, adim, 1, std=(3 * adim) ** -0.5)
        P[pre + ".conv_module.pointwise_cov2.bias"] = 0.02 * randn(adim)
        ln(pre + ".norm_ff", adim)
        ln(pre + ".norm_mha", adim)
        lin(pre + ".feed_forward_macaron.w_1", eunits, adim)
        lin(pre + ".feed_forward_macaron.w_2", adim, eunits)
        ln(pre + ".norm_ff_macaron", adim)
        ln(pre + ".norm_conv", adim)
        ln(pre + ".norm_final", adim)
    ln("encoder.after_norm", adim)
    P["decoder.embed.0.weight"] = randn(odim, adim, std=adim ** -0.5)
    for i in range(dlayers):
        pre = f"decoder.decoders.{i}"
        for a in ("self_attn", "src_attn"):
            for c in ("q", "k", "v", "out"):
                lin(f"{pre}.{a}.linear_{c}", adim, adim)
        lin(pre + ".feed_forward.w_1", eunits, adim)
        lin(pre + ".feed_forward.w_2", adim, eunits)
        ln(pre + ".norm1", adim)
        ln(pre + ".norm2", adim)
        ln(pre + ".norm3", adim)
    ln("decoder.after_norm", adim)
    lin("decoder.output_layer", odim, adim)
    lin("ctc.ctc_lo", odim, adim)
    lin("audio_classifier", n_audio, adim)
    return P


def make_inputs(seed: int, B: int, T: int, S: int = 88, A: int = 2, G: int = 2, V: int = 640, odim: int = 5049,
                min_len: Optional[int] = None, lab_min: int = 3, lab_max: int = 8, extra_tokens: int = 0):
    """SURVEY.md section 8(d) C3/C4 inputs: N(0,1) clips, lengths U{T/2..T} with one full-length sample (frames beyond a
    clip's length are zero, like collate_pad, datamodule/data_module.py:12-43), labels U{1..odim-2} padded with -1."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, 1, S, S, generator=g)
    lo = min_len if min_len is not None else max(T // 2, 1)
    lengths = torch.randint(lo, T + 1, (B,), generator=g)
    lengths[0] = T
    for b in range(B):
        x[b, lengths[b]:] = 0
    tokens = torch.randint(0, V, (B, T * A + extra_tokens, G), generator=g)
    lab_len = torch.randint(lab_min, lab_max + 1, (B,), generator=g)
    label = torch.full((B, int(lab_len.max())), -1, dtype=torch.long)
    for b in range(B):
        label[b, : lab_len[b]] = torch.randint(1, odim - 1, (int(lab_len[b]),), generator=g)
    return x, lengths, tokens, label
