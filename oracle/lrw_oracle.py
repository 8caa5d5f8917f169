"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path (syncvsr_b200/*, bench.py's GPU arm).

CPU restatement, in plain fp32 PyTorch functional ops, of the reference LRW hot path
`TransformerLightningModule.forward` (/root/reference/LRW/video/src/lightning.py:133-191) operating on a
state dict with the reference's own key names. Each function cites the reference lines it follows.

Pinning: tests/test_oracle_cpu.py checks this file (a) against the UNMODIFIED reference module imported
through oracle/ref_loader.py when /root/reference is present, and (b) against tests/golden/*.pt fixtures generated
by tests/golden/make_golden.py from that same reference module. The encoder inside the reference run is the
x-transformers restatement of oracle/xt_encoder.py (the real package is absent: PARITY UNPINNED at that boundary,
see that file's header); stem, ResNet trunk, heads and losses are pinned against the reference's own code.

`q` is an optional quantisation hook applied at every point where the CUDA path stores a tensor in bf16
(and to every GEMM/conv operand); with q=None this is the exact fp32 algorithm.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
QFn = Optional[Callable[[Tensor], Tensor]]


def bf16_ste(x: Tensor) -> Tensor:
    """Round to bf16 (straight-through gradient): models bf16 storage / tensor-core operands."""
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


def _q(q: QFn, x: Tensor) -> Tensor:
    return x if q is None else q(x)


# ----------------------------------------------------------------------------------------------------------------
# BatchNorm (train mode): torch.nn.BatchNorm{2,3}d defaults eps=1e-5, momentum=0.1 (lightning.py:51; torchvision
# resnet BasicBlock bn1/bn2). Normalises with the biased batch variance; running_var gets the unbiased one.
# ----------------------------------------------------------------------------------------------------------------
def batchnorm(x: Tensor, prefix: str, P: Dict[str, Tensor], train: bool, new_stats: Dict[str, Tensor],
              eps: float = 1e-5, momentum: float = 0.1) -> Tensor:
    dims = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    w, b = P[prefix + ".weight"], P[prefix + ".bias"]
    if train:
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            new_stats[prefix + ".running_mean"] = (1 - momentum) * P[prefix + ".running_mean"] + momentum * mean
            new_stats[prefix + ".running_var"] = (1 - momentum) * P[prefix + ".running_var"] + momentum * var * (
                n / max(n - 1, 1))
    else:
        mean, var = P[prefix + ".running_mean"], P[prefix + ".running_var"]
    return (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + eps) * w.view(shape) + b.view(shape)


# ----------------------------------------------------------------------------------------------------------------
# stem3d: Conv3d(1,64,(5,7,7),(1,2,2),(2,3,3),bias=False) -> BatchNorm3d -> GELU(erf) -> MaxPool3d((1,3,3),(1,2,2),(0,1,1))
# (lightning.py:49-54)
# ----------------------------------------------------------------------------------------------------------------
def stem3d(videos: Tensor, P, train: bool, new_stats, q: QFn = None, cap: Optional[dict] = None) -> Tensor:
    x = F.conv3d(_q(q, videos), _q(q, P["stem3d.0.weight"]), None, (1, 2, 2), (2, 3, 3))
    x = _q(q, x)
    if cap is not None:
        cap["stem_conv"] = x.detach()
    x = batchnorm(x, "stem3d.1", P, train, new_stats)
    x = F.gelu(x)
    x = F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    return _q(q, x)


# ----------------------------------------------------------------------------------------------------------------
# ResNet-18 trunk layer1..layer4 (timm.create_model("resnet18") == torchvision resnet18 modules; call sites
# lightning.py:55,114-117): BasicBlock = relu(bn2(conv2(relu(bn1(conv1(x))))) + shortcut)
# ----------------------------------------------------------------------------------------------------------------
def basic_block(x: Tensor, prefix: str, P, stride: int, train: bool, new_stats, q: QFn = None) -> Tensor:
    out = F.conv2d(x, _q(q, P[prefix + ".conv1.weight"]), None, stride, 1)
    out = _q(q, out)
    out = _q(q, F.relu(batchnorm(out, prefix + ".bn1", P, train, new_stats)))
    out = F.conv2d(out, _q(q, P[prefix + ".conv2.weight"]), None, 1, 1)
    out = _q(q, out)
    out = batchnorm(out, prefix + ".bn2", P, train, new_stats)
    if (prefix + ".downsample.0.weight") in P:
        sc = F.conv2d(x, _q(q, P[prefix + ".downsample.0.weight"]), None, stride, 0)
        sc = _q(q, sc)
        sc = batchnorm(sc, prefix + ".downsample.1", P, train, new_stats)
    else:
        sc = x
    return _q(q, F.relu(out + sc))


def forward_videos(videos: Tensor, P, train: bool, new_stats, q: QFn = None, cap: Optional[dict] = None) -> Tensor:
    """lightning.py:112-119: stem -> transpose(1,2).flatten(0,1) -> layer1..4 -> mean((2,3)) -> unflatten."""
    B = videos.shape[0]
    h = stem3d(videos, P, train, new_stats, q, cap).transpose(1, 2).flatten(0, 1)
    if cap is not None:
        cap["stem_out"] = h.detach()
    bi = 0
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for b, st in ((0, stride), (1, 1)):
            h = basic_block(h, f"resnet.layer{li}.{b}", P, st, train, new_stats, q)
            if cap is not None:
                cap[f"block{bi}.out"] = h.detach()
            bi += 1
    return h.mean((2, 3)).unflatten(0, (B, -1))


# ----------------------------------------------------------------------------------------------------------------
# x_transformers.Encoder 1.9.2 as called at lightning.py:95-105,158 (restated; see oracle/xt_encoder.py header)
# ----------------------------------------------------------------------------------------------------------------
def rmsnorm(x: Tensor, g: Tensor, eps: float = 1e-8) -> Tensor:
    norm = torch.norm(x, dim=-1, keepdim=True) * (x.shape[-1] ** -0.5)
    return x / norm.clamp(min=eps) * g


def rotary_table(n: int, rot_dim: int = 32) -> Tensor:
    inv_freq = 1.0 / (10000 ** (torch.arange(0, rot_dim, 2).float() / rot_dim))
    f = torch.einsum("i,j->ij", torch.arange(n).float(), inv_freq)
    return torch.cat((f, f), dim=-1)


def _rotary(t: Tensor, freqs: Tensor) -> Tensor:
    rot = freqs.shape[-1]
    tl, tr = t[..., :rot], t[..., rot:]
    x1, x2 = tl.chunk(2, dim=-1)
    tl = tl * freqs.cos() + torch.cat((-x2, x1), dim=-1) * freqs.sin()
    return torch.cat((tl, tr), dim=-1)


def encoder(x: Tensor, P, depth: int, heads: int, q: QFn = None, skip: Optional[set] = None) -> Tensor:
    """x: [B, n, D]. `skip` = set of sublayer indices dropped by layer_dropout on this step (host RNG in the reference)."""
    B, n, D = x.shape
    dh = 64
    freqs = rotary_table(n).to(x.device)
    for i in range(2 * depth):
        if skip and i in skip:
            continue
        pre = f"encoder.layers.{i}"
        y = _q(q, rmsnorm(x, P[pre + ".0.g"]))
        if i % 2 == 0:
            qq, kk, vv = (
                _q(q, F.linear(y, _q(q, P[f"{pre}.1.to_{c}.weight"]))).view(B, n, heads, dh).transpose(1, 2)
                for c in "qkv"
            )
            qq, kk, vv = _rotary(qq, freqs), _rotary(kk, freqs), _rotary(vv, freqs)
            dots = torch.einsum("bhid,bhjd->bhij", qq, kk) * dh ** -0.5
            attn = F.softmax(dots, dim=-1, dtype=torch.float32)
            o = torch.einsum("bhij,bhjd->bhid", attn, vv).transpose(1, 2).reshape(B, n, heads * dh)
            y = F.linear(_q(q, o), _q(q, P[pre + ".1.to_out.weight"]))
        else:
            hcat = _q(q, F.linear(y, _q(q, P[pre + ".1.ff.0.proj.weight"]), P[pre + ".1.ff.0.proj.bias"]))
            val, gate = hcat.chunk(2, dim=-1)
            u = _q(q, val * F.gelu(gate))
            y = F.linear(u, _q(q, P[pre + ".1.ff.3.weight"]), P[pre + ".1.ff.3.bias"])
        x = y + x
    return x


def hf_bert_encoder(x: Tensor, P: Dict[str, Tensor], cfg: dict, train: bool) -> Tensor:
    """BertModel(BertConfig(**cfg))(inputs_embeds=x, output_hidden_states=True).last_hidden_state with the module's
    parameters taken from (and sharing autograd with) P["encoder.<name>"]."""
    from torch.func import functional_call
    from transformers import BertConfig, BertModel

    m = BertModel(BertConfig(**cfg))
    m.train(train)
    params = {k[len("encoder."):]: v for k, v in P.items() if k.startswith("encoder.")}
    for k, v in m.state_dict().items():  # members forward() never touches (word embeddings, pooler) keep their init
        params.setdefault(k, v)
    return functional_call(m, params, kwargs=dict(inputs_embeds=x, output_hidden_states=True)).last_hidden_state


def make_hf_params(P: Dict[str, Tensor], cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Replaces the x-transformers encoder entries of a make_params() dict by seeded BertModel parameters."""
    g = torch.Generator().manual_seed(seed)
    D, I, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
    out = {k: v for k, v in P.items() if not k.startswith("encoder.")}

    def randn(*shape, std=0.05):
        return torch.randn(*shape, generator=g) * std

    def ln(prefix):
        out[prefix + ".weight"] = 1.0 + 0.05 * randn(D, std=1.0)
        out[prefix + ".bias"] = 0.02 * randn(D, std=1.0)

    e = "encoder.embeddings."
    out[e + "position_embeddings.weight"] = randn(cfg["max_position_embeddings"], D, std=0.3)
    out[e + "token_type_embeddings.weight"] = randn(2, D, std=0.3)
    ln(e + "LayerNorm")
    for i in range(L):
        pre = f"encoder.encoder.layer.{i}."
        for name in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
            out[pre + name + ".weight"] = randn(D, D, std=(3 * D) ** -0.5 * 2)
            out[pre + name + ".bias"] = 0.02 * randn(D, std=1.0)
        ln(pre + "attention.output.LayerNorm")
        out[pre + "intermediate.dense.weight"] = randn(I, D, std=(3 * D) ** -0.5 * 2)
        out[pre + "intermediate.dense.bias"] = 0.02 * randn(I, std=1.0)
        out[pre + "output.dense.weight"] = randn(D, I, std=(3 * I) ** -0.5 * 2)
        out[pre + "output.dense.bias"] = 0.02 * randn(D, std=1.0)
        ln(pre + "output.LayerNorm")
    return out


# ----------------------------------------------------------------------------------------------------------------
# audio target indexing (integer, bit-exact): lightning.py:147,170-171 == README.md:47-53
#   tokens[:, :T*A] -> flatten; row r of logits_audio.reshape(-1, V) is (b, t, c) with c = a*G + g and
#   r = ((b*T + t)*A + a)*G + g, whose target is audio_tokens[b, t*A + a, g].
# ----------------------------------------------------------------------------------------------------------------
def audio_targets(audio_tokens: Tensor, T: int, A: int) -> Tensor:
    return audio_tokens[:, : T * A].flatten()


def lrw_forward(P: Dict[str, Tensor], videos: Tensor, audio_tokens: Tensor, labels: Tensor, word_mask: Tensor, *,
                depth: int = 12, heads: int = 8, audio_alignment: int = 4, vq_groups: int = 2,
                audio_vocab_size: int = 320, lambda_audio: float = 10.0, label_smoothing: float = 0.0,
                use_wb: bool = False, train: bool = True, q: QFn = None, skip: Optional[set] = None,
                cap: Optional[dict] = None, hf_bert: Optional[dict] = None):
    """lightning.py:133-191. Returns the reference's metric dict plus last_hidden_state / logits and new BN buffers."""
    new_stats: Dict[str, Tensor] = {}
    emb = forward_videos(videos, P, train, new_stats, q, cap)
    if use_wb:
        emb = torch.cat((emb, word_mask.unsqueeze(-1)), dim=-1)
    B, T, D = emb.shape
    tok = audio_targets(audio_tokens, T, audio_alignment)
    x = torch.cat((P["cls_token"].expand(B, -1, -1), emb), dim=1)
    if hf_bert is not None:
        # lightning.py:90-92,152-156 (`type: huggingface`): the encoder IS the third-party transformers.BertModel, which is
        # importable here, so the oracle runs the library itself on the reference-named parameters ("encoder.*")
        last = hf_bert_encoder(x, P, hf_bert, train)
    else:
        last = encoder(x, P, depth, heads, q, skip)

    lq = _q(q, last)
    logits_category = F.linear(lq[:, 0, :], _q(q, P["category_classifier.weight"]), P["category_classifier.bias"]).float()
    loss_category = F.cross_entropy(logits_category, labels, label_smoothing=label_smoothing)
    logits_audio = F.linear(lq[:, 1:, :], _q(q, P["audio_projection.weight"]), P["audio_projection.bias"]).float()
    logits_audio = logits_audio.reshape(B, T, audio_alignment * vq_groups, audio_vocab_size)
    loss_audio = F.cross_entropy(logits_audio.reshape(-1, audio_vocab_size), tok)
    loss_total = loss_category + loss_audio * lambda_audio

    hard = labels.argmax(dim=-1) if labels.dim() == 2 else labels
    corrects = logits_category.topk(5, dim=1)[1] == hard.unsqueeze(1)
    return {
        "loss_total": loss_total,
        "loss_category": loss_category,
        "loss_audio": loss_audio,
        "accuracy_top1": corrects[:, 0].float().mean(),
        "accuracy_top5": corrects.float().amax(1).mean(),
        "last_hidden_state": last,
        "logits_audio": logits_audio,
        "logits_category": logits_category,
        "inputs_embeds": emb,
        "new_stats": new_stats,
    }


# ----------------------------------------------------------------------------------------------------------------
# deterministic synthetic parameters / inputs shared by golden generation, CPU tests and GPU parity tests
# ----------------------------------------------------------------------------------------------------------------
def make_params(seed: int = 0, depth: int = 12, dim: int = 512, heads: int = 8, n_audio: int = 2560,
                num_labels: int = 500) -> Dict[str, Tensor]:
    """Reference-named parameters with reference-like init scales, drawn in a fixed order from one generator
    (so the same values are reproduced wherever the same torch version runs)."""
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    P: Dict[str, Tensor] = {}

    def bn(prefix, c):
        P[prefix + ".weight"] = 1.0 + 0.1 * randn(c)
        P[prefix + ".bias"] = 0.1 * randn(c)
        P[prefix + ".running_mean"] = torch.zeros(c)
        P[prefix + ".running_var"] = torch.ones(c)

    P["stem3d.0.weight"] = randn(64, 1, 5, 7, 7, std=math.sqrt(2.0 / (64 * 245)) * 4)
    bn("stem3d.1", 64)
    cin = 64
    for li, c in ((1, 64), (2, 128), (3, 256), (4, 512)):
        for b in (0, 1):
            pre = f"resnet.layer{li}.{b}"
            P[pre + ".conv1.weight"] = randn(c, cin if b == 0 else c, 3, 3, std=math.sqrt(2.0 / (c * 9)))
            bn(pre + ".bn1", c)
            P[pre + ".conv2.weight"] = randn(c, c, 3, 3, std=math.sqrt(2.0 / (c * 9)))
            bn(pre + ".bn2", c)
            if b == 0 and li > 1:
                P[pre + ".downsample.0.weight"] = randn(c, cin, 1, 1, std=math.sqrt(2.0 / c))
                bn(pre + ".downsample.1", c)
        cin = c
    inner = heads * 64
    for i in range(depth):
        a, f = f"encoder.layers.{2 * i}", f"encoder.layers.{2 * i + 1}"
        P[a + ".0.g"] = 1.0 + 0.05 * randn(dim)
        for c in "qkv":
            P[f"{a}.1.to_{c}.weight"] = randn(inner, dim, std=(3 * dim) ** -0.5)
        P[a + ".1.to_out.weight"] = randn(dim, inner, std=(3 * inner) ** -0.5)
        P[f + ".0.g"] = 1.0 + 0.05 * randn(dim)
        P[f + ".1.ff.0.proj.weight"] = randn(8 * dim, dim, std=(3 * dim) ** -0.5)
        P[f + ".1.ff.0.proj.bias"] = 0.02 * randn(8 * dim)
        P[f + ".1.ff.3.weight"] = randn(dim, 4 * dim, std=(12 * dim) ** -0.5)
        P[f + ".1.ff.3.bias"] = 0.02 * randn(dim)
    P["audio_projection.weight"] = randn(n_audio, dim, std=(3 * dim) ** -0.5)
    P["audio_projection.bias"] = 0.02 * randn(n_audio)
    P["category_classifier.weight"] = randn(num_labels, dim, std=(3 * dim) ** -0.5)
    P["category_classifier.bias"] = 0.02 * randn(num_labels)
    P["cls_token"] = randn(1, 1, dim)
    return P


def make_inputs(seed: int, B: int, T: int = 29, S: int = 88, A: int = 4, G: int = 2, V: int = 320,
                num_labels: int = 500, extra_tokens: int = 0):
    """SURVEY.md section 8(d) synthetic inputs: N(0,1) clips, uniform audio tokens / labels."""
    g = torch.Generator().manual_seed(seed)
    videos = torch.randn(B, 1, T, S, S, generator=g)
    tokens = torch.randint(0, V, (B, T * A + extra_tokens, G), generator=g)
    labels = torch.randint(0, num_labels, (B,), generator=g)
    word_mask = torch.zeros(B, 1)
    return videos, tokens, labels, word_mask
