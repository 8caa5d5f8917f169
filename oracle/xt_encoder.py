"""TEST INFRASTRUCTURE ONLY (oracle). Not imported by the product path.

PyTorch restatement of `x_transformers.Encoder` as pinned by the reference
(`pip install x-transformers==1.9.2`, /root/reference/LRW/video/setup.sh:30) for the exact call made at
/root/reference/LRW/video/src/lightning.py:95-105 and invoked at lightning.py:158:

    Encoder(dim=D, depth=12, heads=8, attn_dropout=0., layer_dropout=.2, ff_dropout=.3,
            use_rmsnorm=True, ff_glu=True, rotary_pos_emb=True)

PARITY UNPINNED: the x-transformers package is a third-party dependency that is not vendored under
/root/reference and is not installable here (no network); the reference holds no test or golden vector
for it. This file restates the published 1.9.x algorithm:
  * 2*depth sublayers ('a','f')*depth, pre-norm, plain residual, no final norm in AttentionLayers
  * RMSNorm:  x / clamp(||x||_2 * D^-0.5, 1e-8) * g
  * Attention: dim_head=64, to_q/k/v/out bias-free, rotary (32 dims/head) on q, k AND v,
               softmax in fp32, scale = 64^-0.5
  * FeedForward(glu=True, mult=4): Linear(D, 2*4D) -> x * GELU(gate) -> Dropout -> Linear(4D, D)
  * layer_dropout: each sublayer skipped with prob p (host RNG) in training.
The three details the survey flags as uncertain are switches (rotary_v, final_norm), defaults as above.
"""
from __future__ import annotations

import random

import torch
import torch.nn as nn
import torch.nn.functional as F


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-8):
        super().__init__()
        self.scale = dim ** -0.5
        self.eps = eps
        self.g = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        norm = torch.norm(x, dim=-1, keepdim=True) * self.scale
        return x / norm.clamp(min=self.eps) * self.g


def rotary_freqs(n: int, rot_dim: int, device=None) -> torch.Tensor:
    inv_freq = 1.0 / (10000 ** (torch.arange(0, rot_dim, 2, device=device).float() / rot_dim))
    t = torch.arange(n, device=device).float()
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    return torch.cat((freqs, freqs), dim=-1)  # [n, rot_dim]


def rotate_half(x):
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary(t, freqs):
    rot = freqs.shape[-1]
    tl, tr = t[..., :rot], t[..., rot:]
    tl = tl * freqs.cos() + rotate_half(tl) * freqs.sin()
    return torch.cat((tl, tr), dim=-1)


class Attention(nn.Module):
    def __init__(self, dim: int, heads: int = 8, dim_head: int = 64, dropout: float = 0.0, rotary_v: bool = True):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head, self.scale, self.rotary_v = heads, dim_head, dim_head ** -0.5, rotary_v
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, freqs):
        b, n, _ = x.shape
        h, d = self.heads, self.dim_head
        q, k, v = (f(x).view(b, n, h, d).transpose(1, 2) for f in (self.to_q, self.to_k, self.to_v))
        q, k = apply_rotary(q, freqs), apply_rotary(k, freqs)
        if self.rotary_v:
            v = apply_rotary(v, freqs)
        dots = torch.einsum("bhid,bhjd->bhij", q, k) * self.scale
        attn = F.softmax(dots, dim=-1, dtype=torch.float32).type(dots.dtype)
        attn = self.dropout(attn)
        out = torch.einsum("bhij,bhjd->bhid", attn, v)
        out = out.transpose(1, 2).reshape(b, n, h * d)
        return self.to_out(out)


class GLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4, dropout: float = 0.0):
        super().__init__()
        inner = int(dim * mult)
        self.ff = nn.Sequential(GLU(dim, inner), nn.Identity(), nn.Dropout(dropout), nn.Linear(inner, dim))

    def forward(self, x):
        return self.ff(x)


class Encoder(nn.Module):
    """Signature-compatible with the reference's call (lightning.py:95-105)."""

    def __init__(self, dim, depth, heads=8, attn_dropout=0.0, layer_dropout=0.0, ff_dropout=0.0, use_rmsnorm=True,
                 ff_glu=True, rotary_pos_emb=True, rotary_v=True, final_norm=False, **_unused):
        super().__init__()
        assert use_rmsnorm and ff_glu and rotary_pos_emb, "restatement covers the reference's configuration only"
        self.dim, self.depth, self.layer_dropout = dim, depth, layer_dropout
        self.rot_dim = max(64 // 2, 32)
        self.layers = nn.ModuleList()
        for _ in range(depth):
            self.layers.append(nn.ModuleList([RMSNorm(dim), Attention(dim, heads, 64, attn_dropout, rotary_v)]))
            self.layers.append(nn.ModuleList([RMSNorm(dim), FeedForward(dim, 4, ff_dropout)]))
        self.final_norm = RMSNorm(dim) if final_norm else None

    def forward(self, x):
        freqs = rotary_freqs(x.shape[1], self.rot_dim, x.device)
        for i, (norm, block) in enumerate(self.layers):
            if self.training and self.layer_dropout > 0.0 and random.random() < self.layer_dropout:
                continue
            y = norm(x)
            y = block(y, freqs) if i % 2 == 0 else block(y)
            x = y + x
        if self.final_norm is not None:
            x = self.final_norm(x)
        return x
