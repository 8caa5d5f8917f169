"""TEST / BASELINE INFRASTRUCTURE ONLY (oracle) -- never imported by the product path (syncvsr_b200/*).

The reference LRW module rebuilt from stock `torch.nn` modules, for ONE purpose: timing what the reference runs on
`torch.cuda` (BASELINE.json configs[1] "vs reference torch.cuda"; SURVEY.md section 2.2: "the bar to beat on the same box
is PyTorch-eager under autocast(bf16)") next to the native arm in bench.py (`--impl eager`, `gpu_baseline`).
The reference tree is not on the GPU box and its third-party imports (timm, x-transformers, pytorch-lightning) are not
installed anywhere, so this file restates the module graph of /root/reference/LRW/video/src/lightning.py:36-191:

    stem3d   nn.Sequential(Conv3d, BatchNorm3d, GELU, MaxPool3d)                      lightning.py:49-54
    resnet   timm.create_model("resnet18") == torchvision.models.resnet18 modules     lightning.py:55
    encoder  x_transformers.Encoder 1.9.2 restatement (oracle/xt_encoder.py)          lightning.py:95-105
    heads    nn.Linear category_classifier / audio_projection + F.cross_entropy       lightning.py:82,107,161-174

every FLOP dispatched through ATen -> cuDNN / cuBLAS exactly as the reference does (no custom kernel, no fusion).
`step()` = what Lightning runs per iteration under `precision: bf16` + `gradient_clip_val` + AdamW (train.py:23-33).
tests/test_oracle_cpu.py pins forward() of this module to oracle/lrw_oracle.py on the same state dict."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .xt_encoder import Encoder


def _resnet18_trunk() -> nn.Module:
    import torchvision

    return torchvision.models.resnet18()


class EagerLRW(nn.Module):
    def __init__(self, depth: int = 12, dim: int = 512, heads: int = 8, A: int = 4, G: int = 2, V: int = 320,
                 num_labels: int = 500, lambda_audio: float = 10.0, layer_dropout: float = 0.0, ff_dropout: float = 0.0):
        super().__init__()
        self.A, self.G, self.V, self.lambda_audio = A, G, V, lambda_audio
        self.stem3d = nn.Sequential(
            nn.Conv3d(1, 64, (5, 7, 7), (1, 2, 2), (2, 3, 3), bias=False), nn.BatchNorm3d(64), nn.GELU(),
            nn.MaxPool3d((1, 3, 3), (1, 2, 2), (0, 1, 1)))
        self.resnet = _resnet18_trunk()
        self.encoder = Encoder(dim=dim, depth=depth, heads=heads, attn_dropout=0.0, layer_dropout=layer_dropout,
                               ff_dropout=ff_dropout, use_rmsnorm=True, ff_glu=True, rotary_pos_emb=True)
        self.audio_projection = nn.Linear(dim, A * G * V)
        self.category_classifier = nn.Linear(dim, num_labels)
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))

    def forward_videos(self, videos):  # lightning.py:112-119
        h = self.stem3d(videos).transpose(1, 2).flatten(0, 1)
        for layer in (self.resnet.layer1, self.resnet.layer2, self.resnet.layer3, self.resnet.layer4):
            h = layer(h)
        return h.mean((2, 3)).unflatten(0, (videos.shape[0], -1))

    def forward(self, videos, audio_tokens, labels, word_mask=None):  # lightning.py:133-191 (noWB)
        emb = self.forward_videos(videos)
        B, T, _ = emb.shape
        tok = audio_tokens[:, : T * self.A].flatten()
        x = torch.cat((self.cls_token.expand(B, -1, -1), emb), dim=1)
        last = self.encoder(x)
        logits_category = self.category_classifier(last[:, 0, :]).float()
        loss_category = F.cross_entropy(logits_category, labels)
        logits_audio = self.audio_projection(last[:, 1:, :]).float()
        logits_audio = logits_audio.reshape(B, T, self.A * self.G, self.V)
        loss_audio = F.cross_entropy(logits_audio.reshape(-1, self.V), tok)
        return {"loss_total": loss_category + loss_audio * self.lambda_audio, "loss_category": loss_category,
                "loss_audio": loss_audio, "last_hidden_state": last, "logits_audio": logits_audio}

    def load_oracle_params(self, P) -> None:
        """Reference-named state dict (oracle.lrw_oracle.make_params) -> this module (same keys; the encoder's
        `layers.i.1.ff.*` names are those of the restatement both share)."""
        missing, unexpected = self.load_state_dict(P, strict=False)
        assert not unexpected, unexpected
        assert all(k.startswith(("resnet.conv1", "resnet.bn1", "resnet.fc")) or "num_batches_tracked" in k
                   for k in missing), missing


class EagerStep:
    """One reference training iteration on torch.cuda: autocast(bf16) forward, backward, clip_grad_norm_, AdamW
    (decay on ndim >= 2 only, lightning.py:216-221), zero_grad -- stock PyTorch, channels_last for the 2-D trunk."""

    def __init__(self, model: EagerLRW, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, clip=1.0,
                 channels_last: bool = True):
        self.model, self.clip = model, clip
        if channels_last:
            model.resnet.to(memory_format=torch.channels_last)
        used = [p for k, p in model.named_parameters()
                if not k.startswith(("resnet.conv1", "resnet.bn1", "resnet.fc"))]
        self.params = used
        fused = all(p.is_cuda for p in used)
        self.opt = torch.optim.AdamW([{"params": [p for p in used if p.ndim >= 2]},
                                      {"params": [p for p in used if p.ndim < 2], "weight_decay": 0.0}],
                                     lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, fused=fused)

    def __call__(self, videos, audio_tokens, labels, word_mask=None):
        self.opt.zero_grad(set_to_none=True)
        dev = videos.device.type
        with torch.autocast(dev, dtype=torch.bfloat16):
            out = self.model(videos, audio_tokens, labels, word_mask)
        out["loss_total"].backward()
        torch.nn.utils.clip_grad_norm_(self.params, self.clip)
        self.opt.step()
        return out
