"""On-device mirror of the reference's CutMix (/root/reference/LRW/video/src/augment.py:12-118).

The reference mixes clip by clip in a Python loop: host-RNG decisions (`torch.randint/rand` on CPU tensors, `.item()`),
then in-place slice assignments on `videos` / `audio_tokens`, so a clip that was already mixed can be the source of a
later one. `cutmix_plan` draws the SAME decisions in the SAME order from torch's CPU generator and replays the swaps on
two small index tables (which ORIGINAL clip does frame t / audio row a of clip i come from); `CutMix.forward` then runs
one gather kernel (`svsr_cutmix_gather`) that moves the frames, the int64 token rows (bit-exact), and builds the mixed
soft labels and word masks. Quirks kept on purpose: the audio rows swapped are the VIDEO frame indices [s, s+len), not
scaled by the audio alignment (augment.py:104-109); `len = int(T * rate)` with rate in [0,1); a fair coin decides
whether a clip is mixed at all (augment.py:87)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn

from ._lib import check, lib, ptr, stream_ptr


@dataclass
class CutMixPlan:
    vsrc: torch.Tensor   # int32 [B, T]   source clip of every video frame
    asrc: torch.Tensor   # int32 [B, Ta]  source clip of every audio-token row
    tgt: torch.Tensor    # int32 [B]      partner clip
    rate: torch.Tensor   # fp32  [B]      mix rate
    mixed: torch.Tensor  # uint8 [B]      1 where labels / word mask are mixed


def cutmix_plan(B: int, T: int, Ta: int) -> CutMixPlan:
    """Host decisions + replay of augment.py:37-111 on index tables (CPU tensors; consumes torch's global CPU RNG in the
    reference's order: randint(B), rand(B), then per clip randint(0,2) and, if a cut happens, randint(0, T-len))."""
    target_ids = torch.randint(0, B, (B,))
    target_rates = torch.rand(B)
    vsrc = torch.arange(B, dtype=torch.int32).unsqueeze(1).repeat(1, T)
    asrc = torch.arange(B, dtype=torch.int32).unsqueeze(1).repeat(1, Ta)
    mixed = torch.zeros(B, dtype=torch.uint8)
    for i in range(B):
        mix_rate = target_rates[i].item()
        cut_out_flag = torch.randint(0, 2, (1,))[0].item()
        if cut_out_flag == 1:
            n = int(T * mix_rate)
            if n > 0:
                s = torch.randint(0, T - n, (1,)).item()
                j = int(target_ids[i])
                vsrc[i, s:s + n] = vsrc[j, s:s + n].clone()  # the partner's CURRENT frames (it may already be mixed)
                asrc[i, s:s + n] = asrc[j, s:s + n].clone()  # same (unscaled) row range of the token sequence
                mixed[i] = 1
    return CutMixPlan(vsrc, asrc, target_ids.to(torch.int32), target_rates.clone(), mixed)


class CutMix(nn.Module):
    def __init__(self, num_labels: int, wav2vec=None) -> None:
        super().__init__()
        self.num_labels = num_labels
        self.wav2vec = wav2vec  # frozen quantiser (augment.py:39-41): PyTorch module of the caller, off the hot path

    @torch.no_grad()
    def forward(self, videos: torch.Tensor, audios: torch.Tensor, labels: torch.Tensor,
                word_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        if self.wav2vec:
            audio_tokens = self.wav2vec.feature_extractor(audios)
            audio_tokens = self.wav2vec.vector_quantizer.forward_idx(audio_tokens)[1]
        else:
            audio_tokens = audios
        if not videos.is_cuda:
            raise ValueError("CutMix runs on the device: pass CUDA tensors (there is no CPU path)")
        B, _, T = videos.shape[:3]
        dev = videos.device
        videos = videos.float().contiguous()
        audio_tokens = audio_tokens.to(dev, torch.long).contiguous()
        labels = labels.to(dev, torch.long).contiguous()
        wm = word_mask.to(dev, torch.float32).contiguous()
        Ta, G = audio_tokens.shape[1], audio_tokens.shape[2]
        plan = cutmix_plan(B, T, Ta)
        t = {k: getattr(plan, k).to(dev, non_blocking=True) for k in ("vsrc", "asrc", "tgt", "rate", "mixed")}
        v_out, a_out = torch.empty_like(videos), torch.empty_like(audio_tokens)
        soft = torch.empty(B, self.num_labels, device=dev, dtype=torch.float32)
        wm_out = torch.empty_like(wm)
        frame = videos[0, 0, 0].numel()
        check(lib().svsr_cutmix_gather(ptr(videos), ptr(v_out), ptr(t["vsrc"]), C.c_int(B), C.c_int(T), C.c_int64(frame),
                                       ptr(audio_tokens), ptr(a_out), ptr(t["asrc"]), C.c_int(Ta), C.c_int(G), ptr(labels),
                                       ptr(t["tgt"]), ptr(t["rate"]), ptr(t["mixed"]), ptr(soft), C.c_int(self.num_labels),
                                       ptr(wm), ptr(wm_out), C.c_int(wm.shape[1]), stream_ptr()), "svsr_cutmix_gather")
        return v_out, a_out, soft, wm_out
