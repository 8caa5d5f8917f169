"""Device-side mirror of the reference's data path (/root/reference/LRW/video/src/data.py:32-68 `__getitem__`, and the
transform pipelines built in data.py:156-171).

The reference decodes every JPEG frame with TurboJPEG and runs the torchvision transform stack per sample on CPU
DataLoader workers, then ships f32 tensors to the GPU. Here a batch goes to the device as its JPEG bytes
(~4 KB per 96x112 frame instead of 43 KB of f32) and
  * `JpegBatchDecoder.decode` -> `svsr_jpeg_parse` (host marker walk) + `svsr_jpeg_decode_gray` (Huffman + integer IDCT
    kernels, bit-identical to libjpeg-turbo's luminance output),
  * `VideoTransform` -> `svsr_video_transform`: one fused pass for x/255, flip, (random-resized / centre) crop + antialiased
    bilinear resize, TimeMask and Normalize.
The random decisions are drawn on the host by `transform_plan` with the same calls, in the same order, as the reference's
modules make them for one clip after another (torch's global CPU generator for RandomHorizontalFlip and
RandomResizedCrop.get_params, Python's `random` for TimeMask, augment.py:131-137), so a seeded run picks the same boxes.
There is no CPU fallback: without libsvsr.so these raise."""
from __future__ import annotations

import ctypes as C
import math
import random
from typing import Sequence, Tuple

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr

MEAN, STD = 0.421, 0.165  # data.py:150


def _rrc_params(height: int, width: int, scale=(0.6, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)) -> Tuple[int, int, int, int]:
    """torchvision RandomResizedCrop.get_params (data.py:160 uses scale=(0.6, 1.0), default ratio): same RNG calls."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def transform_plan(B: int, T: int, H: int, W: int, crop_size: Tuple[int, int], train: bool, use_rrc: bool = True,
                   use_timemask: bool = True, use_val_resize: bool = False, time_mask_T: float = 0.6 * 25):
    """Returns (xform int32 [B, 8] CPU tensor, (OH, OW)). Row = {flip, top, left, crop_h, crop_w, mask_t0, mask_t1, 0}.

    train (data.py:157-164): per clip RandomHorizontalFlip(0.5) -> RandomResizedCrop(crop_size, scale=(0.6, 1)) if use_rrc
    -> TimeMask(T=15, n_mask=1) if use_timemask. eval (data.py:167-172): Resize(crop_size) if use_val_resize else
    CenterCrop(crop_size)."""
    xf = torch.zeros(B, 8, dtype=torch.int32)
    out_size = tuple(crop_size)
    for b in range(B):
        if train:
            flip = bool(torch.rand(1) < 0.5)
            if use_rrc:
                top, left, h, w = _rrc_params(H, W)
            else:
                top, left, h, w = 0, 0, H, W
                out_size = (H, W)
            m0 = m1 = 0
            if use_timemask:  # augment.py:131-137 with n_mask = 1
                t = random.randint(0, int(min(time_mask_T, T)))
                m0 = random.randint(0, T - t)
                m1 = m0 + t
            xf[b] = torch.tensor([int(flip), top, left, h, w, m0, m1, 0], dtype=torch.int32)
        else:
            if use_val_resize:
                top, left, h, w = 0, 0, H, W
            else:  # torchvision center_crop
                h, w = crop_size
                if h > H or w > W:
                    raise ValueError(f"CenterCrop {crop_size} larger than the {H}x{W} frames")
                top, left = int(round((H - h) / 2.0)), int(round((W - w) / 2.0))
            xf[b] = torch.tensor([0, top, left, h, w, 0, 0, 0], dtype=torch.int32)
    return xf, out_size


class VideoTransform:
    """frames u8 [B, T, H, W] (device) -> videos f32 [B, 1, T, OH, OW] ready for TransformerLightningModule.forward."""

    def __init__(self, crop_size: Tuple[int, int], train: bool, use_rrc: bool = True, use_timemask: bool = True,
                 use_val_resize: bool = False):
        self.crop_size, self.train = tuple(crop_size), train
        self.use_rrc, self.use_timemask, self.use_val_resize = use_rrc, use_timemask, use_val_resize

    def plan(self, B: int, T: int, H: int, W: int):
        return transform_plan(B, T, H, W, self.crop_size, self.train, self.use_rrc, self.use_timemask, self.use_val_resize)

    @torch.no_grad()
    def __call__(self, frames: torch.Tensor, plan=None) -> torch.Tensor:
        if frames.dtype != torch.uint8 or frames.dim() != 4 or not frames.is_cuda or not frames.is_contiguous():
            raise ValueError("VideoTransform expects a contiguous CUDA uint8 tensor [B, T, H, W]")
        B, T, H, W = frames.shape
        xf, (OH, OW) = plan if plan is not None else self.plan(B, T, H, W)
        xf_d = xf.to(frames.device, non_blocking=True)
        out = torch.empty(B, 1, T, OH, OW, device=frames.device, dtype=torch.float32)
        clip_sum = torch.empty(B, device=frames.device, dtype=torch.float64)
        tm = int(self.train and self.use_timemask)
        check(lib().svsr_video_transform(ptr(frames), ptr(xf_d), ptr(out), ptr(clip_sum), C.c_int(B), C.c_int(T), C.c_int(H),
                                         C.c_int(W), C.c_int(OH), C.c_int(OW), C.c_float(MEAN), C.c_float(STD), C.c_int(tm),
                                         stream_ptr()), "svsr_video_transform")
        return out


class JpegBatchDecoder:
    """data.py:41 for a whole batch: list of JPEG byte strings (all frames the same size) -> u8 [n, H, W] on the device."""

    DESC_INTS, HUFF_BYTES, QCAP, HCAP = 24, 1536, 64, 64

    def __init__(self, device: torch.device | str = "cuda"):
        self.device = torch.device(device)
        self._stage = None  # reusable pinned staging buffer for the JPEG bytes
        self._copied = None  # event recorded after the last H2D copy out of it

    def _pinned(self, nbytes: int) -> torch.Tensor:
        if self._stage is None or self._stage.numel() < nbytes:
            self._stage = torch.empty(max(nbytes * 5 // 4, 1 << 20), dtype=torch.uint8).pin_memory()
        return self._stage[:nbytes]

    def parse(self, frames: Sequence[bytes]):
        """Host side only (works without a GPU): concatenated blob, per-frame descriptors, table pools, geometry."""
        n = len(frames)
        if n == 0:
            raise ValueError("JpegBatchDecoder: empty batch")
        blob = np.frombuffer(b"".join(frames), dtype=np.uint8)
        offsets = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(f) for f in frames], out=offsets[1:])
        desc = np.zeros((n, self.DESC_INTS), dtype=np.int32)
        qt = np.zeros((self.QCAP, 64), dtype=np.uint16)
        ht = np.zeros((self.HCAP, self.HUFF_BYTES), dtype=np.uint8)
        nq, nh = C.c_int(0), C.c_int(0)
        cp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(lib().svsr_jpeg_parse(cp(blob), cp(offsets), C.c_int(n), cp(desc), cp(qt), C.c_int(self.QCAP), C.byref(nq),
                                    cp(ht), C.c_int(self.HCAP), C.byref(nh)), "svsr_jpeg_parse")
        W, H = int(desc[0, 2]), int(desc[0, 3])
        if not (np.all(desc[:, 2] == W) and np.all(desc[:, 3] == H)):
            raise ValueError("JpegBatchDecoder: all frames of a batch must have the same size")
        nc = desc[:, 5]
        h0 = np.where(nc == 1, 1, desc[:, 6])
        v0 = np.where(nc == 1, 1, desc[:, 7])
        hmax = np.where(nc == 1, 1, np.maximum(desc[:, 6], np.maximum(desc[:, 11], desc[:, 16])))
        vmax = np.where(nc == 1, 1, np.maximum(desc[:, 7], np.maximum(desc[:, 12], desc[:, 17])))
        bw = int(np.max(-(-W // (8 * hmax)) * h0))
        bh = int(np.max(-(-H // (8 * vmax)) * v0))
        return dict(blob=blob, desc=desc, qt=qt[:max(nq.value, 1)], ht=ht[:max(nh.value, 1)], W=W, H=H, bw=bw, bh=bh, n=n)

    @torch.no_grad()
    def decode(self, frames: Sequence[bytes]) -> torch.Tensor:
        p = self.parse(frames)
        dev = self.device
        up = lambda a: torch.from_numpy(a).to(dev, non_blocking=True)  # noqa: E731
        if self._copied is not None:
            self._copied.synchronize()  # the previous batch's H2D copy has left the staging buffer
        stage = self._pinned(p["blob"].size)
        stage.numpy()[:] = p["blob"]
        blob = stage.to(dev, non_blocking=True)
        self._copied = torch.cuda.Event()
        self._copied.record()
        desc, qt, ht = up(p["desc"]), up(p["qt"].view(np.int16)), up(p["ht"])
        n, W, H, bw, bh = p["n"], p["W"], p["H"], p["bw"], p["bh"]
        coef = torch.empty(n * bw * bh * 64, device=dev, dtype=torch.int16)
        out = torch.empty(n, H, W, device=dev, dtype=torch.uint8)
        check(lib().svsr_jpeg_decode_gray(ptr(blob), ptr(desc), C.c_int(n), ptr(qt), ptr(ht), ptr(coef), ptr(out), C.c_int(W),
                                          C.c_int(H), C.c_int(bw), C.c_int(bh), stream_ptr()), "svsr_jpeg_decode_gray")
        return out


def load_clips(samples: Sequence[dict], decoder: JpegBatchDecoder, transform: VideoTransform, plan=None) -> torch.Tensor:
    """samples: the reference's pkl dicts ({"video": [JPEG bytes] * T, ...}, preprocess_pkl.py:118-225), all T equal.
    Returns videos f32 [B, 1, T, OH, OW] on the device (what the DataLoader of data.py:185-192 would have collated)."""
    T = len(samples[0]["video"])
    if any(len(s["video"]) != T for s in samples):
        raise ValueError("load_clips: clips of one batch must have the same number of frames")
    frames = decoder.decode([f for s in samples for f in s["video"]])
    return transform(frames.view(len(samples), T, frames.shape[1], frames.shape[2]), plan)
