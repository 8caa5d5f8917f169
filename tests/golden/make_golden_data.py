"""Generates tests/golden/jpeg.pt and tests/golden/video_transform.pt in the build container.

jpeg.pt: JPEG files written by Pillow's bundled libjpeg-turbo -- including the format the reference's preprocessing
writes (preprocess_pkl.py:182 `TurboJPEG().encode(frame)`: 96x112 colour, quality 85, 4:2:2) -- and the luminance planes
libjpeg-turbo decodes from them in grayscale output mode (what data.py:41 `decode(img, pixel_format=TJPF_GRAY)` returns):
Pillow draft mode "L", cross-checked against OpenCV IMREAD_GRAYSCALE (both drive libjpeg with JCS_GRAYSCALE, islow IDCT).

video_transform.pt: outputs of the reference's OWN transform stacks (data.py:157-171: torchvision modules + the
reference's FunctionalModule and TimeMask from LRW/video/src/augment.py) under fixed torch / `random` seeds, with the
decisions (flip, crop box, mask span) recorded through the same RNG replay that syncvsr_b200/data.py performs.
TimeMask is built with T=15 (data.py:162 passes 0.6*25 = 15.0; `random.randint` needs an int on Python >= 3.12)."""
import io
import random
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent


def synth_frame(h, w, seed, color):
    """Smooth textured image (low-pass noise + edges) so the files have realistic run/size statistics."""
    rng = np.random.default_rng(seed)
    base = rng.normal(size=(h // 4 + 2, w // 4 + 2, 3 if color else 1))
    img = np.kron(base, np.ones((4, 4, 1)))[:h, :w]
    yy, xx = np.mgrid[0:h, 0:w]
    img = img * 40 + 110 + 50 * np.sin(xx / 7.0 + seed)[..., None] + rng.normal(size=img.shape) * 6
    img[h // 3:h // 3 + 5] += 60  # a hard edge
    img = np.clip(img, 0, 255).astype(np.uint8)
    return img if color else img[..., 0]


def encode(img, **kw):
    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", **kw)
    return buf.getvalue()


def decode_libjpeg_gray(data):
    import cv2
    from PIL import Image

    im = Image.open(io.BytesIO(data))
    im.draft("L", im.size)
    a = np.asarray(im.convert("L") if im.mode != "L" else im)
    b = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
    assert im.mode == "L" and np.array_equal(a, b), "Pillow draft-L and OpenCV grayscale decodes disagree"
    return a


def make_jpeg():
    cases = {}
    specs = [
        ("lrw_422_q85", 96, 112, True, dict(quality=85, subsampling=1)),   # TurboJPEG.encode defaults
        ("lrw_422_q85_b", 96, 112, True, dict(quality=85, subsampling=1)),
        ("gray_96", 96, 96, False, dict(quality=90)),
        ("gray_odd", 37, 50, False, dict(quality=75)),
        ("c420_odd", 47, 33, True, dict(quality=60, subsampling=2)),
        ("c444_opt", 40, 56, True, dict(quality=95, subsampling=0, optimize=True)),
        ("gray_q100", 24, 24, False, dict(quality=100)),
        ("gray_q5", 64, 64, False, dict(quality=5)),
        ("c422_restart", 96, 112, True, dict(quality=85, subsampling=1, restart_marker_blocks=3)),
        ("gray_restart_rows", 48, 40, False, dict(quality=80, restart_marker_rows=1)),
    ]
    for i, (name, h, w, color, kw) in enumerate(specs):
        data = encode(synth_frame(h, w, 100 + i, color), **kw)
        cases[name] = dict(jpeg=data, gray=torch.from_numpy(decode_libjpeg_gray(data).copy()))
        print(name, len(data), "bytes", cases[name]["gray"].shape, int(cases[name]["gray"].sum()))
    # noise image: long codes, large coefficients
    rng = np.random.default_rng(5)
    data = encode(rng.integers(0, 256, size=(32, 48), dtype=np.uint8), quality=98)
    cases["gray_noise"] = dict(jpeg=data, gray=torch.from_numpy(decode_libjpeg_gray(data).copy()))
    torch.save(cases, OUT / "jpeg.pt")


def make_transform():
    import torchvision
    from oracle import ref_loader as rl
    from syncvsr_b200.data import transform_plan

    rl.load_reference_lrw()
    import augment as ref_aug  # the reference's module

    mean, std = 0.421, 0.165
    cases = {}
    for name, (T, H, W, crop, train, rrc, tmask, val_resize, seed) in {
        "train_rrc_tm": (6, 96, 96, (96, 96), True, True, True, False, 3),
        "train_rrc_112": (5, 48, 56, (44, 44), True, True, True, False, 4),
        "train_rrc_long": (29, 24, 24, (24, 24), True, True, True, False, 8),
        "train_plain": (12, 40, 48, (40, 48), True, False, True, False, 5),
        "val_center": (4, 96, 96, (88, 88), False, False, False, False, 6),
        "val_resize": (3, 96, 112, (88, 88), False, False, False, True, 7),
    }.items():
        B = 3
        g = torch.Generator().manual_seed(seed)
        frames = torch.randint(0, 256, (B, T, H // 4, W // 4), generator=g, dtype=torch.uint8)
        frames = frames.repeat_interleave(4, 2).repeat_interleave(4, 3).contiguous()  # blocky texture
        frames = (frames.float() * 0.8 + torch.rand(B, T, H, W, generator=g) * 50).clamp(0, 255).to(torch.uint8)
        if train:
            tf = torch.nn.Sequential(
                ref_aug.FunctionalModule(lambda x: x / 255.0),
                torchvision.transforms.RandomHorizontalFlip(p=0.5),
                torchvision.transforms.RandomResizedCrop(size=crop, scale=(0.6, 1.0)) if rrc else torch.nn.Identity(),
                torchvision.transforms.Grayscale(),
                ref_aug.TimeMask(T=15, n_mask=1) if tmask else torch.nn.Identity(),
                torchvision.transforms.Normalize(mean, std),
            )
        else:
            tf = torch.nn.Sequential(
                ref_aug.FunctionalModule(lambda x: x / 255.0),
                torchvision.transforms.Resize(crop) if val_resize else torchvision.transforms.CenterCrop(crop),
                torchvision.transforms.Grayscale(),
                torchvision.transforms.Normalize(mean, std),
            )
        torch.manual_seed(seed), random.seed(seed)
        outs = []
        for b in range(B):  # Dataset.__getitem__ (data.py:42-46, 68): [T,1,H,W] u8 -> transform -> [1,T,H,W]
            v = frames[b].unsqueeze(1)
            outs.append(tf(v).permute(1, 0, 2, 3))
        ref = torch.stack(outs)
        torch.manual_seed(seed), random.seed(seed)
        xf, size = transform_plan(B, T, H, W, crop, train, rrc, tmask, val_resize)
        cases[name] = dict(frames=frames, out=ref.clone(), xform=xf, size=size, seed=seed,
                           cfg=dict(crop=crop, train=train, rrc=rrc, tmask=tmask, val_resize=val_resize))
        print(name, tuple(ref.shape), float(ref.sum()), xf.tolist())
    torch.save(cases, OUT / "video_transform.pt")


if __name__ == "__main__":
    make_jpeg()
    make_transform()
