"""Data path on the device (svsr_jpeg_decode_gray, svsr_video_transform through syncvsr_b200/data.py) against the
libjpeg-turbo / reference-transform fixtures and the CPU oracle. JPEG decoding is integer work: bit-exact. The transform is
fp32 arithmetic: |diff| <= 1e-5 (normalised units; one fp32 rounding of products that ATen and the kernel order differently)."""
import random
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
TOL = 1e-5


@pytest.fixture(scope="module")
def data():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from syncvsr_b200 import data as d

    return d


@pytest.fixture(scope="module")
def jpegs():
    return torch.load(GOLD / "jpeg.pt")


@pytest.fixture(scope="module")
def xforms():
    return torch.load(GOLD / "video_transform.pt")


def test_jpeg_decode_bit_exact_per_file(data, jpegs):
    dec = data.JpegBatchDecoder()
    for name, c in jpegs.items():
        got = dec.decode([c["jpeg"]])[0].cpu()
        assert torch.equal(got, c["gray"]), f"{name}: max diff {(got.int() - c['gray'].int()).abs().max()}"


def test_jpeg_decode_full_batch_mixed_tables(data, jpegs):
    """B=64 x T=29 frames in the reference's on-disk format (96x112, 4:2:2, q85), files with and without restart
    intervals interleaved in one launch."""
    names = ["lrw_422_q85", "c422_restart", "lrw_422_q85_b"]
    n = 64 * 29
    files = [jpegs[names[i % 3]]["jpeg"] for i in range(n)]
    out = data.JpegBatchDecoder().decode(files)
    assert out.shape == (n, 96, 112) and out.dtype == torch.uint8
    for k, nm in enumerate(names):
        ref = jpegs[nm]["gray"].cuda()
        assert bool((out[k::3] == ref).all()), nm


def test_jpeg_decode_matches_oracle_on_fresh_files(data):
    """Files encoded on this box (Pillow) that are not in the fixtures: GPU == CPU oracle == libjpeg-turbo."""
    import io

    from PIL import Image

    from oracle import data_oracle as do

    rng = np.random.default_rng(9)
    files, refs = [], []
    for i in range(6):
        img = rng.integers(0, 256, size=(12, 14), dtype=np.uint8).repeat(8, 0).repeat(8, 1)
        img = np.clip(img.astype(np.int32) + rng.integers(-20, 20, size=img.shape), 0, 255).astype(np.uint8)
        buf = io.BytesIO()
        Image.fromarray(np.stack([img, img[::-1], img[:, ::-1]], -1)).save(buf, format="JPEG", quality=40 + 10 * i,
                                                                           subsampling=1)
        files.append(buf.getvalue())
        im = Image.open(io.BytesIO(files[-1]))
        im.draft("L", im.size)
        refs.append(np.asarray(im))
    out = data.JpegBatchDecoder().decode(files).cpu().numpy()
    for i in range(6):
        assert np.array_equal(out[i], refs[i]) and np.array_equal(do.jpeg_decode_gray(files[i]), refs[i])


def test_jpeg_errors_are_loud(data, jpegs):
    dec = data.JpegBatchDecoder()
    with pytest.raises(Exception, match="same size"):
        dec.decode([jpegs["gray_odd"]["jpeg"], jpegs["gray_96"]["jpeg"]])
    with pytest.raises(Exception, match="empty"):
        dec.decode([])


def test_video_transform_matches_reference_stack(data, xforms):
    for name, c in xforms.items():
        cfg = c["cfg"]
        vt = data.VideoTransform(cfg["crop"], cfg["train"], cfg["rrc"], cfg["tmask"], cfg["val_resize"])
        got = vt(c["frames"].cuda(), (c["xform"], c["size"])).cpu()
        assert got.shape == c["out"].shape, name
        assert (got - c["out"]).abs().max().item() <= TOL, f"{name}: {(got - c['out']).abs().max().item()}"


def test_video_transform_full_batch_against_oracle(data):
    from oracle import data_oracle as do

    B, T, H, W = 64, 29, 96, 96
    g = torch.Generator().manual_seed(21)
    frames = torch.randint(0, 256, (B, T, H, W), generator=g, dtype=torch.uint8)
    vt = data.VideoTransform((96, 96), train=True)
    torch.manual_seed(5), random.seed(5)
    xf, size = vt.plan(B, T, H, W)
    got = vt(frames.cuda(), (xf, size)).cpu()
    assert got.shape == (B, 1, T, 96, 96)
    for b in range(0, B, 7):
        flip, top, left, h, w, m0, m1, _ = xf[b].tolist()
        ref = do.video_transform(frames[b], bool(flip), (top, left, h, w), size, (m0, m1))
        assert (got[b] - ref).abs().max().item() <= TOL
    # properties at full size: masked frames are constant, un-masked frames of un-flipped full-image crops are the input
    for b in range(B):
        flip, top, left, h, w, m0, m1, _ = xf[b].tolist()
        if m1 > m0:
            blk = got[b, 0, m0:m1]
            assert float(blk.max() - blk.min()) == 0.0


def test_load_clips_jpeg_to_model_input(data, jpegs):
    from oracle import data_oracle as do

    names = ["lrw_422_q85", "lrw_422_q85_b", "c422_restart"]
    T = 5
    samples = [{"video": [jpegs[names[(b + t) % 3]]["jpeg"] for t in range(T)]} for b in range(3)]
    vt = data.VideoTransform((88, 88), train=False)
    out = data.load_clips(samples, data.JpegBatchDecoder(), vt).cpu()
    assert out.shape == (3, 1, T, 88, 88)
    for b in range(3):
        fr = torch.stack([jpegs[names[(b + t) % 3]]["gray"] for t in range(T)])
        ref = do.video_transform(fr, False, (4, 12, 88, 88), (88, 88))
        assert (out[b] - ref).abs().max().item() <= TOL
