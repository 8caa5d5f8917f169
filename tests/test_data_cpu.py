"""Data-path oracle pinned to libjpeg-turbo / the reference's transform stack (fixtures from tests/golden/make_golden_data.py),
plus the host-side pieces of the product path that need no GPU (marker parser, RNG replay of the transform decisions)."""
import random
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import data_oracle as do

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def jpegs():
    return torch.load(GOLD / "jpeg.pt")


@pytest.fixture(scope="module")
def xforms():
    return torch.load(GOLD / "video_transform.pt")


def test_jpeg_oracle_is_bit_identical_to_libjpeg_turbo(jpegs):
    assert len(jpegs) >= 10
    for name, c in jpegs.items():
        got = do.jpeg_decode_gray(c["jpeg"])
        assert got.shape == tuple(c["gray"].shape), name
        assert np.array_equal(got, c["gray"].numpy()), f"{name}: {np.abs(got.astype(int) - c['gray'].numpy()).max()}"


def test_islow_idct_known_answers():
    # DC only: every pixel = clamp(round(dc / 8) + 128); zero block = mid-grey
    z = np.zeros((8, 8), np.int64)
    assert np.all(do.idct_islow(z) == 128)
    z[0, 0] = 80
    assert np.all(do.idct_islow(z) == 138)
    z[0, 0] = 8 * 300
    assert np.all(do.idct_islow(z) == 255)      # saturates through the range-limit table
    z[0, 0] = -8 * 300
    assert np.all(do.idct_islow(z) == 0)


def test_transform_oracle_matches_reference_stack(xforms):
    for name, c in xforms.items():
        cfg = c["cfg"]
        for b in range(c["frames"].shape[0]):
            flip, top, left, h, w, m0, m1, _ = c["xform"][b].tolist()
            got = do.video_transform(c["frames"][b], bool(flip), (top, left, h, w), c["size"],
                                     (m0, m1) if cfg["train"] and cfg["tmask"] else None)
            torch.testing.assert_close(got, c["out"][b], rtol=1e-5, atol=1e-5, msg=lambda m: f"{name}[{b}]: {m}")


def test_transform_plan_replays_the_reference_rng_order(xforms):
    from syncvsr_b200.data import transform_plan

    for name, c in xforms.items():
        cfg = c["cfg"]
        B, T, H, W = c["frames"].shape
        torch.manual_seed(c["seed"]), random.seed(c["seed"])
        xf, size = transform_plan(B, T, H, W, cfg["crop"], cfg["train"], cfg["rrc"], cfg["tmask"], cfg["val_resize"])
        assert torch.equal(xf, c["xform"]) and tuple(size) == tuple(c["size"]), name
    # eval plans consume no randomness
    s0 = torch.get_rng_state()
    transform_plan(2, 29, 96, 96, (88, 88), train=False)
    assert torch.equal(s0, torch.get_rng_state())


def test_jpeg_parser_runs_on_the_host(jpegs):
    from syncvsr_b200.data import JpegBatchDecoder

    dec = JpegBatchDecoder("cpu")
    lrw = [jpegs["lrw_422_q85"]["jpeg"], jpegs["lrw_422_q85_b"]["jpeg"], jpegs["c422_restart"]["jpeg"]]
    p = dec.parse(lrw)
    assert (p["W"], p["H"], p["bw"], p["bh"]) == (112, 96, 14, 12)
    assert p["desc"][:, 5].tolist() == [3, 3, 3] and p["desc"][:, 4].tolist() == [0, 0, 3]
    assert p["desc"][0, 6:8].tolist() == [2, 1]                     # 4:2:2 luminance sampling
    assert len(p["qt"]) == 2 and len(p["ht"]) == 4                  # identical tables are pooled across frames
    p = dec.parse([jpegs["gray_odd"]["jpeg"]])
    assert (p["W"], p["H"], p["bw"], p["bh"]) == (50, 37, 7, 5) and p["desc"][0, 5] == 1
    with pytest.raises(Exception, match="same size"):
        dec.parse([jpegs["gray_odd"]["jpeg"], jpegs["gray_96"]["jpeg"]])
    with pytest.raises(Exception, match="SOI"):
        dec.parse([b"not a jpeg file"])
    import io
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(jpegs["gray_96"]["gray"].numpy()).save(buf, format="JPEG", progressive=True)
    with pytest.raises(Exception, match="progressive"):
        dec.parse([buf.getvalue()])
