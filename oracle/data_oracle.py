"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's data path. Imported by tests/
and tools/bench_datapath.py's CPU baseline leg, never by the product package.

What it restates
  * `jpeg_decode_gray`: data.py:41 `TurboJPEG().decode(img, pixel_format=TJPF_GRAY)`. The arithmetic lives in a
    third-party dependency that is NOT under /root/reference: PyTurboJPEG (unpinned in LRW/video/setup.sh) -> libjpeg-turbo.
    This file restates the published algorithm: ITU-T T.81 sequential Huffman decoding (Annex F.2: DECODE, RECEIVE,
    EXTEND, restart intervals, byte stuffing) and libjpeg's integer "islow" inverse DCT (jidctint.c: LL&M factorisation,
    CONST_BITS = 13, PASS1_BITS = 2, range-limit table with RANGE_MASK), the default DCT of tjDecompress2. Grayscale
    output of a YCbCr file is its luminance plane (libjpeg skips the chroma components when out_color_space is
    JCS_GRAYSCALE).
    PINNED: tests/golden/jpeg.pt holds JPEG files written by Pillow's bundled libjpeg-turbo (the 4:2:2 quality-85 colour
    files TurboJPEG.encode defaults to, preprocess_pkl.py:182, plus gray / 4:2:0 / 4:4:4 / odd sizes / restart intervals /
    optimised Huffman tables) together with the planes libjpeg-turbo itself decodes (Pillow draft mode "L" and OpenCV
    IMREAD_GRAYSCALE, which agree); tests/test_data_cpu.py requires this oracle to reproduce them bit for bit.
  * `video_transform`: the train / val transform pipelines of data.py:157-171 with the random decisions made explicit
    (flip, crop box, TimeMask span): x/255 -> hflip -> resized_crop (bilinear, antialias) -> TimeMask (augment.py:120-143)
    -> Normalize(0.421, 0.165). PINNED: tests/golden/video_transform.pt was produced by the reference's own
    nn.Sequential (torchvision transforms + the reference's FunctionalModule / TimeMask) under fixed seeds.
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                   21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53,
                   60, 61, 54, 47, 55, 62, 63])


# ---------------------------------------------------------------------------------------------------- JPEG -------
class _Huff:
    """T.81 Annex C code assignment + F.2.2.3 decode tables."""

    def __init__(self, bits, vals):
        self.vals = vals
        self.maxcode, self.valoff = [-1] * 18, [0] * 17
        code = k = 0
        for length in range(1, 17):
            self.valoff[length] = k - code
            code += bits[length]
            k += bits[length]
            self.maxcode[length] = code - 1 if bits[length] else -1
            code <<= 1


class _Bits:
    def __init__(self, data: bytes, pos: int):
        self.d, self.p, self.buf, self.cnt, self.marker = data, pos, 0, 0, False

    def _byte(self):
        if self.marker or self.p >= len(self.d):
            return 0
        b = self.d[self.p]
        if b == 0xFF:
            b2 = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
            if b2 == 0:
                self.p += 2
                return 0xFF
            self.marker = True  # libjpeg feeds zero bits past a marker
            return 0
        self.p += 1
        return b

    def get(self, n: int) -> int:
        while self.cnt < n:
            self.buf = (self.buf << 8) | self._byte()
            self.cnt += 8
        self.cnt -= n
        v = (self.buf >> self.cnt) & ((1 << n) - 1)
        self.buf &= (1 << self.cnt) - 1
        return v

    def decode(self, h: _Huff) -> int:
        code = 0
        for length in range(1, 17):
            code = (code << 1) | self.get(1)
            if code <= h.maxcode[length]:
                return h.vals[code + h.valoff[length]]
        return 0

    def receive_extend(self, s: int) -> int:
        r = self.get(s)
        return r - (1 << s) + 1 if r < (1 << (s - 1)) else r

    def restart(self):
        self.buf = self.cnt = 0
        if self.p + 1 < len(self.d) and self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7:
            self.p += 2
        self.marker = False


def _idct_1d(v, shift):
    """jidctint.c jpeg_idct_islow, one 8-point pass on int64 arrays v[0..7] (broadcast over the other axis)."""
    z2, z3 = v[2], v[6]
    z1 = (z2 + z3) * 4433
    tmp2 = z1 + z3 * (-15137)
    tmp3 = z1 + z2 * 6270
    z2, z3 = v[0], v[4]
    tmp0 = (z2 + z3) << 13
    tmp1 = (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = v[7], v[5], v[3], v[1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * 9633
    tmp0, tmp1, tmp2, tmp3 = tmp0 * 2446, tmp1 * 16819, tmp2 * 25172, tmp3 * 12299
    z1, z2, z3, z4 = z1 * -7373, z2 * -20995, z3 * -16069 + z5, z4 * -3196 + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    rnd = 1 << (shift - 1)
    out = [tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3]
    return [(o + rnd) >> shift for o in out]


def idct_islow(blocks: np.ndarray) -> np.ndarray:
    """blocks: int [..., 8, 8] dequantised coefficients (row = vertical frequency) -> u8 [..., 8, 8]."""
    b = blocks.astype(np.int64)
    cols = np.stack(_idct_1d([b[..., r, :] for r in range(8)], 13 - 2), axis=-2)      # pass 1 along columns
    rows = np.stack(_idct_1d([cols[..., :, c] for c in range(8)], 13 + 2 + 3), axis=-1)  # pass 2 along rows
    m = rows & 1023  # sample_range_limit + CENTERJSAMPLE, indexed with RANGE_MASK
    return np.where(m < 128, m + 128, np.where(m < 512, 255, np.where(m < 896, 0, m - 896))).astype(np.uint8)


def jpeg_decode_gray(data: bytes) -> np.ndarray:
    """Luminance plane u8 [H, W] of a baseline (sequential Huffman, 8-bit) JPEG file."""
    assert data[:2] == b"\xff\xd8", "SOI expected"
    p = 2
    qt, dc, ac = {}, {}, {}
    comps, W, H, ri = [], 0, 0, 0
    while True:
        assert data[p] == 0xFF, "marker expected"
        while data[p + 1] == 0xFF:
            p += 1
        m, ln = data[p + 1], (data[p + 2] << 8) | data[p + 3]
        seg, end = p + 4, p + 2 + ln
        if m == 0xDB:
            while seg < end:
                pq, tq = data[seg] >> 4, data[seg] & 15
                seg += 1
                t = np.zeros(64, np.int64)
                for i in range(64):
                    t[ZIGZAG[i]] = (data[seg] << 8) | data[seg + 1] if pq else data[seg]
                    seg += 2 if pq else 1
                qt[tq] = t
        elif m == 0xC4:
            while seg < end:
                tc, th = data[seg] >> 4, data[seg] & 15
                bits = [0] + list(data[seg + 1:seg + 17])
                n = sum(bits)
                (ac if tc else dc)[th] = _Huff(bits, list(data[seg + 17:seg + 17 + n]))
                seg += 17 + n
        elif m in (0xC0, 0xC1):
            assert data[seg] == 8
            H, W = (data[seg + 1] << 8) | data[seg + 2], (data[seg + 3] << 8) | data[seg + 4]
            comps = [dict(h=data[seg + 7 + 3 * c] >> 4, v=data[seg + 7 + 3 * c] & 15, q=data[seg + 8 + 3 * c])
                     for c in range(data[seg + 5])]
        elif m == 0xC2 or (0xC5 <= m <= 0xCF and m not in (0xC8, 0xCC)):
            raise ValueError("progressive / lossless / arithmetic JPEG is outside the reference's data format")
        elif m == 0xDD:
            ri = (data[seg] << 8) | data[seg + 1]
        elif m == 0xDA:
            ns = data[seg]
            assert ns == len(comps), "one interleaved scan expected"
            for c in range(ns):
                comps[c]["dc"], comps[c]["ac"] = dc[data[seg + 2 + 2 * c] >> 4], ac[data[seg + 2 + 2 * c] & 15]
            p = end
            break
        p = end
    if len(comps) == 1:  # a single-component scan is never interleaved: one block per MCU whatever the sampling factors
        comps[0]["h"] = comps[0]["v"] = 1
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    mx_n, my_n = -(-W // (8 * hmax)), -(-H // (8 * vmax))
    bw, bh = mx_n * comps[0]["h"], my_n * comps[0]["v"]
    coef = np.zeros((bh, bw, 64), np.int64)
    br = _Bits(data, p)
    pred = [0] * len(comps)
    left = ri
    for my in range(my_n):
        for mx in range(mx_n):
            if ri:
                if left == 0:
                    br.restart()
                    pred = [0] * len(comps)
                    left = ri
                left -= 1
            for ci, c in enumerate(comps):
                for by in range(c["v"]):
                    for bx in range(c["h"]):
                        blk = np.zeros(64, np.int64)
                        s = br.decode(c["dc"])
                        if s:
                            pred[ci] += br.receive_extend(s)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = br.decode(c["ac"])
                            r, s = rs >> 4, rs & 15
                            if s:
                                k += r
                                blk[ZIGZAG[k & 63]] = br.receive_extend(s)
                                k += 1
                            elif r == 15:
                                k += 16
                            else:
                                break
                        if ci == 0:
                            coef[my * c["v"] + by, mx * c["h"] + bx] = blk
    pix = idct_islow((coef * qt[comps[0]["q"]]).reshape(bh, bw, 8, 8))
    return np.ascontiguousarray(pix.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)[:H, :W])


# ------------------------------------------------------------------------------------------------ transform ------
def video_transform(frames_u8, flip: bool, box, size, mask_span=None, mean: float = 0.421, std: float = 0.165):
    """frames_u8: torch u8 [T, H, W]; box = (top, left, h, w); size = (OH, OW); mask_span = (t0, t1) or None.
    Returns f32 [1, T, OH, OW] -- what Dataset.__getitem__ hands to the collate function (data.py:44-46, 68)."""
    import torch
    import torch.nn.functional as F

    x = frames_u8.unsqueeze(1).float() / 255.0            # data.py:158 FunctionalModule(lambda x: x / 255.0)
    if flip:
        x = x.flip(-1)                                    # RandomHorizontalFlip, one coin per clip (data.py:159)
    top, left, h, w = box
    x = x[..., top:top + h, left:left + w]                # resized_crop = crop + resize(bilinear, antialias=True)
    if (h, w) != tuple(size):
        x = F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=False, antialias=True)
    if mask_span is not None and mask_span[1] > mask_span[0]:
        x = x.clone()
        x[mask_span[0]:mask_span[1]] = x.mean()           # augment.py:141
    x = (x - mean) / std                                  # Normalize (data.py:163)
    return x.permute(1, 0, 2, 3).contiguous()
