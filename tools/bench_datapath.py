"""Data path timing: one LRW batch (B=64 clips x T=29 frames, 96x112 4:2:2 q85 JPEG, the format
preprocess_pkl.py:182 writes) -> model input f32 [64,1,29,96,96].
  GPU arm: JpegBatchDecoder.decode + VideoTransform (host marker parse and H2D of the JPEG bytes INSIDE the timed region),
           CUDA events on the launching stream; kernel-only numbers are timed separately with inputs resident.
  CPU arm: what the reference's DataLoader workers do per frame -- libjpeg-turbo decode in grayscale mode (OpenCV here;
           PyTurboJPEG is not installed) + the oracle transform (torch CPU ops), single thread, on a bounded sample.
Prints one JSON line."""
import io
import json
import os
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncvsr_b200 import data as D  # noqa: E402
from syncvsr_b200._lib import lib, ptr, stream_ptr, check  # noqa: E402


def synth_files(n_unique=58):
    from PIL import Image

    rng = np.random.default_rng(0)
    files = []
    for i in range(n_unique):
        base = rng.normal(size=(26, 30, 3))
        img = np.kron(base, np.ones((4, 4, 1)))[:96, :112] * 40 + 110 + rng.normal(size=(96, 112, 3)) * 6
        buf = io.BytesIO()
        Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(buf, format="JPEG", quality=85, subsampling=1)
        files.append(buf.getvalue())
    return files


def ev_time(fn, n, warm):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B, T = 64, 29
    uniq = synth_files()
    files = [uniq[i % len(uniq)] for i in range(B * T)]
    dec, vt = D.JpegBatchDecoder(), D.VideoTransform((96, 96), train=True)
    torch.manual_seed(0), random.seed(0)
    plan = vt.plan(B, T, 96, 112)

    def full():
        fr = dec.decode(files)
        return vt(fr.view(B, T, 96, 112), plan)

    ms_full = ev_time(full, 10, 3)
    # kernel-only legs, inputs resident
    p = dec.parse(files)
    t0 = time.perf_counter()
    for _ in range(5):
        dec.parse(files)
    ms_parse = (time.perf_counter() - t0) / 5 * 1e3
    dev = torch.device("cuda")
    blob, desc = torch.from_numpy(p["blob"].copy()).to(dev), torch.from_numpy(p["desc"]).to(dev)
    qt, ht = torch.from_numpy(p["qt"].view(np.int16)).to(dev), torch.from_numpy(p["ht"]).to(dev)
    n, bw, bh = p["n"], p["bw"], p["bh"]
    coef = torch.empty(n * bw * bh * 64, device=dev, dtype=torch.int16)
    out = torch.empty(n, 96, 112, device=dev, dtype=torch.uint8)
    import ctypes as C

    def dec_only():
        check(lib().svsr_jpeg_decode_gray(ptr(blob), ptr(desc), C.c_int(n), ptr(qt), ptr(ht), ptr(coef), ptr(out), C.c_int(112),
                                          C.c_int(96), C.c_int(bw), C.c_int(bh), stream_ptr()), "decode")

    ms_dec = ev_time(dec_only, 10, 3)
    frames = out.view(B, T, 96, 112)
    ms_tf = ev_time(lambda: vt(frames, plan), 10, 3)

    # CPU arm: bounded sample of 4 clips
    import cv2

    from oracle import data_oracle as do

    torch.set_num_threads(1)
    nclip = 4
    t0 = time.perf_counter()
    for b in range(nclip):
        fr = np.stack([cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_GRAYSCALE) for f in files[b * T:(b + 1) * T]])
        flip, top, left, h, w, m0, m1, _ = plan[0][b].tolist()
        do.video_transform(torch.from_numpy(fr), bool(flip), (top, left, h, w), plan[1], (m0, m1))
    cpu_ms_per_clip = (time.perf_counter() - t0) / nclip * 1e3
    jpeg_bytes = sum(len(f) for f in files)
    print(json.dumps({
        "workload": "LRW batch 64x29 frames, 96x112 4:2:2 q85 JPEG -> f32 [64,1,29,96,96] (flip + RRC + TimeMask + Normalize)",
        "gpu_ms_per_batch_e2e": ms_full, "gpu_clips_per_s_e2e": B / ms_full * 1e3,
        "host_parse_ms": ms_parse, "jpeg_decode_kernels_ms": ms_dec, "transform_kernel_ms": ms_tf,
        "jpeg_bytes_per_batch": jpeg_bytes, "f32_bytes_per_batch": B * T * 96 * 96 * 4,
        "transform_GBps": (B * T * 96 * 112 + B * T * 96 * 96 * 4) / ms_tf * 1e-6,
        "cpu_ms_per_clip_1thread": cpu_ms_per_clip, "cpu_clips_per_s_1thread": 1e3 / cpu_ms_per_clip,
        "cpu_sample": f"{nclip} clips, OpenCV/libjpeg-turbo grayscale decode + oracle transform, 1 thread"}))


if __name__ == "__main__":
    main()
