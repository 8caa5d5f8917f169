"""Per-launch time of the 64 -> 64 channel 3x3 convolution (resnet.layer1 geometry) with the halo-tile kernel on / off.
CUDA events on the launching stream, 20 launches after 5 warm-ups, inputs (131 MB per tensor) larger than L2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncvsr_b200 import ops  # noqa: E402


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    N, H, W = 64 * 29, 22, 22
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, H, W, 64, device="cuda", generator=g).bfloat16()
    dy = torch.randn(N, H, W, 64, device="cuda", generator=g).bfloat16()
    base = torch.randn(N, H, W, 64, device="cuda", generator=g).bfloat16()
    w = (torch.randn(64, 64, 3, 3, device="cuda", generator=g) * 0.05).bfloat16()
    wp, wd = ops.pack_conv_weight(w), ops.pack_conv_weight_dgrad(w)
    flops = 2.0 * N * H * W * 64 * 576
    for mode in ("1", "0"):
        os.environ["SVSR_HALO_CONV"] = mode
        rows = [("fprop", lambda: ops.conv2d_fprop(x, wp, 3, 3, 1, 1)),
                ("fprop+bnstats", lambda: ops.conv2d_fprop_bnstats(x, wp, 3, 3, 1, 1)),
                ("dgrad", lambda: ops.conv2d_dgrad(dy, wd, H, W, 3, 3, 1, 1)),
                ("dgrad+resid", lambda: ops.conv2d_dgrad(dy, wd, H, W, 3, 3, 1, 1, resid=base))]
        for name, fn in rows:
            us = timeit(fn)
            print(f"halo={mode} {name:14s} {us:8.1f} us  {flops / us * 1e-6:7.1f} TFLOP/s  "
                  f"{(2 * x.numel() * 2) / us * 1e-3:7.1f} GB/s (in+out)")


if __name__ == "__main__":
    main()
