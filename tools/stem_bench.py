"""Per-launch time of the stem's temporal 5-tap / 64 -> 64 column contraction (forward + fused BatchNorm statistics, and
its weight gradient) at the bench geometry [64 clips, 29 frames, 44*44 pixels, 64] with the temporal-halo kernels
(csrc/igemm_stem.cu, csrc/wgrad_stem.cu) on / off. CUDA events on the launching stream, 20 launches after 5 warm-ups,
tensors (460 MB each) larger than L2. Wrapper overhead (output allocation) is included on both sides."""
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
    N, T, W = 64, 29, 44 * 44
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, T, W, 64, device="cuda", generator=g).bfloat16()
    dz = (torch.randn(N, T, W, 64, device="cuda", generator=g) * 0.1).bfloat16()
    w = (torch.randn(64, 64, 5, 1, device="cuda", generator=g) * 0.05).bfloat16()
    wp = ops.pack_conv_weight(w)
    taps = [(kt - 2, 0) for kt in range(5)]
    out = torch.zeros(320, 64, device="cuda")
    flops = 2.0 * N * T * W * 64 * 320
    for mode in ("1", "0"):
        os.environ["SVSR_STEM_HALO"] = mode
        rows = [("fprop+bnstats", lambda: ops.conv_taps_fprop_bnstats(x, wp, taps), 2 * x.numel() * 2),
                ("wgrad", lambda: ops.conv_taps_wgrad(x, dz, taps, out=out), 2 * x.numel() * 2)]
        for name, fn, nbytes in rows:
            us = timeit(fn)
            print(f"stem_halo={mode} {name:14s} {us:8.1f} us  {flops / us * 1e-6:7.1f} TFLOP/s (K padded 245 -> 320)  "
                  f"{nbytes / us * 1e-3:7.1f} GB/s (distinct bytes)")


def direct():
    """The stem without the patch tensor (csrc/stem_direct.cu) against the patch tensor + halo kernels, launch by launch,
    from the fp32 video: patch gather | bf16 cast, forward contraction, weight gradient. Buffers preallocated."""
    import ctypes as C

    from syncvsr_b200._lib import check, lib, ptr, stream_ptr

    B, T, S = 64, 29, 88
    npix = (S // 2) ** 2
    g = torch.Generator(device="cuda").manual_seed(1)
    v = torch.randn(B, 1, T, S, S, device="cuda", generator=g)
    wp = (torch.randn(64, 320, device="cuda", generator=g) * 0.05).bfloat16()
    dz = (torch.randn(B, T, npix, 64, device="cuda", generator=g) * 0.1).bfloat16()
    P = torch.empty(B, T, npix, 64, device="cuda", dtype=torch.bfloat16)
    vb = torch.empty(B, T, S, S, device="cuda", dtype=torch.bfloat16)
    y = torch.empty(B, T, npix, 64, device="cuda", dtype=torch.bfloat16)
    st = torch.zeros(2, 64, device="cuda", dtype=torch.float64)
    out = torch.zeros(320, 64, device="cuda")
    L, i = lib(), C.c_int
    taps = [(kt - 2, 0) for kt in range(5)]
    os.environ["SVSR_STEM_HALO"] = "1"
    rows = [
        ("patch gather (fp32 video -> 460 MB patch tensor)",
         lambda: check(L.svsr_stem_patch(ptr(v), ptr(P), i(B), i(T), i(S), i(S), stream_ptr()), "patch")),
        ("forward, patch tensor by TMA", lambda: ops.conv_taps_fprop_bnstats(P, wp, taps)),
        ("weight gradient, patch tensor by TMA", lambda: ops.conv_taps_wgrad(P, dz, taps, out=out)),
        ("bf16 copy of the video", lambda: vb.copy_(v.view(B, T, S, S))),
        ("forward, window rows built in shared memory",
         lambda: check(L.svsr_stem_conv_direct(ptr(vb), ptr(wp), ptr(y), ptr(st), i(B), i(T), i(S), i(S), stream_ptr()), "fwd")),
        ("weight gradient, window rows built in shared memory",
         lambda: check(L.svsr_stem_wgrad_direct(ptr(vb), ptr(dz), ptr(out), i(64), i(B), i(T), i(S), i(S), stream_ptr()), "wg")),
    ]
    for name, fn in rows:
        print(f"{name:52s} {timeit(fn):8.1f} us")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "direct":
        direct()
    else:
        main()
