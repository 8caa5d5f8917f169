"""Warm per-launch time (wrapper allocations / memsets included: ~2 us) of the small encoder-side kernels at the LRW bench geometry (B=64: M = 64*30 rows, D = 512,
8 heads): CUDA events around 50 back-to-back launches (L2-resident operands, as in the step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncvsr_b200 import ops  # noqa: E402


def timeit(fn, n=20, reps=5):
    """n launches captured in one CUDA graph (the Python wrappers cost more host time than these kernels run for)."""
    fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3


def main():
    B, n, H, D = 64, 30, 8, 512
    M = B * n
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, D, device="cuda", generator=g)
    gam = torch.ones(D, device="cuda")
    dy = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    y, inv = ops.rmsnorm_fwd(x, gam)
    dx = torch.zeros(M, D, device="cuda")
    print(f"rmsnorm_fwd   {timeit(lambda: ops.rmsnorm_fwd(x, gam)):7.1f} us")
    print(f"rmsnorm_bwd   {timeit(lambda: ops.rmsnorm_bwd(dy, x, gam, inv, dx)):7.1f} us")
    qkv = torch.randn(M, 3 * H * 64, device="cuda", generator=g).bfloat16()
    rot = ops.rotary_table(n)
    d_o = torch.randn(M, H * 64, device="cuda", generator=g).bfloat16()
    print(f"attention_fwd {timeit(lambda: ops.attention_fwd(qkv, rot, B, n, H)):7.1f} us")
    print(f"attention_bwd {timeit(lambda: ops.attention_bwd(qkv, rot, d_o, B, n, H)):7.1f} us")
    h = torch.randn(M, 4096, device="cuda", generator=g).bfloat16()
    du = torch.randn(M, 2048, device="cuda", generator=g).bfloat16()
    print(f"geglu_fwd     {timeit(lambda: ops.geglu_fwd(h)):7.1f} us")
    print(f"geglu_bwd     {timeit(lambda: ops.geglu_bwd(h, du)):7.1f} us")


if __name__ == "__main__":
    main()
