"""Split-K sweep of the weight-gradient kernel on the LRW trunk / encoder shapes (B=64: 1856 images): per-launch time for
the cost model's choice and for forced split factors (SVSR_WGRAD_NSPLIT), CUDA events, 10 launches after 3 warm-ups."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncvsr_b200 import ops  # noqa: E402


def timeit(fn, n=10, warm=3):
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
    N = 64 * 29
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()  # noqa: E731
    convs = [(22, 64, 64, 3, 1), (22, 64, 128, 3, 2), (11, 128, 128, 3, 1), (11, 128, 256, 3, 2), (6, 256, 256, 3, 1),
             (6, 256, 512, 3, 2), (3, 512, 512, 3, 1), (22, 64, 128, 1, 2)]
    sweeps = [1, 2, 4, 8, 16, 32, 64, 148, 296]
    for (S, Cin, Cout, R, stride) in convs:
        pad = R // 2
        OS = (S + 2 * pad - R) // stride + 1
        x, dy = rn(N, S, S, Cin), rn(N, OS, OS, Cout)
        out = torch.zeros(R * R * Cin, Cout, device="cuda")
        fn = lambda: ops.conv2d_wgrad(x, dy, R, R, stride, pad, out=out)  # noqa: E731
        os.environ.pop("SVSR_WGRAD_NSPLIT", None)
        base = timeit(fn)
        os.environ["SVSR_WGRAD_SPLIT_LEGACY"] = "1"
        legacy = timeit(fn)
        os.environ.pop("SVSR_WGRAD_SPLIT_LEGACY")
        res = []
        for n in sweeps:
            os.environ["SVSR_WGRAD_NSPLIT"] = str(n)
            res.append((n, timeit(fn)))
        os.environ.pop("SVSR_WGRAD_NSPLIT", None)
        fl = 2.0 * N * OS * OS * R * R * Cin * Cout
        print(f"conv {S}x{S} {Cin}->{Cout} k{R} s{stride}: model {base:7.1f} us ({fl / base * 1e-6:6.0f} TF/s) legacy {legacy:7.1f} | " +
              " ".join(f"n{n}:{t:.0f}" for n, t in res), flush=True)
    M = 64 * 30
    for (Nout, K) in [(512, 512), (1536, 512), (4096, 512), (512, 2048)]:
        dy, x = rn(M, Nout), rn(M, K)
        acc = torch.zeros(Nout, K, device="cuda")
        fn = lambda: ops.gemm_wgrad(dy, x, out=acc)  # noqa: E731
        os.environ.pop("SVSR_WGRAD_NSPLIT", None)
        base = timeit(fn)
        res = []
        for n in [1, 2, 4, 8, 15]:
            os.environ["SVSR_WGRAD_NSPLIT"] = str(n)
            res.append((n, timeit(fn)))
        os.environ.pop("SVSR_WGRAD_NSPLIT", None)
        print(f"linear wgrad [{Nout},{K}] over {M} rows: model {base:6.1f} us | " + " ".join(f"n{n}:{t:.0f}" for n, t in res), flush=True)


if __name__ == "__main__":
    main()
