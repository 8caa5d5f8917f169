"""Developer micro-benchmark: tcgen05 implicit-GEMM kernel vs cuBLAS (torch.matmul) on the dense GEMM shapes of the
encoder / Conformer, warm L2, CUDA events, 200 iterations each."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from syncvsr_b200 import ops  # noqa: E402

SHAPES = [(1920, 1536, 512), (1920, 512, 512), (1920, 4096, 512), (1920, 512, 2048), (1856, 2560, 512),
          (2400, 2304, 768), (2400, 768, 768), (2400, 3072, 768), (2400, 768, 3072), (2400, 1536, 768),
          (2400, 5049, 768), (8192, 8192, 8192)]


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


for M, N, K in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out32 = torch.empty(M, N, device="cuda", dtype=torch.float32)
    n = 20 if M * N * K > 1e11 else 200
    t_mine = timeit(lambda: ops.gemm(a, b, out=out), n) if N % 8 == 0 else float("nan")
    t_mine32 = timeit(lambda: ops.gemm(a, b, out=out32), n) if N % 8 == 0 else float("nan")
    t_blas = timeit(lambda: torch.matmul(a, b.T, out=out), n)
    fl = 2.0 * M * N * K
    print(f"M={M:5d} N={N:5d} K={K:5d}  igemm bf16-out {t_mine:8.1f} us {fl/t_mine/1e6:7.0f} TF/s | fp32-out {t_mine32:8.1f} us |"
          f" cuBLAS {t_blas:8.1f} us {fl/t_blas/1e6:7.0f} TF/s")
