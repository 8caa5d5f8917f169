"""One dense GEMM shape through the tcgen05 implicit-GEMM kernel, a few launches (for ncu --set full): gemm_one.py M N K"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from syncvsr_b200 import ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(6):
    ops.gemm(a, b, out=out)
torch.cuda.synchronize()
