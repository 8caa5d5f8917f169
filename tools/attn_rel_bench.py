"""The d_k = 64 attention core of the LRS path at the bench geometries (C3: B=16, T=150; C4: B=8, T=250; H=12), forward and
backward, tcgen05 kernels (attention_rel_tc.cu) vs the CUDA-core fp32 kernels (SVSR_ATTN_TC=0). Run under
`ncu --metrics gpu__time_duration.sum -k regex:attn_rel\\|attention_core` for per-kernel times; alone it prints CUDA-event
times of the fwd / bwd wrappers (output allocations included)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncvsr_b200 import ops  # noqa: E402


def run(B, T, H, tc, reps):
    os.environ["SVSR_ATTN_TC"] = "1" if tc else "0"
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = (0.7 * torch.randn(B * T, 3 * D, device="cuda", generator=g)).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    pos = (0.7 * torch.randn(2 * T - 1, D, device="cuda", generator=g)).bfloat16()
    bu = 0.3 * torch.randn(H, 64, device="cuda", generator=g)
    bv = 0.3 * torch.randn(H, 64, device="cuda", generator=g)
    klen = torch.full((B,), T, dtype=torch.int32, device="cuda")
    d_o = (0.5 * torch.randn(B * T, D, device="cuda", generator=g)).bfloat16()
    o, lse = ops.attention_core_fwd(q, k, v, B, H, T, T, p=pos, bias_u=bu, bias_v=bv, klen=klen)
    ops.attention_core_bwd(q, k, v, o, lse, d_o, B, H, T, T, p=pos, bias_u=bu, bias_v=bv, klen=klen)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(reps):
        ops.attention_core_fwd(q, k, v, B, H, T, T, p=pos, bias_u=bu, bias_v=bv, klen=klen)
    ev[1].record()
    for _ in range(reps):
        ops.attention_core_bwd(q, k, v, o, lse, d_o, B, H, T, T, p=pos, bias_u=bu, bias_v=bv, klen=klen)
    ev[2].record()
    torch.cuda.synchronize()
    flops_f = 2.0 * B * H * T * T * 64 * 3  # QK^T, QP^T (useful band), PV
    print(f"B={B} T={T} H={H} {'tcgen05 ' if tc else 'cudacore'}: fwd {ev[0].elapsed_time(ev[1]) / reps * 1e3:7.1f} us  "
          f"bwd {ev[1].elapsed_time(ev[2]) / reps * 1e3:7.1f} us  (fwd useful {flops_f / 1e9:.2f} GF)")


if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    for B, T in ((16, 150), (8, 250)):
        for tc in (True, False):
            run(B, T, 12, tc, reps)
