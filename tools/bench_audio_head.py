"""BASELINE.json configs[4] -- audio-head stress: seq_len=400, alignment=4, vq_groups=8, vocab=1024 (A*G*V = 32768 logits
per frame). Times the head as the native path runs it today (tcgen05 projection GEMM -> fp32 logits -> one-pass CE that
also emits the bf16 logits gradient -> input-gradient GEMM -> weight-gradient GEMM) and reports achieved bandwidth
against (a) the algorithmic bytes of a fully fused head (X + W + tokens + dX + dW) and (b) the bytes
this unfused pipeline actually has to move (adds the fp32 logits written+read and the bf16 gradient written+read x2)."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from syncvsr_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--H", type=int, default=768)
    ap.add_argument("--T", type=int, default=400)
    a = ap.parse_args()
    B, T, H, A, G, V = a.B, a.T, a.H, 4, 8, 1024
    N = A * G * V
    M = B * T
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, H, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, H, device="cuda", generator=g) * 0.02).bfloat16()
    wt = w.t().contiguous()
    bias = torch.zeros(N, device="cuda")
    tokens = torch.randint(0, V, (B, T * A, G), device="cuda", generator=g)
    logits = torch.empty(M, N, device="cuda", dtype=torch.float32)
    dx = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(N, H, device="cuda")
    rows = M * A * G

    def step():
        ops.gemm(x, w, bias=bias, out=logits)
        loss, dl, _bad = ops.audio_ce(logits, tokens, T, A, G, V, dscale=1.0 / rows)
        ops.gemm(dl, wt, out=dx)
        ops.gemm_wgrad(dl, x, out=dw)
        return loss

    for _ in range(3):
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fused_bytes = M * H * 2 + N * H * 2 + tokens.numel() * 8 + M * H * 2 + N * H * 4
    actual_bytes = fused_bytes + M * N * 4 * 2 + M * N * 2 * 3
    flops = 3 * 2.0 * M * N * H
    print(json.dumps({"workload": f"audio head stress B={B} T={T} H={H} A*G*V={N}", "ms": ms,
                      "loss_per_row": float(loss), "tflops": flops / ms / 1e9,
                      "algorithmic_fused_MB": fused_bytes / 1e6, "fused_GBps": fused_bytes / ms / 1e6,
                      "unfused_pipeline_MB": actual_bytes / 1e6, "unfused_GBps": actual_bytes / ms / 1e6,
                      "frames_per_s": M / ms * 1e3}))


if __name__ == "__main__":
    main()
