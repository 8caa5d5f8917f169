"""Developer GPU probe: runs each kernel case in its own subprocess (a trapped kernel poisons the CUDA
context) with a timeout, and writes a summary to gpurun_out/gpu_check.txt.

    python tools/gpu_check.py            # run all cases
    python tools/gpu_check.py CASE       # run one case in-process
"""
from __future__ import annotations

import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _rel_err(a, b):
    import torch

    a = a.float()
    b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), (a - b).abs().max().item()


def _time(fn, iters=20, warm=3):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case_gemm(M, N, K, fp32_out=False, bias=False, resid=False):
    import torch
    from syncvsr_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bi = torch.randn(N, device="cuda", generator=g) if bias else None
    rs = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if resid else None
    out = ops.gemm(a, b, bias=bi, resid=rs, out_dtype=torch.float32 if fp32_out else torch.bfloat16)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().T
    if bias:
        ref = ref + bi
    if resid:
        ref = ref + rs.float()
    rel, mx = _rel_err(out, ref)
    ms = _time(lambda: ops.gemm(a, b, bias=bi, resid=rs, out_dtype=torch.float32 if fp32_out else torch.bfloat16))
    tf = 2.0 * M * N * K / ms / 1e9
    print(f"gemm M={M} N={N} K={K} fp32={fp32_out} bias={bias} resid={resid}: rel={rel:.3e} max={mx:.3e} "
          f"{ms*1e3:.1f}us {tf:.1f}TF/s {'OK' if rel < 1e-2 else 'FAIL'}")


def case_conv(N, H, W, Cin, Cout, R, stride, pad):
    import torch
    import torch.nn.functional as F
    from syncvsr_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, R, R, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    wp = ops.pack_conv_weight(w)
    y = ops.conv2d_fprop(x, wp, R, R, stride, pad)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    rel, mx = _rel_err(y, ref)
    ms = _time(lambda: ops.conv2d_fprop(x, wp, R, R, stride, pad))
    OH = ref.shape[1]
    tf = 2.0 * N * OH * OH * Cout * Cin * R * R / ms / 1e9
    print(f"conv N={N} {H}x{W} Cin={Cin} Cout={Cout} k={R} s={stride} p={pad}: rel={rel:.3e} max={mx:.3e} "
          f"{ms*1e3:.1f}us {tf:.1f}TF/s {'OK' if rel < 1e-2 else 'FAIL'}")


CASES = {
    "gemm_small": lambda: case_gemm(128, 64, 64),
    "gemm_k512": lambda: case_gemm(256, 128, 512),
    "gemm_n256": lambda: case_gemm(1920, 512, 512, bias=True),
    "gemm_fp32": lambda: case_gemm(1920, 2560, 512, fp32_out=True, bias=True),
    "gemm_resid": lambda: case_gemm(1920, 512, 2048, resid=True),
    "gemm_ragged": lambda: case_gemm(100, 500, 512, fp32_out=True, bias=True),
    "gemm_big": lambda: case_gemm(8192, 4096, 4096),
    "conv_l1": lambda: case_conv(58, 22, 22, 64, 64, 3, 1, 1),
    "conv_l2": lambda: case_conv(58, 11, 11, 128, 128, 3, 1, 1),
    "conv_l3": lambda: case_conv(58, 6, 6, 256, 256, 3, 1, 1),
    "conv_l4": lambda: case_conv(58, 3, 3, 512, 512, 3, 1, 1),
    "conv_l2s2": lambda: case_conv(58, 22, 22, 64, 128, 3, 2, 1),
    "conv_l3s2": lambda: case_conv(58, 11, 11, 128, 256, 3, 2, 1),
    "conv_ds": lambda: case_conv(58, 22, 22, 64, 128, 1, 2, 0),
    "conv_l1_big": lambda: case_conv(1856, 22, 22, 64, 64, 3, 1, 1),
    "conv_l2_big": lambda: case_conv(1856, 11, 11, 128, 128, 3, 1, 1),
    "conv_l3_big": lambda: case_conv(1856, 6, 6, 256, 256, 3, 1, 1),
    "conv_l4_big": lambda: case_conv(1856, 3, 3, 512, 512, 3, 1, 1),
}


def main():
    if len(sys.argv) > 1 and sys.argv[1] != "--all":
        for name in sys.argv[1:]:
            CASES[name]()
        return
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    lines = []
    for name in CASES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=180)
            tail = (r.stdout.strip().splitlines() or ["<no stdout>"])[-1]
            if r.returncode != 0:
                err = " | ".join(r.stderr.strip().splitlines()[-4:])
                tail = f"{name}: EXIT {r.returncode}: {tail} :: {err}"
        except subprocess.TimeoutExpired:
            tail = f"{name}: TIMEOUT"
        line = f"[{time.time()-t0:5.1f}s] {tail}"
        print(line, flush=True)
        lines.append(line)
    (out_dir / "gpu_check.txt").write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
