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


def case_dgrad(N, H, W, Cin, Cout, R, stride, pad, resid=False):
    import torch
    import torch.nn.functional as F
    from syncvsr_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(3)
    OH = (H + 2 * pad - R) // stride + 1
    dy = torch.randn(N, OH, OH, Cout, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, R, R, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    rs = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16) if resid else None
    wd = ops.pack_conv_weight_dgrad(w)
    dx = ops.conv2d_dgrad(dy, wd, H, W, R, R, stride, pad, resid=rs)
    torch.cuda.synchronize()
    x = torch.zeros(N, Cin, H, W, device="cuda", requires_grad=True)
    y = F.conv2d(x, w.float(), stride=stride, padding=pad)
    (ref,) = torch.autograd.grad(y, x, dy.float().permute(0, 3, 1, 2))
    ref = ref.permute(0, 2, 3, 1)
    if resid:
        ref = ref + rs.float()
    rel, mx = _rel_err(dx, ref)
    ms = _time(lambda: ops.conv2d_dgrad(dy, wd, H, W, R, R, stride, pad, resid=rs))
    tf = 2.0 * N * OH * OH * Cout * Cin * R * R / ms / 1e9
    print(f"dgrad N={N} {H}x{W} Cin={Cin} Cout={Cout} k={R} s={stride} p={pad} resid={resid}: rel={rel:.3e} "
          f"max={mx:.3e} {ms*1e3:.1f}us {tf:.1f}TF/s {'OK' if rel < 1e-2 else 'FAIL'}")


def case_wgrad(N, H, W, Cin, Cout, R, stride, pad):
    import torch
    import torch.nn.functional as F
    from syncvsr_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(4)
    OH = (H + 2 * pad - R) // stride + 1
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
    dy = (torch.randn(N, OH, OH, Cout, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    dw = ops.unpack_conv_wgrad(ops.conv2d_wgrad(x, dy, R, R, stride, pad), Cin, R, R)
    torch.cuda.synchronize()
    w = torch.zeros(Cout, Cin, R, R, device="cuda", requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, stride=stride, padding=pad)
    (ref,) = torch.autograd.grad(y, w, dy.float().permute(0, 3, 1, 2))
    rel, mx = _rel_err(dw, ref)
    out = torch.zeros(R * R * Cin, Cout, device="cuda")
    ms = _time(lambda: ops.conv2d_wgrad(x, dy, R, R, stride, pad, out=out))
    tf = 2.0 * N * OH * OH * Cout * Cin * R * R / ms / 1e9
    print(f"wgrad N={N} {H}x{W} Cin={Cin} Cout={Cout} k={R} s={stride} p={pad}: rel={rel:.3e} max={mx:.3e} "
          f"{ms*1e3:.1f}us {tf:.1f}TF/s {'OK' if rel < 1e-3 else 'FAIL'}")


def case_gemm_wgrad(M, N, K):
    import torch
    from syncvsr_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    dy = (torch.randn(M, N, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    dw = ops.gemm_wgrad(dy, x)
    torch.cuda.synchronize()
    ref = dy.float().T @ x.float()
    rel, mx = _rel_err(dw, ref)
    out = torch.zeros(N, K, device="cuda")
    ms = _time(lambda: ops.gemm_wgrad(dy, x, out=out))
    tf = 2.0 * M * N * K / ms / 1e9
    print(f"gemm_wgrad M={M} N={N} K={K}: rel={rel:.3e} max={mx:.3e} {ms*1e3:.1f}us {tf:.1f}TF/s "
          f"{'OK' if rel < 1e-3 else 'FAIL'}")


def case_rowshift():
    import ctypes as C
    import torch
    import subprocess
    from pathlib import Path
    from syncvsr_b200._lib import check, lib, ptr, stream_ptr

    # the probe lives outside the product library: build it on demand against libsvsr.so (tensor-map + error helpers)
    root = Path(__file__).resolve().parent
    out_dir = root / "probes" / "_build"
    out_dir.mkdir(parents=True, exist_ok=True)
    so = out_dir / "libprobe.so"
    pkg = root.parent / "syncvsr_b200"
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                    "-lineinfo", "-Xcompiler", "-fPIC", "-shared", "-o", str(so), str(root / "probes" / "debug_probe.cu"),
                    f"-L{pkg}", "-lsvsr", f"-Xlinker=-rpath={pkg}", "-lcuda"], check=True)
    lib()  # libsvsr.so first (RTLD_GLOBAL not needed: the probe links against it)
    probe = C.CDLL(str(so))

    g = torch.Generator(device="cuda").manual_seed(6)
    res = []
    for mode in (0, 2, 1, 3):
        mn = mode & 1
        a = torch.randn(256, 128 if mn else 64, device="cuda", generator=g).to(torch.bfloat16)
        b = torch.randn(256 if mn else 64, 64, device="cuda", generator=g).to(torch.bfloat16)
        for shift in (0, 8, 1, 3, 13, 50):
            out = torch.zeros(128, 64, device="cuda")
            check(probe.svsr_debug_rowshift(ptr(a), ptr(b), ptr(out), C.c_int(shift), C.c_int(mode), stream_ptr()),
                  "rowshift")
            torch.cuda.synchronize()
            if mn:
                ref = a[shift:shift + 128].float().T @ b[shift:shift + 128].float()
            else:
                ref = a[shift:shift + 128].float() @ b.float().T
            rel, _ = _rel_err(out, ref)
            res.append(f"m{mode}s{shift}:{rel:.1e}")
    print("rowshift " + " ".join(res))


CASES = {
    "rowshift": case_rowshift,
    "dgrad_l1": lambda: case_dgrad(58, 22, 22, 64, 64, 3, 1, 1),
    "dgrad_l1r": lambda: case_dgrad(58, 22, 22, 64, 64, 3, 1, 1, resid=True),
    "dgrad_l2s2": lambda: case_dgrad(58, 22, 22, 64, 128, 3, 2, 1),
    "dgrad_l3s2": lambda: case_dgrad(58, 11, 11, 128, 256, 3, 2, 1),
    "dgrad_ds": lambda: case_dgrad(58, 22, 22, 64, 128, 1, 2, 0),
    "dgrad_ds3": lambda: case_dgrad(58, 11, 11, 128, 256, 1, 2, 0),
    "wgrad_l1": lambda: case_wgrad(58, 22, 22, 64, 64, 3, 1, 1),
    "wgrad_l2": lambda: case_wgrad(58, 11, 11, 128, 128, 3, 1, 1),
    "wgrad_l3": lambda: case_wgrad(58, 6, 6, 256, 256, 3, 1, 1),
    "wgrad_l4": lambda: case_wgrad(58, 3, 3, 512, 512, 3, 1, 1),
    "wgrad_l2s2": lambda: case_wgrad(58, 22, 22, 64, 128, 3, 2, 1),
    "wgrad_l3s2": lambda: case_wgrad(58, 11, 11, 128, 256, 3, 2, 1),
    "wgrad_ds": lambda: case_wgrad(58, 22, 22, 64, 128, 1, 2, 0),
    "gemm_wgrad1": lambda: case_gemm_wgrad(1920, 512, 512),
    "gemm_wgrad2": lambda: case_gemm_wgrad(1920, 4096, 512),
    "gemm_wgrad3": lambda: case_gemm_wgrad(1856, 2560, 512),
    "gemm_wgrad4": lambda: case_gemm_wgrad(1920, 512, 2048),
    "wgrad_l1_big": lambda: case_wgrad(1856, 22, 22, 64, 64, 3, 1, 1),
    "wgrad_l2_big": lambda: case_wgrad(1856, 11, 11, 128, 128, 3, 1, 1),
    "wgrad_l3_big": lambda: case_wgrad(1856, 6, 6, 256, 256, 3, 1, 1),
    "wgrad_l4_big": lambda: case_wgrad(1856, 3, 3, 512, 512, 3, 1, 1),
    "dgrad_l1_big": lambda: case_dgrad(1856, 22, 22, 64, 64, 3, 1, 1),
    "dgrad_l2s2_big": lambda: case_dgrad(1856, 22, 22, 64, 128, 3, 2, 1),
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
    if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
        for name in sys.argv[1:]:
            CASES[name]()
        return
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    lines = []
    names = list(CASES)
    if sys.argv[1:2] == ["--prefix"]:
        names = [n for n in names if any(n.startswith(p) for p in sys.argv[2:])]
    for name in names:
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
