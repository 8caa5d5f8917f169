"""Thin Python wrappers over the C ABI: torch tensors in, raw pointers across the boundary.

These are the per-kernel entry points the parity tests call; the model-level path (lightning.py)
drives the same kernels through the native step executor."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib, ptr, stream_ptr


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def gemm(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor | None = None, resid: torch.Tensor | None = None,
         out_dtype=torch.bfloat16, alpha: float = 1.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """out[M,N] = alpha * a[M,K] @ b[N,K].T (+bias) (+resid)."""
    _req(a, torch.bfloat16, "a"), _req(b, torch.bfloat16, "b")
    M, K = a.shape
    N, K2 = b.shape
    assert K == K2
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    if bias is not None:
        _req(bias, torch.float32, "bias")
    rc = lib().svsr_gemm_bf16(ptr(a), C.c_int(K), ptr(b), C.c_int(K), ptr(out), C.c_int(N), ptr(bias), ptr(resid),
                              C.c_int(M), C.c_int(N), C.c_int(K), C.c_int(int(out.dtype == torch.float32)),
                              C.c_int(int(resid is not None and resid.dtype == torch.float32)), C.c_float(alpha),
                              stream_ptr())
    check(rc, "svsr_gemm_bf16")
    return out


def conv2d_fprop(x: torch.Tensor, w_packed: torch.Tensor, R: int, S: int, stride: int, pad: int,
                 resid: torch.Tensor | None = None, out_dtype=torch.bfloat16) -> torch.Tensor:
    """x: [N,H,W,Cin] bf16 NHWC; w_packed: [Cout, R*S*Cin] bf16 -> y: [N,OH,OW,Cout]."""
    _req(x, torch.bfloat16, "x"), _req(w_packed, torch.bfloat16, "w_packed")
    N, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    assert w_packed.shape[1] == R * S * Cin
    OH = (H + 2 * pad - R) // stride + 1
    OW = (W + 2 * pad - S) // stride + 1
    y = torch.empty(N, OH, OW, Cout, device=x.device, dtype=out_dtype)
    rc = lib().svsr_conv2d_fprop(ptr(x), ptr(w_packed), ptr(y), ptr(resid), C.c_int(N), C.c_int(H), C.c_int(W),
                                 C.c_int(Cin), C.c_int(Cout), C.c_int(R), C.c_int(S), C.c_int(stride), C.c_int(pad),
                                 C.c_int(int(out_dtype == torch.float32)), stream_ptr())
    check(rc, "svsr_conv2d_fprop")
    return y


def pack_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,R,S] (torch layout) -> [Cout, R*S*Cin] bf16, K index = (r*S+s)*Cin + c."""
    Cout, Cin, R, S = w.shape
    return w.permute(0, 2, 3, 1).reshape(Cout, R * S * Cin).to(torch.bfloat16).contiguous()
