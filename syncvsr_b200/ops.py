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


def pack_conv_weight_dgrad(w: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,R,S] -> [Cin, R*S*Cout] bf16, K index = (r*S+s)*Cout + co (operand B of the input-gradient conv)."""
    Cout, Cin, R, S = w.shape
    return w.permute(1, 2, 3, 0).reshape(Cin, R * S * Cout).to(torch.bfloat16).contiguous()


def conv2d_dgrad(dy: torch.Tensor, wd_packed: torch.Tensor, H: int, W: int, R: int, S: int, stride: int, pad: int,
                 resid: torch.Tensor | None = None, out_dtype=torch.bfloat16) -> torch.Tensor:
    _req(dy, torch.bfloat16, "dy"), _req(wd_packed, torch.bfloat16, "wd_packed")
    N, OH, OW, Cout = dy.shape
    Cin = wd_packed.shape[0]
    dx = torch.zeros(N, H, W, Cin, device=dy.device, dtype=out_dtype)
    rc = lib().svsr_conv2d_dgrad(ptr(dy), ptr(wd_packed), ptr(dx), ptr(resid), C.c_int(N), C.c_int(H), C.c_int(W),
                                 C.c_int(Cin), C.c_int(Cout), C.c_int(R), C.c_int(S), C.c_int(stride), C.c_int(pad),
                                 C.c_int(int(out_dtype == torch.float32)), stream_ptr())
    check(rc, "svsr_conv2d_dgrad")
    return dx


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, R: int, S: int, stride: int, pad: int,
                 out: torch.Tensor | None = None) -> torch.Tensor:
    """Returns dw as fp32 [R*S*Cin, Cout] (row (r*S+s)*Cin+ci); accumulates into `out` if given."""
    _req(x, torch.bfloat16, "x"), _req(dy, torch.bfloat16, "dy")
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    if out is None:
        out = torch.zeros(R * S * Cin, Cout, device=x.device, dtype=torch.float32)
    rc = lib().svsr_conv2d_wgrad(ptr(x), ptr(dy), ptr(out), C.c_int(N), C.c_int(H), C.c_int(W), C.c_int(Cin),
                                 C.c_int(Cout), C.c_int(R), C.c_int(S), C.c_int(stride), C.c_int(pad), stream_ptr())
    check(rc, "svsr_conv2d_wgrad")
    return out


def unpack_conv_wgrad(dw: torch.Tensor, Cin: int, R: int, S: int) -> torch.Tensor:
    """[R*S*Cin, Cout] -> torch layout [Cout,Cin,R,S]."""
    Cout = dw.shape[1]
    return dw.view(R, S, Cin, Cout).permute(3, 2, 0, 1).contiguous()


def gemm_wgrad(dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """dw[N,K] (fp32) += dy[M,N].T @ x[M,K]."""
    _req(dy, torch.bfloat16, "dy"), _req(x, torch.bfloat16, "x")
    M, N = dy.shape
    K = x.shape[1]
    if out is None:
        out = torch.zeros(N, K, device=x.device, dtype=torch.float32)
    rc = lib().svsr_gemm_wgrad(ptr(dy), C.c_int(N), ptr(x), C.c_int(K), ptr(out), C.c_int(K), C.c_int(M), C.c_int(N),
                               C.c_int(K), stream_ptr())
    check(rc, "svsr_gemm_wgrad")
    return out
