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


# ---------------------------------------------------------------------------------------------------------------
# non-GEMM operators (thin wrappers; scratch buffers allocated here with torch)
# ---------------------------------------------------------------------------------------------------------------
def _i(v):
    return C.c_int(int(v))


def stem_patch(videos: torch.Tensor) -> torch.Tensor:
    _req(videos, torch.float32, "videos")
    B, _, T, H, W = videos.shape
    OH, OW = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    out = torch.empty(B, T, OH * OW, 64, device=videos.device, dtype=torch.bfloat16)
    check(lib().svsr_stem_patch(ptr(videos), ptr(out), _i(B), _i(T), _i(H), _i(W), stream_ptr()), "svsr_stem_patch")
    return out


def batchnorm_fwd(x, gamma, beta, running_mean, running_var, train=True, res=None, res_coef=None, relu=False,
                  eps=1e-5, momentum=0.1):
    """x: bf16 [..., C] channels-last. Returns (out bf16, coef fp32 [4, C])."""
    _req(x, torch.bfloat16, "x")
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    out = torch.empty_like(x)
    coef = torch.empty(4, Cc, device=x.device, dtype=torch.float32)
    scratch = torch.zeros(2 * Cc, device=x.device, dtype=torch.float64)
    check(lib().svsr_batchnorm_fwd(ptr(x), C.c_int64(rows), _i(Cc), ptr(gamma), ptr(beta), ptr(running_mean),
                                   ptr(running_var), C.c_float(eps), C.c_float(momentum), _i(train), ptr(res),
                                   ptr(res_coef), _i(relu), ptr(out), ptr(coef), ptr(scratch), stream_ptr()),
          "svsr_batchnorm_fwd")
    return out, coef


def batchnorm_bwd(dout, relu_ref, c, coef, want_gmask=False):
    """Returns (dc bf16, dgamma, dbeta, gmask|None)."""
    Cc = c.shape[-1]
    rows = c.numel() // Cc
    dc = torch.empty_like(c)
    dgamma = torch.zeros(Cc, device=c.device)
    dbeta = torch.zeros(Cc, device=c.device)
    gm = torch.empty_like(c) if want_gmask else None
    scratch = torch.zeros(2 * Cc, device=c.device, dtype=torch.float64)
    kcoef = torch.empty(2 * Cc, device=c.device)
    check(lib().svsr_batchnorm_bwd(ptr(dout), ptr(relu_ref), ptr(c), C.c_int64(rows), _i(Cc), ptr(coef), ptr(dgamma),
                                   ptr(dbeta), ptr(dc), ptr(gm), ptr(scratch), ptr(kcoef), stream_ptr()),
          "svsr_batchnorm_bwd")
    return dc, dgamma, dbeta, gm


def stem_bn_gelu_pool_fwd(y0, coef):
    N, IH, IW, Cc = y0.shape
    assert Cc == 64
    OH, OW = (IH + 2 - 3) // 2 + 1, (IW + 2 - 3) // 2 + 1
    out = torch.empty(N, OH, OW, 64, device=y0.device, dtype=torch.bfloat16)
    am = torch.empty(N, OH, OW, 64, device=y0.device, dtype=torch.uint8)
    check(lib().svsr_stem_bn_gelu_pool_fwd(ptr(y0), ptr(coef), ptr(out), ptr(am), _i(N), _i(IH), _i(IW), stream_ptr()),
          "svsr_stem_bn_gelu_pool_fwd")
    return out, am


def stem_pool_gelu_bwd(dout, argmax, y0, coef):
    N, IH, IW, _ = y0.shape
    dz = torch.empty_like(y0)
    check(lib().svsr_stem_pool_gelu_bwd(ptr(dout), ptr(argmax), ptr(y0), ptr(coef), ptr(dz), _i(N), _i(IH), _i(IW),
                                        stream_ptr()), "svsr_stem_pool_gelu_bwd")
    return dz


def stem_bwd_fused(dout, argmax, y0, coef):
    """Pool scatter * GELU' -> BatchNorm backward in two passes. Returns (dc bf16, dgamma, dbeta)."""
    N, IH, IW, _ = y0.shape
    dout = dout.clone()  # the op overwrites its upstream gradient
    dc = torch.empty_like(y0)
    dgamma = torch.zeros(64, device=y0.device)
    dbeta = torch.zeros(64, device=y0.device)
    scratch = torch.zeros(128, device=y0.device, dtype=torch.float64)
    kcoef = torch.empty(128, device=y0.device)
    check(lib().svsr_stem_bwd_fused(ptr(dout), ptr(argmax), ptr(y0), ptr(coef), ptr(dgamma), ptr(dbeta), ptr(dc),
                                    ptr(scratch), ptr(kcoef), _i(N), _i(IH), _i(IW), stream_ptr()),
          "svsr_stem_bwd_fused")
    return dc, dgamma, dbeta


def meanpool_cls_fwd(a, cls, B, T):
    N, H, W, Cc = a.shape
    xs = torch.empty(B, T + 1, Cc, device=a.device, dtype=torch.float32)
    check(lib().svsr_meanpool_cls_fwd(ptr(a), ptr(cls), ptr(xs), _i(B), _i(T), _i(H * W), _i(Cc), stream_ptr()),
          "svsr_meanpool_cls_fwd")
    return xs


def meanpool_cls_bwd(dx, HW):
    B, T1, Cc = dx.shape
    T = T1 - 1
    dout = torch.empty(B * T, HW, Cc, device=dx.device, dtype=torch.bfloat16)
    dcls = torch.zeros(Cc, device=dx.device)
    check(lib().svsr_meanpool_cls_bwd(ptr(dx), ptr(dout), ptr(dcls), _i(B), _i(T), _i(HW), _i(Cc), stream_ptr()),
          "svsr_meanpool_cls_bwd")
    return dout, dcls


def rmsnorm_fwd(x, g, eps=1e-8):
    M, D = x.shape
    y = torch.empty(M, D, device=x.device, dtype=torch.bfloat16)
    inv = torch.empty(M, device=x.device)
    check(lib().svsr_rmsnorm_fwd(ptr(x), ptr(g), ptr(y), ptr(inv), _i(M), _i(D), C.c_float(eps), stream_ptr()),
          "svsr_rmsnorm_fwd")
    return y, inv


def rmsnorm_bwd(dy, x, g, inv, dx, eps=1e-8):
    """dx (fp32) is accumulated in place; returns (dx_bf16, dg)."""
    M, D = x.shape
    dxb = torch.empty(M, D, device=x.device, dtype=torch.bfloat16)
    dg = torch.zeros(D, device=x.device)
    check(lib().svsr_rmsnorm_bwd(ptr(dy), ptr(x), ptr(g), ptr(inv), ptr(dx), ptr(dxb), ptr(dg), _i(M), _i(D),
                                 C.c_float(eps), stream_ptr()), "svsr_rmsnorm_bwd")
    return dxb, dg


def rotary_table(n, device="cuda"):
    tab = torch.empty(n, 32, device=device)
    check(lib().svsr_rotary_table(ptr(tab), _i(n), stream_ptr()), "svsr_rotary_table")
    return tab


def attention_fwd(qkv, rot, B, n, heads, rotary_v=True):
    o = torch.empty(B * n, heads * 64, device=qkv.device, dtype=torch.bfloat16)
    check(lib().svsr_attention_fwd(ptr(qkv), ptr(rot), ptr(o), _i(B), _i(n), _i(heads), _i(rotary_v), stream_ptr()),
          "svsr_attention_fwd")
    return o


def attention_bwd(qkv, rot, d_o, B, n, heads, rotary_v=True):
    dqkv = torch.empty_like(qkv)
    check(lib().svsr_attention_bwd(ptr(qkv), ptr(rot), ptr(d_o), ptr(dqkv), _i(B), _i(n), _i(heads), _i(rotary_v),
                                   stream_ptr()), "svsr_attention_bwd")
    return dqkv


def geglu_fwd(h, p_drop=0.0, seed=0):
    M, F2 = h.shape
    u = torch.empty(M, F2 // 2, device=h.device, dtype=torch.bfloat16)
    check(lib().svsr_geglu_fwd(ptr(h), ptr(u), _i(M), _i(F2 // 2), C.c_float(p_drop), C.c_uint64(seed), stream_ptr()),
          "svsr_geglu_fwd")
    return u


def geglu_bwd(h, du, p_drop=0.0, seed=0):
    dh = torch.empty_like(h)
    check(lib().svsr_geglu_bwd(ptr(h), ptr(du), ptr(dh), _i(h.shape[0]), _i(h.shape[1] // 2), C.c_float(p_drop),
                               C.c_uint64(seed), stream_ptr()), "svsr_geglu_bwd")
    return dh


def audio_ce(logits, tokens, T, A, G, V, dscale=1.0, want_grad=True):
    """logits fp32 [B*T, A*G*V]; tokens int64 [B, >=T*A, G]. Returns (loss_mean, dlogits bf16|None, bad_token_flag)."""
    _req(logits, torch.float32, "logits"), _req(tokens, torch.int64, "tokens")
    B = tokens.shape[0]
    dl = torch.empty(logits.shape, device=logits.device, dtype=torch.bfloat16) if want_grad else None
    acc = torch.zeros(8, device=logits.device, dtype=torch.float64)
    bad = torch.zeros(1, device=logits.device, dtype=torch.int32)
    check(lib().svsr_audio_ce(ptr(logits), _i(logits.shape[1]), ptr(tokens), C.c_int64(tokens.stride(0)), _i(B), _i(T),
                              _i(A), _i(G), _i(V), ptr(dl), ptr(acc), ptr(bad), C.c_float(dscale), stream_ptr()),
          "svsr_audio_ce")
    return acc[0] / (B * T * A * G), dl, bad


def category_ce(logits, labels, num_classes, label_smoothing=0.0, dscale=1.0):
    """logits fp32 [B, ld>=C]; labels int64 [B] or fp32 [B,C]. Returns (loss_mean, top1, top5, dlogits bf16 [B, ld])."""
    B, ld = logits.shape
    hard = labels if labels.dtype == torch.int64 else None
    soft = labels if labels.dtype == torch.float32 else None
    dl = torch.empty(B, ld, device=logits.device, dtype=torch.bfloat16)
    acc = torch.zeros(8, device=logits.device, dtype=torch.float64)
    check(lib().svsr_category_ce(ptr(logits), _i(ld), ptr(hard), ptr(soft), _i(B), _i(num_classes),
                                 C.c_float(label_smoothing), ptr(dl), _i(ld), ptr(acc), C.c_float(dscale),
                                 stream_ptr()), "svsr_category_ce")
    return acc[1] / B, acc[2] / B, acc[3] / B, dl


def conv2d_fprop_generic(x: torch.Tensor, w_packed: torch.Tensor, taps, out_dtype=torch.bfloat16) -> torch.Tensor:
    """x [N,H,W,Cin] bf16, w_packed [Cout, len(taps)*Cin]; taps = [(dh, dw), ...]; same-size output."""
    _req(x, torch.bfloat16, "x"), _req(w_packed, torch.bfloat16, "w_packed")
    N, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    n = len(taps)
    dh = (C.c_int * n)(*[int(t[0]) for t in taps])
    dw = (C.c_int * n)(*[int(t[1]) for t in taps])
    y = torch.empty(N, H, W, Cout, device=x.device, dtype=out_dtype)
    check(lib().svsr_conv_taps_fprop(ptr(x), ptr(w_packed), ptr(y), _i(N), _i(H), _i(W), _i(Cin), _i(Cout), _i(n), dh,
                                     dw, _i(out_dtype == torch.float32), stream_ptr()), "svsr_conv_taps_fprop")
    return y


def conv2d_fprop_bnstats(x: torch.Tensor, w_packed: torch.Tensor, R: int, S: int, stride: int, pad: int):
    """conv fprop + fused per-channel (sum, sum of squares) of the output; returns (y bf16, stats fp64 [2, Cout])."""
    _req(x, torch.bfloat16, "x"), _req(w_packed, torch.bfloat16, "w_packed")
    N, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    OH = (H + 2 * pad - R) // stride + 1
    OW = (W + 2 * pad - S) // stride + 1
    y = torch.empty(N, OH, OW, Cout, device=x.device, dtype=torch.bfloat16)
    stats = torch.zeros(2, Cout, device=x.device, dtype=torch.float64)
    check(lib().svsr_conv2d_fprop_bnstats(ptr(x), ptr(w_packed), ptr(y), ptr(stats), _i(N), _i(H), _i(W), _i(Cin),
                                          _i(Cout), _i(R), _i(S), _i(stride), _i(pad), stream_ptr()),
          "svsr_conv2d_fprop_bnstats")
    return y, stats
