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


def conv2d_dgrad_bnbwd(dy: torch.Tensor, wd_packed: torch.Tensor, H: int, W: int, R: int, S: int, stride: int, pad: int,
                       c0: torch.Tensor, coef0: torch.Tensor, resid: torch.Tensor | None = None,
                       relu_mask: torch.Tensor | None = None, self_mask: bool = False, c1: torch.Tensor | None = None,
                       coef1: torch.Tensor | None = None, dx: torch.Tensor | None = None):
    """Input gradient with the consumer BatchNorm's backward reduction fused into the epilogue (svsr_conv2d_dgrad_bnbwd).
    Returns (dx bf16 [N,H,W,Cin] already masked, stats0 fp64 [2,Cin], stats1 | None)."""
    _req(dy, torch.bfloat16, "dy"), _req(wd_packed, torch.bfloat16, "wd_packed"), _req(c0, torch.bfloat16, "c0")
    N, OH, OW, Cout = dy.shape
    Cin = wd_packed.shape[0]
    if dx is None:
        dx = torch.zeros(N, H, W, Cin, device=dy.device, dtype=torch.bfloat16)
    st0 = torch.zeros(2, Cin, device=dy.device, dtype=torch.float64)
    st1 = torch.zeros(2, Cin, device=dy.device, dtype=torch.float64) if c1 is not None else None
    rc = lib().svsr_conv2d_dgrad_bnbwd(ptr(dy), ptr(wd_packed), ptr(dx), ptr(resid), C.c_int(N), C.c_int(H), C.c_int(W),
                                       C.c_int(Cin), C.c_int(Cout), C.c_int(R), C.c_int(S), C.c_int(stride), C.c_int(pad),
                                       ptr(relu_mask), C.c_int(int(self_mask)), ptr(c0), ptr(coef0), ptr(st0), ptr(c1),
                                       ptr(coef1), ptr(st1), stream_ptr())
    check(rc, "svsr_conv2d_dgrad_bnbwd")
    return dx, st0, st1


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


def stem_conv_direct(videos: torch.Tensor, w_packed: torch.Tensor):
    """Conv3d(1,64,(5,7,7),(1,2,2),(2,3,3)) without the patch tensor: returns (y0 bf16 [B,T,OH*OW,64], stats fp64 [2,64])."""
    _req(videos, torch.float32, "videos"), _req(w_packed, torch.bfloat16, "w_packed")
    B, _, T, H, W = videos.shape
    vb = videos.to(torch.bfloat16).contiguous()
    y = torch.empty(B, T, (H // 2) * (W // 2), 64, device=videos.device, dtype=torch.bfloat16)
    stats = torch.zeros(2, 64, device=videos.device, dtype=torch.float64)
    check(lib().svsr_stem_conv_direct(ptr(vb), ptr(w_packed), ptr(y), ptr(stats), _i(B), _i(T), _i(H), _i(W),
                                      stream_ptr()), "svsr_stem_conv_direct")
    return y, stats


def stem_wgrad_direct(videos: torch.Tensor, dz: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Weight gradient of stem_conv_direct: fp32 [320, 64] (row kt*64 + kh*8 + kw); accumulates into `out`."""
    _req(videos, torch.float32, "videos"), _req(dz, torch.bfloat16, "dz")
    B, _, T, H, W = videos.shape
    vb = videos.to(torch.bfloat16).contiguous()
    if out is None:
        out = torch.zeros(320, 64, device=videos.device, dtype=torch.float32)
    check(lib().svsr_stem_wgrad_direct(ptr(vb), ptr(dz), ptr(out), _i(out.stride(0)), _i(B), _i(T), _i(H), _i(W),
                                       stream_ptr()), "svsr_stem_wgrad_direct")
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


def attention_qkv_fwd(xn, w, rot, B, n, heads, rotary_v=True):
    """Fused to_q|to_k|to_v projection + rotary + softmax + PV. xn bf16 [B*n, K], w bf16 [3*heads*64, K]; returns (qkv, o)."""
    qkv = torch.empty(B * n, 3 * heads * 64, device=xn.device, dtype=torch.bfloat16)
    o = torch.empty(B * n, heads * 64, device=xn.device, dtype=torch.bfloat16)
    check(lib().svsr_attention_qkv_fwd(ptr(xn), _i(xn.stride(0)), ptr(w), _i(w.shape[1]), ptr(rot), ptr(qkv), ptr(o), _i(B),
                                       _i(n), _i(heads), _i(rotary_v), stream_ptr()), "svsr_attention_qkv_fwd")
    return qkv, o


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


class AudioHead:
    """The fused audio head as an operator: audio_projection + reshape + log-softmax + NLL (lightning.py:82,168-171)
    without fp32 logits in HBM. forward() -> mean loss; backward() -> (dx bf16 [B*T, K], dw fp32 [A*G*V, K], db fp32)."""

    def __init__(self, B: int, T: int, A: int, G: int, V: int, K: int, device="cuda"):
        self.B, self.T, self.A, self.G, self.V, self.K = B, T, A, G, V, K
        rows, N = B * T, A * G * V
        self.part = torch.empty(rows * (N // 64), 2, device=device, dtype=torch.float32)
        self.xt = torch.zeros(rows * A * G, device=device, dtype=torch.float32)
        self.lse = torch.empty(rows * A * G, device=device, dtype=torch.float32)
        self.acc = torch.zeros(1, device=device, dtype=torch.float64)
        self.bad = torch.zeros(1, device=device, dtype=torch.int32)
        self.dlogits = None

    def _args(self, x, w, bias, tokens):
        _req(x, torch.bfloat16, "x"), _req(w, torch.bfloat16, "w"), _req(tokens, torch.int64, "tokens")
        assert x.shape == (self.B * self.T, self.K) and w.shape == (self.A * self.G * self.V, self.K)
        return (ptr(x), _i(self.K), ptr(w), _i(self.K), _i(self.K), ptr(bias), ptr(tokens), C.c_int64(tokens.stride(0)),
                _i(self.B), _i(self.T), _i(self.A), _i(self.G), _i(self.V))

    def forward(self, x, w, bias, tokens):
        self.acc.zero_(), self.bad.zero_()
        check(lib().svsr_audio_head_fwd(*self._args(x, w, bias, tokens), ptr(self.part), ptr(self.xt), ptr(self.lse),
                                        ptr(self.acc), ptr(self.bad), stream_ptr()), "svsr_audio_head_fwd")
        return self.acc[0] / (self.B * self.T * self.A * self.G)

    def backward(self, x, w, wt, bias, tokens, dscale=None, dx=None, dw=None, want_db=True):
        """wt = w.t().contiguous() (the input-gradient GEMM's operand). dscale defaults to 1/rows (mean reduction)."""
        rows = self.B * self.T * self.A * self.G
        N = self.A * self.G * self.V
        if self.dlogits is None:
            self.dlogits = torch.empty(self.B * self.T, N, device=x.device, dtype=torch.bfloat16)
        check(lib().svsr_audio_head_bwd(*self._args(x, w, bias, tokens), ptr(self.lse),
                                        C.c_float(1.0 / rows if dscale is None else dscale), ptr(None), ptr(self.dlogits),
                                        ptr(self.bad), stream_ptr()), "svsr_audio_head_bwd")
        dx = gemm(self.dlogits, wt, out=dx)
        dw = gemm_wgrad(self.dlogits, x, out=dw)
        db = self.dlogits.float().sum(0) if want_db else None
        return dx, dw, db


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


def conv_taps_fprop_bnstats(x: torch.Tensor, w_packed: torch.Tensor, taps):
    """conv2d_fprop_generic with bf16 output + fused per-channel (sum, sum of squares): returns (y, stats fp64 [2, Cout])."""
    _req(x, torch.bfloat16, "x"), _req(w_packed, torch.bfloat16, "w_packed")
    N, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    n = len(taps)
    dh = (C.c_int * n)(*[int(t[0]) for t in taps])
    dw = (C.c_int * n)(*[int(t[1]) for t in taps])
    y = torch.empty(N, H, W, Cout, device=x.device, dtype=torch.bfloat16)
    stats = torch.zeros(2, Cout, device=x.device, dtype=torch.float64)
    check(lib().svsr_conv_taps_fprop_bnstats(ptr(x), ptr(w_packed), ptr(y), ptr(stats), _i(N), _i(H), _i(W), _i(Cin),
                                             _i(Cout), _i(n), dh, dw, stream_ptr()), "svsr_conv_taps_fprop_bnstats")
    return y, stats


def conv_taps_wgrad(x: torch.Tensor, dy: torch.Tensor, taps, out: torch.Tensor | None = None) -> torch.Tensor:
    """Weight gradient of conv2d_fprop_generic: fp32 [len(taps)*Cin, Cout], row t*Cin+ci; accumulates into `out`."""
    _req(x, torch.bfloat16, "x"), _req(dy, torch.bfloat16, "dy")
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    n = len(taps)
    dh = (C.c_int * n)(*[int(t[0]) for t in taps])
    dw = (C.c_int * n)(*[int(t[1]) for t in taps])
    if out is None:
        out = torch.zeros(n * Cin, Cout, device=x.device, dtype=torch.float32)
    check(lib().svsr_conv_taps_wgrad(ptr(x), ptr(dy), ptr(out), _i(N), _i(H), _i(W), _i(Cin), _i(Cout), _i(n), dh, dw,
                                     stream_ptr()), "svsr_conv_taps_wgrad")
    return out


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


# ----------------------------------------------------------------------------------------------------------------
# LRS sentence-level operators (csrc/conformer.cu)
# ----------------------------------------------------------------------------------------------------------------
def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-12):
    """x fp32 [M,D] -> (y bf16, y fp32, stats fp32 [M,2])."""
    _req(x, torch.float32, "x")
    M, D = x.shape
    yb = torch.empty(M, D, device=x.device, dtype=torch.bfloat16)
    yf = torch.empty(M, D, device=x.device, dtype=torch.float32)
    stats = torch.empty(M, 2, device=x.device, dtype=torch.float32)
    check(lib().svsr_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(yb), ptr(yf), ptr(stats), C.c_int(M), C.c_int(D),
                                   C.c_float(eps), stream_ptr()), "svsr_layernorm_fwd")
    return yb, yf, stats


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, stats: torch.Tensor,
                  dx: torch.Tensor | None = None):
    """dy bf16 or fp32 [M,D]; returns (dx fp32 [= or += if given], dgamma, dbeta)."""
    M, D = x.shape
    acc = dx is not None
    if dx is None:
        dx = torch.empty(M, D, device=x.device, dtype=torch.float32)
    dg = torch.zeros(D, device=x.device)
    db = torch.zeros(D, device=x.device)
    is_f32 = dy.dtype == torch.float32
    check(lib().svsr_layernorm_bwd(ptr(None if is_f32 else dy), ptr(dy if is_f32 else None), ptr(x), ptr(gamma),
                                   ptr(stats), ptr(dx), C.c_int(int(acc)), ptr(dg), ptr(db), C.c_int(M), C.c_int(D),
                                   stream_ptr()), "svsr_layernorm_bwd")
    return dx, dg, db


def glu_fwd(h: torch.Tensor) -> torch.Tensor:
    _req(h, torch.bfloat16, "h")
    M, C2 = h.shape
    u = torch.empty(M, C2 // 2, device=h.device, dtype=torch.bfloat16)
    check(lib().svsr_glu_fwd(ptr(h), ptr(u), C.c_int64(M), C.c_int(C2 // 2), stream_ptr()), "svsr_glu_fwd")
    return u


def glu_bwd(h: torch.Tensor, du: torch.Tensor) -> torch.Tensor:
    _req(h, torch.bfloat16, "h"), _req(du, torch.bfloat16, "du")
    dh = torch.empty_like(h)
    check(lib().svsr_glu_bwd(ptr(h), ptr(du), ptr(dh), C.c_int64(h.shape[0]), C.c_int(h.shape[1] // 2), stream_ptr()),
          "svsr_glu_bwd")
    return dh


def dwconv1d_fwd(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, flip: bool = False) -> torch.Tensor:
    """x bf16 [B,T,C]; w fp32 [C,K]."""
    _req(x, torch.bfloat16, "x"), _req(w, torch.float32, "w")
    B, T, Cc = x.shape
    y = torch.empty_like(x)
    check(lib().svsr_dwconv1d_fwd(ptr(x), ptr(w), ptr(bias), ptr(y), C.c_int(B), C.c_int(T), C.c_int(Cc),
                                  C.c_int(w.shape[1]), C.c_int(int(flip)), stream_ptr()), "svsr_dwconv1d_fwd")
    return y


def dwconv1d_wgrad(x: torch.Tensor, dy: torch.Tensor, K: int):
    B, T, Cc = x.shape
    dw = torch.zeros(Cc, K, device=x.device)
    db = torch.zeros(Cc, device=x.device)
    check(lib().svsr_dwconv1d_wgrad(ptr(x), ptr(dy), ptr(dw), ptr(db), C.c_int(B), C.c_int(T), C.c_int(Cc), C.c_int(K),
                                    stream_ptr()), "svsr_dwconv1d_wgrad")
    return dw, db


def bn_col_reduce(x: torch.Tensor, dout: torch.Tensor | None = None, coef: torch.Tensor | None = None) -> torch.Tensor:
    """x bf16 [rows, C] -> fp64 [2, C] (mode 0: sum, sum of squares; mode 1 with dout/coef: sum g, sum g*xhat)."""
    rows, Cc = x.shape
    stats = torch.zeros(2, Cc, device=x.device, dtype=torch.float64)
    check(lib().svsr_bn_col_reduce(ptr(x), ptr(dout), ptr(coef), C.c_int64(rows), C.c_int(Cc), ptr(stats),
                                   C.c_int(0 if dout is None else 1), stream_ptr()), "svsr_bn_col_reduce")
    return stats


def _attn_args(q, k, v, p, bias_u, bias_v, klen, causal, B, H, Tq, Tk, scale):
    return (ptr(q), C.c_int(q.stride(0)), ptr(k), C.c_int(k.stride(0)), ptr(v), C.c_int(v.stride(0)), ptr(p),
            C.c_int(p.stride(0) if p is not None else 0), ptr(bias_u), ptr(bias_v), ptr(klen), C.c_int(int(causal)),
            C.c_int(B), C.c_int(H), C.c_int(Tq), C.c_int(Tk), C.c_float(scale))


def attention_core_fwd(q, k, v, B: int, H: int, Tq: int, Tk: int, p=None, bias_u=None, bias_v=None, klen=None,
                       causal: bool = False, scale: float = 0.125, drop_p: float = 0.0, drop_seed: int = 0):
    """q [B*Tq, >=H*64] / k, v [B*Tk, ...] bf16 row-major VIEWS (row pitch = stride(0)); returns (o bf16 [B*Tq,H*64], lse)."""
    o = torch.empty(B * Tq, H * 64, device=q.device, dtype=torch.bfloat16)
    lse = torch.empty(B, H, Tq, device=q.device, dtype=torch.float32)
    check(lib().svsr_attention_core_fwd(*_attn_args(q, k, v, p, bias_u, bias_v, klen, causal, B, H, Tq, Tk, scale),
                                        ptr(o), C.c_int(H * 64), ptr(lse), C.c_float(drop_p), C.c_uint64(drop_seed),
                                        stream_ptr()), "svsr_attention_core_fwd")
    return o, lse


def attention_core_bwd(q, k, v, o, lse, d_o, B: int, H: int, Tq: int, Tk: int, p=None, bias_u=None, bias_v=None,
                       klen=None, causal: bool = False, scale: float = 0.125, drop_p: float = 0.0, drop_seed: int = 0):
    """Returns dq, dk, dv (bf16, same pitches as q/k/v: pass contiguous-row views), dp fp32, dbias_u, dbias_v."""
    L = lib()
    L.svsr_attention_scratch_bytes.restype = C.c_int64
    dq = torch.zeros(q.shape[0], q.stride(0), device=q.device, dtype=torch.bfloat16)
    dk = torch.zeros(k.shape[0], k.stride(0), device=q.device, dtype=torch.bfloat16)
    dv = torch.zeros(v.shape[0], v.stride(0), device=q.device, dtype=torch.bfloat16)
    dp = torch.zeros(2 * Tk - 1, H * 64, device=q.device) if p is not None else None
    dbu = torch.zeros(H, 64, device=q.device) if bias_u is not None else None
    dbv = torch.zeros(H, 64, device=q.device) if bias_v is not None else None
    scratch = torch.empty(L.svsr_attention_scratch_bytes(B, H, Tq, Tk), device=q.device, dtype=torch.uint8)
    check(L.svsr_attention_core_bwd(*_attn_args(q, k, v, p, bias_u, bias_v, klen, causal, B, H, Tq, Tk, scale),
                                    ptr(o), C.c_int(H * 64), ptr(lse), ptr(d_o), ptr(dq), ptr(dk), ptr(dv), ptr(dp),
                                    ptr(dbu), ptr(dbv), ptr(scratch), C.c_float(drop_p), C.c_uint64(drop_seed),
                                    stream_ptr()), "svsr_attention_core_bwd")
    return dq[:, : H * 64], dk[:, : H * 64], dv[:, : H * 64], dp, dbu, dbv


def ctc_loss(logits: torch.Tensor, V: int, labels: torch.Tensor, in_len: torch.Tensor, B: int, T: int,
             dscale: float = 1.0):
    """logits fp32 [B*T, ld]; labels int64 [B,Lmax] (-1 padded); in_len int32 [B]. Returns (sum nll, dlogits bf16)."""
    L = lib()
    L.svsr_ctc_scratch_bytes.restype = C.c_int64
    _req(logits, torch.float32, "logits"), _req(labels, torch.int64, "labels"), _req(in_len, torch.int32, "in_len")
    ld, Lmax = logits.shape[1], labels.shape[1]
    dl = torch.empty(B * T, ld, device=logits.device, dtype=torch.bfloat16)
    acc = torch.zeros(4, device=logits.device, dtype=torch.float64)
    scratch = torch.empty(L.svsr_ctc_scratch_bytes(B, T, Lmax), device=logits.device, dtype=torch.uint8)
    check(L.svsr_ctc_loss(ptr(logits), C.c_int(ld), C.c_int(V), ptr(labels), C.c_int(Lmax), ptr(in_len), C.c_int(B),
                          C.c_int(T), ptr(dl), ptr(acc), C.c_int(0), C.c_float(dscale), ptr(scratch), stream_ptr()),
          "svsr_ctc_loss")
    return acc[0], dl


def label_smoothing_loss(logits: torch.Tensor, V: int, target: torch.Tensor, smoothing: float, dscale: float = 1.0):
    """logits fp32 [rows, ld]; target int64 [rows]. Returns (acc fp64 [KL sum, #correct, #scored], dlogits bf16)."""
    _req(logits, torch.float32, "logits"), _req(target, torch.int64, "target")
    rows, ld = logits.shape
    dl = torch.empty(rows, ld, device=logits.device, dtype=torch.bfloat16)
    acc = torch.zeros(4, device=logits.device, dtype=torch.float64)
    check(lib().svsr_label_smoothing_loss(ptr(logits), C.c_int(ld), C.c_int(V), ptr(target), C.c_int(rows),
                                          C.c_float(smoothing), ptr(dl), ptr(acc), C.c_int(0), C.c_float(dscale),
                                          stream_ptr()), "svsr_label_smoothing_loss")
    return acc, dl


def gemm_ex(a: torch.Tensor, b: torch.Tensor, bias=None, resid=None, out_dtype=torch.bfloat16, alpha: float = 1.0,
            bias_scale: float = 1.0, relu: bool = False, relu_mask: torch.Tensor | None = None, drop_p: float = 0.0,
            drop_seed: int = 0) -> torch.Tensor:
    """out = act(dropout(alpha * a @ b.T + bias_scale * bias) + resid) * [relu_mask > 0]."""
    _req(a, torch.bfloat16, "a"), _req(b, torch.bfloat16, "b")
    M, K = a.shape
    N = b.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    check(lib().svsr_gemm_bf16_ex(ptr(a), C.c_int(K), ptr(b), C.c_int(K), ptr(out), C.c_int(N), ptr(bias), ptr(resid),
                                  C.c_int(M), C.c_int(N), C.c_int(K), C.c_int(int(out_dtype == torch.float32)),
                                  C.c_int(int(resid is not None and resid.dtype == torch.float32)), C.c_float(alpha),
                                  C.c_float(bias_scale), C.c_int(int(relu)), ptr(relu_mask), C.c_float(drop_p),
                                  C.c_uint64(drop_seed), stream_ptr()),
          "svsr_gemm_bf16_ex")
    return out


def dropout_mask(n: int, p: float, seed: int, device="cuda") -> torch.Tensor:
    """uint8 keep-mask of the counter-based dropout (element i kept iff mask[i] == 1)."""
    out = torch.empty(n, device=device, dtype=torch.uint8)
    check(lib().svsr_dropout_mask(ptr(out), C.c_int64(n), C.c_float(p), C.c_uint64(seed), stream_ptr()), "svsr_dropout_mask")
    return out
