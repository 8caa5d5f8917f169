"""Kernel-level parity (through the C ABI) against plain fp32 PyTorch on IDENTICAL inputs.

Tolerances: the kernels compute in fp32 from bf16-stored operands, so outputs stored in bf16 carry one bf16 rounding
(2^-9 relative, rel-L2 ~2e-3); fp32 outputs (weight gradients, losses, statistics) must agree to ~1e-4.
Integer work (audio-token indexing) is checked bit-exactly."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_TOL = 4e-3
F32_TOL = 2e-4


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from syncvsr_b200 import ops as o

    return o


def randn(*shape, seed=0, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


# ---------------------------------------------------------------- GEMM / conv (tcgen05) -------------------------
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1920, 512, 512), (1920, 4096, 512), (100, 512, 2048), (64, 504, 512)])
def test_gemm(ops, M, N, K):
    a, b = randn(M, K, seed=1), randn(N, K, seed=2, scale=0.05)
    bias = randn(N, seed=3, dtype=torch.float32)
    out = ops.gemm(a, b, bias=bias, out_dtype=torch.float32)
    ref = a.float() @ b.float().T + bias
    assert rel(out, ref) < F32_TOL


def test_gemm_residual_bf16(ops):
    a, b, r = randn(1920, 2048, seed=1), randn(512, 2048, seed=2, scale=0.05), randn(1920, 512, seed=3)
    out = ops.gemm(a, b, resid=r)
    assert rel(out, a.float() @ b.float().T + r.float()) < BF16_TOL


CONVS = [(6, 22, 22, 64, 64, 3, 1, 1), (6, 11, 11, 128, 128, 3, 1, 1), (7, 6, 6, 256, 256, 3, 1, 1),
         (30, 3, 3, 512, 512, 3, 1, 1), (6, 22, 22, 64, 128, 3, 2, 1), (6, 11, 11, 128, 256, 3, 2, 1),
         (6, 6, 6, 256, 512, 3, 2, 1), (6, 22, 22, 64, 128, 1, 2, 0), (5, 11, 11, 128, 256, 1, 2, 0),
         (3, 24, 24, 64, 64, 3, 1, 1), (3, 12, 12, 128, 128, 3, 1, 1)]


@pytest.mark.parametrize("N,H,W,Cin,Cout,R,stride,pad", CONVS)
def test_conv_fprop_dgrad_wgrad(ops, N, H, W, Cin, Cout, R, stride, pad):
    x = randn(N, H, W, Cin, seed=4)
    w = randn(Cout, Cin, R, R, seed=5, scale=0.05)
    xt = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wt = w.float().requires_grad_(True)
    ref = F.conv2d(xt, wt, stride=stride, padding=pad)
    y = ops.conv2d_fprop(x, ops.pack_conv_weight(w), R, R, stride, pad)
    assert rel(y, ref.permute(0, 2, 3, 1)) < BF16_TOL
    dy = randn(*y.shape, seed=6, scale=0.1)
    gx, gw = torch.autograd.grad(ref, (xt, wt), dy.float().permute(0, 3, 1, 2))
    dx = ops.conv2d_dgrad(dy, ops.pack_conv_weight_dgrad(w), H, W, R, R, stride, pad)
    assert rel(dx, gx.permute(0, 2, 3, 1)) < BF16_TOL
    dw = ops.unpack_conv_wgrad(ops.conv2d_wgrad(x, dy, R, R, stride, pad), Cin, R, R)
    assert rel(dw, gw) < F32_TOL


@pytest.mark.parametrize("N,H,W,Cin,Cout,R,stride,pad", [(40, 22, 22, 64, 64, 3, 1, 1), (33, 11, 11, 128, 128, 3, 1, 1),
                                                          (35, 22, 22, 64, 128, 3, 2, 1), (50, 6, 6, 256, 512, 1, 2, 0),
                                                          (700, 3, 3, 512, 512, 3, 1, 1)])
def test_conv_fused_batchnorm_statistics(ops, N, H, W, Cin, Cout, R, stride, pad):
    x = randn(N, H, W, Cin, seed=14)
    w = randn(Cout, Cin, R, R, seed=15, scale=0.05)
    y, stats = ops.conv2d_fprop_bnstats(x, ops.pack_conv_weight(w), R, R, stride, pad)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), stride=stride, padding=pad)
    assert rel(y, ref.permute(0, 2, 3, 1)) < BF16_TOL
    # the statistics are those of the stored (bf16) conv output -- exactly the tensor BatchNorm then normalises
    yd = y.double()
    assert rel(stats[0], yd.sum((0, 1, 2))) < 1e-4
    assert rel(stats[1], (yd ** 2).sum((0, 1, 2))) < 1e-4
    assert rel(stats[1], (ref.double() ** 2).sum((0, 2, 3))) < 2e-3


BNB_CASES = [
    # N, H, W, Cin, Cout, stride, mask ("self" | "ref"), resid, two BatchNorms     (the launches of frontend_backward)
    (40, 22, 22, 64, 64, 1, "self", False, False),    # layer1 conv2 dgrad (halo kernel): bn1's own ReLU
    (40, 22, 22, 64, 64, 1, "ref", True, False),      # layer1.1 conv1 dgrad (halo): + identity shortcut, relu(out) of layer1.0
    (3, 24, 24, 64, 64, 1, "ref", True, False),       # 96x96 crops
    (35, 22, 22, 64, 128, 2, "ref", True, False),     # layer2.0 conv1 dgrad: stride 2 (4 pixel classes), in-place shortcut
    (33, 11, 11, 128, 128, 1, "self", False, False),  # layer2 conv2 dgrad (generic kernel)
    (33, 11, 11, 128, 128, 1, "ref", True, True),     # layer2.1 conv1 dgrad: feeds bn2 AND downsample.1 of layer2.0
    (41, 6, 6, 256, 256, 1, "ref", True, True),       # layer3.1 conv1 dgrad, two BatchNorms over 256 channels
    (300, 3, 3, 512, 512, 1, "ref", True, False),     # layer4.1 conv1 dgrad, 512 channels
    (300, 3, 3, 512, 512, 1, "self", False, False),
]


@pytest.mark.parametrize("N,H,W,Cin,Cout,stride,mask,with_resid,two", BNB_CASES)
def test_dgrad_with_fused_batchnorm_backward_statistics(ops, N, H, W, Cin, Cout, stride, mask, with_resid, two):
    """svsr_conv2d_dgrad_bnbwd: the masked input gradient equals (plain dgrad + shortcut) * mask bit for bit, and the sums it
    leaves are those of the standalone BatchNorm-backward reduce over the STORED gradient: sum g, sum g * xhat."""
    OH = (H + 2 - 3) // stride + 1
    w = randn(Cout, Cin, 3, 3, seed=41, scale=0.05)
    dy = randn(N, OH, OH, Cout, seed=42, scale=0.1)
    wd = ops.pack_conv_weight_dgrad(w)
    c0, c1 = randn(N, H, W, Cin, seed=43), randn(N, H, W, Cin, seed=44)
    resid = randn(N, H, W, Cin, seed=45, scale=0.3) if with_resid else None
    ref_mask = randn(N, H, W, Cin, seed=46)
    g = torch.Generator(device="cuda").manual_seed(47)

    def coef_of(c):
        mean, var = c.float().mean((0, 1, 2)), c.float().var((0, 1, 2), unbiased=False)
        gamma = 1 + 0.2 * torch.randn(Cin, device="cuda", generator=g)
        beta = 0.3 * torch.randn(Cin, device="cuda", generator=g)
        inv = torch.rsqrt(var + 1e-5)
        return torch.stack([mean, inv, gamma * inv, beta - mean * gamma * inv]).contiguous()

    cf0, cf1 = coef_of(c0), coef_of(c1)
    plain = ops.conv2d_dgrad(dy, wd, H, W, 3, 3, stride, 1, resid=resid)
    ambiguous = torch.zeros(N, H, W, Cin, dtype=torch.bool, device="cuda")
    if mask == "self":
        z = c0.double() * cf0[2].double() + cf0[3].double()
        keep = z > 0
        # the kernel evaluates fma(c, scale, shift) in fp32: a pre-activation within rounding of zero may fall either way
        ambiguous = z.abs() < 1e-6 * ((c0.double() * cf0[2].double()).abs() + cf0[3].double().abs())
    else:
        keep = ref_mask.float() > 0
    want = torch.where(keep, plain.float(), torch.zeros((), device="cuda")).bfloat16()
    dx_buf = resid.clone() if (with_resid and stride == 2) else None  # the engine's strided blocks add in place
    dx, st0, st1 = ops.conv2d_dgrad_bnbwd(dy, wd, H, W, 3, 3, stride, 1, c0, cf0, resid=dx_buf if dx_buf is not None else resid,
                                          relu_mask=None if mask == "self" else ref_mask, self_mask=mask == "self",
                                          c1=c1 if two else None, coef1=cf1 if two else None, dx=dx_buf)
    assert torch.equal(dx[~ambiguous], want[~ambiguous]) and int(ambiguous.sum()) < 50
    gd = dx.double()
    for st, c, cf in ((st0, c0, cf0), (st1, c1, cf1)):
        if st is None:
            continue
        xhat = (c.double() - cf[0].double()) * cf[1].double()
        assert rel(st[0], gd.sum((0, 1, 2))) < 1e-5
        assert rel(st[1], (gd * xhat).sum((0, 1, 2))) < 1e-4


@pytest.mark.parametrize("N,H,W", [(40, 22, 22), (3, 24, 24), (5, 7, 25), (2, 1, 2), (9, 13, 5), (1, 30, 11)])
def test_halo_conv_matches_generic_igemm(ops, monkeypatch, N, H, W):
    """igemm_halo.cu (one activation load per tile, resident weights) against the tap-by-tap kernel it replaces for the
    64 -> 64 channel 3x3 convs, and against fp32 PyTorch: forward (+ fused BN statistics) and input gradient (+ residual)."""
    x, w = randn(N, H, W, 64, seed=31), randn(64, 64, 3, 3, seed=32, scale=0.05)
    dy, base = randn(N, H, W, 64, seed=33, scale=0.1), randn(N, H, W, 64, seed=34)
    wp, wd = ops.pack_conv_weight(w), ops.pack_conv_weight_dgrad(w)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SVSR_HALO_CONV", mode)
        y, stats = ops.conv2d_fprop_bnstats(x, wp, 3, 3, 1, 1)
        dx = ops.conv2d_dgrad(dy, wd, H, W, 3, 3, 1, 1, resid=base)
        res[mode] = (y.clone(), stats.clone(), dx.clone(), ops.conv2d_fprop(x, wp, 3, 3, 1, 1, resid=base).clone())
    # same MMA sequence per output pixel -> identical bf16 results
    assert torch.equal(res["1"][0], res["0"][0])
    assert torch.equal(res["1"][2], res["0"][2])
    assert torch.equal(res["1"][3], res["0"][3])
    assert rel(res["1"][1], res["0"][1]) < 1e-6
    xt = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.conv2d(xt, w.float(), padding=1)
    assert rel(res["1"][0], ref.permute(0, 2, 3, 1)) < BF16_TOL
    gx, = torch.autograd.grad(ref, xt, dy.float().permute(0, 3, 1, 2))
    assert rel(res["1"][2], gx.permute(0, 2, 3, 1) + base.float()) < BF16_TOL
    yd = res["1"][0].double()
    assert rel(res["1"][1][0], yd.sum((0, 1, 2))) < 1e-4 and rel(res["1"][1][1], (yd ** 2).sum((0, 1, 2))) < 1e-4


@pytest.mark.parametrize("N,T,W", [(2, 29, 1936), (1, 8, 16), (3, 5, 48), (1, 1, 32), (2, 150, 64), (5, 21, 2304)])
def test_stem_temporal_halo_conv_matches_generic_igemm(ops, monkeypatch, N, T, W):
    """igemm_stem.cu (one activation load per 8-frame x 16-pixel tile, resident weights) against the tap-by-tap kernel
    it replaces for the stem's temporal 5-tap contraction (lightning.py:50), and against fp32 PyTorch: output and fused
    BatchNorm statistics; clip lengths that are not multiples of the 8-frame tile, a single frame, LRS length."""
    x, w = randn(N, T, W, 64, seed=41), randn(64, 64, 5, 1, seed=42, scale=0.05)
    wp = ops.pack_conv_weight(w)  # [64, 5*64], K = kt*64 + c
    taps = [(kt - 2, 0) for kt in range(5)]
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SVSR_STEM_HALO", mode)
        y, stats = ops.conv_taps_fprop_bnstats(x, wp, taps)
        torch.cuda.synchronize()
        res[mode] = (y.clone(), stats.clone())
    assert torch.equal(res["1"][0], res["0"][0])  # same MMA sequence per output -> identical bf16 results
    assert rel(res["1"][1], res["0"][1]) < 1e-6
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=(2, 0)).permute(0, 2, 3, 1)
    assert rel(res["1"][0], ref) < BF16_TOL
    yd = res["1"][0].double()
    assert rel(res["1"][1][0], yd.sum((0, 1, 2))) < 1e-4 and rel(res["1"][1][1], (yd ** 2).sum((0, 1, 2))) < 1e-4
    # a width that is not a multiple of the 16-pixel tile stays on the generic kernel
    x2 = randn(1, 6, 40, 64, seed=43)
    monkeypatch.setenv("SVSR_STEM_HALO", "1")
    y2, _ = ops.conv_taps_fprop_bnstats(x2, wp, taps)
    ref2 = F.conv2d(x2.float().permute(0, 3, 1, 2), w.float(), padding=(2, 0)).permute(0, 2, 3, 1)
    assert rel(y2, ref2) < BF16_TOL


@pytest.mark.parametrize("N,H,W", [(40, 22, 22), (3, 24, 24), (5, 7, 25), (2, 1, 2), (9, 13, 5), (300, 22, 22)])
def test_halo_wgrad_matches_generic_and_autograd(ops, monkeypatch, N, H, W):
    """wgrad_halo.cu (one activation + one gradient load per pixel tile, tap pairs as descriptor offsets) against the
    generic split-K kernel and fp32 autograd; accumulation into a non-zero gradient buffer."""
    x, dy = randn(N, H, W, 64, seed=41), randn(N, H, W, 64, seed=42, scale=0.1)
    base = torch.randn(9 * 64, 64, device="cuda")
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SVSR_HALO_CONV", mode)
        res[mode] = ops.conv2d_wgrad(x, dy, 3, 3, 1, 1, out=base.clone())
    assert rel(res["1"] - base, res["0"] - base) < 2e-5  # same products, different fp32 summation order
    xt = x.float().permute(0, 3, 1, 2)
    wt = torch.zeros(64, 64, 3, 3, device="cuda", requires_grad=True)
    F.conv2d(xt, wt, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    assert rel(ops.unpack_conv_wgrad(res["1"] - base, 64, 3, 3), wt.grad) < F32_TOL


@pytest.mark.parametrize("N,T,W", [(2, 29, 1936), (1, 8, 16), (3, 5, 48), (1, 1, 32), (2, 150, 64), (40, 29, 1936)])
def test_stem_temporal_halo_wgrad_matches_generic_and_autograd(ops, monkeypatch, N, T, W):
    """wgrad_stem.cu (one patch tile + one gradient tile per 8-frame x 16-pixel tile, tap pairs as descriptor offsets)
    against the generic split-K kernel and fp32 autograd; accumulation into a non-zero gradient buffer."""
    x, dy = randn(N, T, W, 64, seed=51), randn(N, T, W, 64, seed=52, scale=0.1)
    taps = [(kt - 2, 0) for kt in range(5)]
    base = torch.randn(5 * 64, 64, device="cuda")
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SVSR_STEM_HALO", mode)
        res[mode] = ops.conv_taps_wgrad(x, dy, taps, out=base.clone())
        torch.cuda.synchronize()
    assert rel(res["1"] - base, res["0"] - base) < 2e-5  # same products, different fp32 summation order
    # fp64 reference as five shifted [64 x pixels] . [pixels x 64] products (2.2 M pixels per sum at the bench geometry:
    # cuDNN's fp32/TF32 backward-weight is itself 5e-3 off there)
    xd, dyd = x.double(), dy.double().reshape(-1, 64)
    ref = torch.zeros(5, 64, 64, device="cuda", dtype=torch.float64)  # [tap, ci, co]
    for kt in range(5):
        xs = torch.zeros_like(xd)
        lo, hi = max(0, 2 - kt), min(T, T + 2 - kt)  # output frames t with 0 <= t + kt - 2 < T
        xs[:, lo:hi] = xd[:, lo + kt - 2:hi + kt - 2]
        ref[kt] = xs.reshape(-1, 64).T @ dyd
    assert rel(res["1"] - base, ref.reshape(320, 64)) < F32_TOL


def test_conv_dgrad_accumulates_residual_in_place(ops):
    dy, w = randn(4, 11, 11, 128, seed=7), randn(128, 64, 3, 3, seed=8, scale=0.05)
    base = randn(4, 22, 22, 64, seed=9)
    ref = ops.conv2d_dgrad(dy, ops.pack_conv_weight_dgrad(w), 22, 22, 3, 3, 2, 1).float() + base.float()
    out = ops.conv2d_dgrad(dy, ops.pack_conv_weight_dgrad(w), 22, 22, 3, 3, 2, 1, resid=base)
    assert rel(out, ref) < BF16_TOL


def test_gemm_wgrad(ops):
    dy, x = randn(1920, 512, seed=10, scale=0.1), randn(1920, 2048, seed=11)
    assert rel(ops.gemm_wgrad(dy, x), dy.float().T @ x.float()) < F32_TOL


def test_stem_conv_matches_conv3d(ops):
    """patch gather + 5-tap temporal implicit GEMM == Conv3d(1,64,(5,7,7),(1,2,2),(2,3,3)) (lightning.py:50)."""
    import ctypes as C
    from syncvsr_b200._lib import check, lib, ptr, stream_ptr

    B, T, S = 2, 29, 88
    v = randn(B, 1, T, S, S, seed=12, dtype=torch.float32)
    w = randn(64, 1, 5, 7, 7, seed=13, scale=0.05, dtype=torch.float32)
    ref = F.conv3d(v.bfloat16().float(), w.bfloat16().float(), None, (1, 2, 2), (2, 3, 3))
    P = ops.stem_patch(v)
    wp = torch.zeros(64, 5, 8, 8, device="cuda")
    wp[:, :, :7, :7] = w[:, 0]
    wp = wp.reshape(64, 320).bfloat16().contiguous()
    # temporal 5-tap conv == conv2d with a 5x1 filter over an [T, OH*OW] "image" of 64 patch channels
    y = ops.conv2d_fprop_generic(P.view(B, T, 44 * 44, 64), wp, taps=[(kt - 2, 0) for kt in range(5)])
    assert rel(y.view(B, T, 44, 44, 64), ref.permute(0, 2, 3, 4, 1)) < BF16_TOL


@pytest.mark.parametrize("B,T,S", [(2, 29, 88), (1, 8, 96), (3, 5, 88), (1, 1, 88), (1, 150, 88), (5, 21, 96)])
def test_stem_without_patch_tensor_matches_patch_path(ops, B, T, S):
    """stem_direct.cu (7x7/s2 window rows built in shared memory from the bf16 video) against the patch tensor + 5-tap
    temporal implicit GEMM it replaces: forward bit-identical (same operand bytes, same MMA sequence), fused BatchNorm
    statistics, weight gradient up to the fp32 order of the cross-CTA reduction; and against Conv3d itself."""
    v = randn(B, 1, T, S, S, seed=61, dtype=torch.float32)
    w = randn(64, 1, 5, 7, 7, seed=62, scale=0.05, dtype=torch.float32)
    wp = torch.zeros(64, 5, 8, 8, device="cuda")
    wp[:, :, :7, :7] = w[:, 0]
    wp = wp.reshape(64, 320).bfloat16().contiguous()
    taps = [(kt - 2, 0) for kt in range(5)]
    P = ops.stem_patch(v)
    y_ref, st_ref = ops.conv_taps_fprop_bnstats(P, wp, taps)
    y, st = ops.stem_conv_direct(v, wp)
    torch.cuda.synchronize()
    assert torch.equal(y, y_ref)
    assert rel(st, st_ref) < 1e-6
    ref = F.conv3d(v.bfloat16().float(), w.bfloat16().float(), None, (1, 2, 2), (2, 3, 3))
    assert rel(y.view(B, T, S // 2, S // 2, 64), ref.permute(0, 2, 3, 4, 1)) < BF16_TOL
    dz = randn(B, T, (S // 2) ** 2, 64, seed=63, scale=0.1)
    base = torch.randn(320, 64, device="cuda")
    g_ref = ops.conv_taps_wgrad(P, dz, taps, out=base.clone())
    g = ops.stem_wgrad_direct(v, dz, out=base.clone())
    torch.cuda.synchronize()
    assert rel(g - base, g_ref - base) < 2e-5
    wt = torch.zeros(64, 1, 5, 7, 7, device="cuda", dtype=torch.float64, requires_grad=True)  # fp64: no TF32 in the reference
    F.conv3d(v.bfloat16().double(), wt, None, (1, 2, 2), (2, 3, 3)).backward(
        dz.double().view(B, T, S // 2, S // 2, 64).permute(0, 4, 1, 2, 3))
    gw = (g - base).view(5, 8, 8, 64)[:, :7, :7].permute(3, 0, 1, 2)  # [co, kt, kh, kw]
    assert rel(gw, wt.grad[:, 0]) < F32_TOL
    pad = (g - base).view(5, 8, 8, 64)  # the padding slots (kh = 7, kw = 7) see zero operands
    assert float(pad[:, 7].abs().max()) == 0.0 and float(pad[:, :, 7].abs().max()) == 0.0


def test_stem_without_patch_tensor_refuses_uncovered_frames(ops):
    from syncvsr_b200._lib import SvsrError

    v = randn(1, 1, 4, 90, 90, seed=64, dtype=torch.float32)  # 45 x 45 output pixels: not a multiple of the 16-pixel tile
    with pytest.raises(SvsrError):
        ops.stem_conv_direct(v, torch.zeros(64, 320, device="cuda", dtype=torch.bfloat16))


# ---------------------------------------------------------------- BatchNorm --------------------------------------
@pytest.mark.parametrize("C,rows", [(64, 5000), (128, 1111), (256, 700), (512, 90)])
def test_batchnorm_fwd_bwd(ops, C, rows):
    x = randn(rows, C, seed=20) * 2 + 0.5
    res = randn(rows, C, seed=21)
    gamma = randn(C, seed=22, dtype=torch.float32).abs() + 0.5
    beta = randn(C, seed=23, dtype=torch.float32)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    out, coef = ops.batchnorm_fwd(x, gamma, beta, rm, rv, train=True, res=res, relu=True)
    xf = x.float().requires_grad_(True)
    gf, bf = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    pre = F.batch_norm(xf, rm2, rv2, gf, bf, True, 0.1, 1e-5)
    ref = F.relu(pre + res.float())
    assert rel(out, ref) < BF16_TOL
    assert rel(rm, rm2) < F32_TOL and rel(rv, rv2) < F32_TOL
    # backward with the kernel's own output as relu reference (identical mask on both sides)
    dout = randn(rows, C, seed=24, scale=0.1)
    mask = (out.float() > 0).float()
    gx, gg, gb = torch.autograd.grad(pre, (xf, gf, bf), dout.float() * mask)
    dc, dgamma, dbeta, gm = ops.batchnorm_bwd(dout, out, x, coef, want_gmask=True)
    assert rel(dc, gx) < BF16_TOL
    assert rel(dgamma, gg) < 1e-3 and rel(dbeta, gb) < 1e-3
    assert torch.equal(gm.float(), dout.float() * mask)


def test_batchnorm_eval_uses_running_stats(ops):
    C = 64
    x = randn(300, C, seed=25)
    gamma, beta = torch.ones(C, device="cuda") * 1.5, torch.ones(C, device="cuda") * 0.1
    rm, rv = randn(C, seed=26, dtype=torch.float32) * 0.1, torch.rand(C, device="cuda") + 0.5
    out, _ = ops.batchnorm_fwd(x, gamma, beta, rm.clone(), rv.clone(), train=False)
    ref = F.batch_norm(x.float(), rm, rv, gamma, beta, False, 0.1, 1e-5)
    assert rel(out, ref) < BF16_TOL


# ---------------------------------------------------------------- stem BN+GELU+pool ------------------------------
def test_stem_bn_gelu_pool_fwd_bwd(ops):
    N, IH = 5, 44
    y0 = randn(N, IH, IH, 64, seed=30)
    coef = torch.zeros(4, 64, device="cuda")
    coef[1] = 1.0
    coef[2] = randn(64, seed=31, dtype=torch.float32).abs() + 0.5  # scale
    coef[3] = randn(64, seed=32, dtype=torch.float32) * 0.3  # shift
    out, am = ops.stem_bn_gelu_pool_fwd(y0, coef)
    z = (y0.float() * coef[2] + coef[3]).permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.max_pool2d(F.gelu(z), 3, 2, 1)
    assert rel(out, ref.permute(0, 2, 3, 1)) < BF16_TOL
    dout = randn(*out.shape, seed=33)
    (gz,) = torch.autograd.grad(ref, z, dout.float().permute(0, 3, 1, 2))
    dz = ops.stem_pool_gelu_bwd(dout, am, y0, coef)
    assert rel(dz, gz.permute(0, 2, 3, 1)) < BF16_TOL


def test_stem_bwd_fused_matches_three_pass_and_autograd(ops):
    """svsr_stem_bwd_fused == stem_pool_gelu_bwd -> batchnorm_bwd (the path it replaces) and == autograd through
    BatchNorm3d(train) -> GELU -> MaxPool (lightning.py:51-53); odd sizes exercise clipped windows."""
    for N, IH, seed in ((5, 44, 60), (3, 7, 70)):
        y0 = randn(N, IH, IH, 64, seed=seed)
        gamma = randn(64, seed=seed + 1, dtype=torch.float32).abs() + 0.5
        beta = randn(64, seed=seed + 2, dtype=torch.float32) * 0.3
        yf = y0.float().reshape(-1, 64)
        mean, var = yf.mean(0), yf.var(0, unbiased=False)
        invstd = (var + 1e-5).rsqrt()
        coef = torch.stack((mean, invstd, gamma * invstd, beta - mean * gamma * invstd)).contiguous()
        out, am = ops.stem_bn_gelu_pool_fwd(y0, coef)
        dout = randn(*out.shape, seed=seed + 3)
        dc, dgamma, dbeta = ops.stem_bwd_fused(dout, am, y0, coef)
        # (a) the three-pass path it replaces (dz rounded to bf16 in between)
        dz = ops.stem_pool_gelu_bwd(dout, am, y0, coef)
        dc3, dg3, db3, _ = ops.batchnorm_bwd(dz, None, y0, coef)
        assert rel(dc, dc3) < BF16_TOL and rel(dgamma, dg3) < 2e-3 and rel(dbeta, db3) < 2e-3
        # (b) autograd on the same (bf16-valued) inputs
        x = y0.float().permute(0, 3, 1, 2).requires_grad_(True)
        g, b = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        ref = F.max_pool2d(F.gelu(F.batch_norm(x, None, None, g, b, True, 0.1, 1e-5)), 3, 2, 1)
        gx, gg, gb = torch.autograd.grad(ref, (x, g, b), dout.float().permute(0, 3, 1, 2))
        assert rel(dc, gx.permute(0, 2, 3, 1)) < BF16_TOL
        assert rel(dgamma, gg) < 2e-3 and rel(dbeta, gb) < 2e-3


def test_meanpool_cls(ops):
    B, T, C = 3, 29, 512
    a = randn(B * T, 3, 3, C, seed=40)
    cls = randn(C, seed=41, dtype=torch.float32)
    xs = ops.meanpool_cls_fwd(a, cls, B, T)
    ref = torch.cat((cls.expand(B, 1, C), a.float().mean((1, 2)).view(B, T, C)), 1)
    assert rel(xs, ref) < 1e-5
    dx = randn(B, T + 1, C, seed=42, dtype=torch.float32)
    dout, dcls = ops.meanpool_cls_bwd(dx, 9)
    assert rel(dout, (dx[:, 1:] / 9).reshape(B * T, 1, C).expand(B * T, 9, C)) < BF16_TOL
    assert rel(dcls, dx[:, 0].sum(0)) < 1e-5


# ---------------------------------------------------------------- encoder pieces ---------------------------------
def test_rmsnorm_fwd_bwd(ops):
    from oracle.lrw_oracle import rmsnorm

    M, D = 300, 512
    x = randn(M, D, seed=50, dtype=torch.float32) * 3
    g = randn(D, seed=51, dtype=torch.float32).abs() + 0.5
    y, inv = ops.rmsnorm_fwd(x, g)
    xr, gr = x.clone().requires_grad_(True), g.clone().requires_grad_(True)
    ref = rmsnorm(xr, gr)
    assert rel(y, ref) < BF16_TOL
    dy = randn(M, D, seed=52)
    gx, gg = torch.autograd.grad(ref, (xr, gr), dy.float())
    dx = randn(M, D, seed=53, dtype=torch.float32)
    dx0 = dx.clone()
    dxb, dg = ops.rmsnorm_bwd(dy, x, g, inv, dx)
    assert rel(dx - dx0, gx) < 1e-4 and rel(dg, gg) < 1e-4
    assert rel(dxb, dx) < BF16_TOL


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("B,n,rotary_v", [(3, 30, True), (3, 30, False), (5, 17, True), (2, 32, True), (3, 50, True),
                                          (2, 64, False), (67, 30, True)])
def test_attention_fwd_bwd(ops, monkeypatch, B, n, rotary_v, tc):
    """x-transformers attention core (rotary q/k[/v], softmax(QK^T/8)V) and its backward: the tcgen05 kernels of
    attention_tc.cu (default; 128/32 or 128/64 (batch, head) pairs per UMMA tile, bf16 operands incl. the probabilities)
    and the fp32 CUDA-core kernels they replace (SVSR_ATTN_TC=0), both against fp32 PyTorch."""
    from oracle.lrw_oracle import _rotary, rotary_table

    monkeypatch.setenv("SVSR_ATTN_TC", "1" if tc else "0")
    H = 8
    qkv = randn(B * n, 3 * H * 64, seed=60)
    rot = ops.rotary_table(n)
    o = ops.attention_fwd(qkv, rot, B, n, H, rotary_v)
    t = qkv.float().requires_grad_(True)
    q, k, v = (t[:, i * 512:(i + 1) * 512].view(B, n, H, 64).transpose(1, 2) for i in range(3))
    fr = rotary_table(n).cuda()
    q, k = _rotary(q, fr), _rotary(k, fr)
    if rotary_v:
        v = _rotary(v, fr)
    attn = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1)
    ref = (attn @ v).transpose(1, 2).reshape(B * n, H * 64)
    tol = 8e-3 if tc else BF16_TOL  # tensor-core path: rotated q/k/v and the probabilities are bf16 MMA operands
    assert rel(o, ref) < tol
    d_o = randn(B * n, H * 64, seed=61)
    (gq,) = torch.autograd.grad(ref, t, d_o.float())
    dqkv = ops.attention_bwd(qkv, rot, d_o, B, n, H, rotary_v)
    assert rel(dqkv, gq) < (1.2e-2 if tc else BF16_TOL)
    for i, name in enumerate("qkv"):
        assert rel(dqkv[:, i * 512:(i + 1) * 512], gq[:, i * 512:(i + 1) * 512]) < (1.5e-2 if tc else 2 * BF16_TOL), name


@pytest.mark.parametrize("B,n,K", [(64, 30, 512), (5, 30, 576), (3, 17, 128)])
def test_fused_qkv_projection_attention_forward(ops, B, n, K):
    """attention_qkv_tc_fwd (projection + rotary + softmax + PV in one kernel) against the two-launch path: the GEMM's qkv
    buffer and the tcgen05 attention core on it; clip counts that are not a multiple of the 4-clip tile, the 576-column
    pitch of the word-boundary variant, short sequences."""
    heads = 8
    xn = randn(B * n, K, seed=61, scale=1.0)
    w = randn(3 * heads * 64, K, seed=62, scale=K ** -0.5)
    rot = ops.rotary_table(n)
    qkv_ref = ops.gemm(xn, w)
    o_ref = ops.attention_fwd(qkv_ref, rot, B, n, heads)
    qkv, o = ops.attention_qkv_fwd(xn, w, rot, B, n, heads)
    assert rel(qkv, qkv_ref) < 1e-3  # same tcgen05 accumulation, one bf16 rounding on both sides
    assert rel(o, o_ref) < BF16_TOL


def test_geglu(ops):
    h = randn(500, 4096, seed=70)
    hr = h.float().requires_grad_(True)
    a, g = hr.chunk(2, -1)
    ref = a * F.gelu(g)
    assert rel(ops.geglu_fwd(h), ref) < BF16_TOL
    du = randn(500, 2048, seed=71)
    (gh,) = torch.autograd.grad(ref, hr, du.float())
    assert rel(ops.geglu_bwd(h, du), gh) < BF16_TOL


def test_geglu_dropout_mask_is_consistent_between_forward_and_backward(ops):
    """Dropout(ff_dropout) after the GLU (x-transformers FeedForward): kept elements are scaled by 1/(1-p), the
    backward regenerates the identical mask from (seed, index)."""
    p, seed = 0.3, 1234567
    h = randn(640, 4096, seed=72)
    u0 = ops.geglu_fwd(h).float()
    u = ops.geglu_fwd(h, p, seed).float()
    keep = u != 0
    frac = 1.0 - keep.float().mean().item()
    assert abs(frac - p) < 5e-3
    assert rel(u[keep], (u0 / (1 - p))[keep]) < BF16_TOL
    assert not torch.equal(ops.geglu_fwd(h, p, seed + 1).float() != 0, keep)  # another seed, another mask
    assert torch.equal(ops.geglu_fwd(h, p, seed).float(), u)  # deterministic
    du = randn(640, 2048, seed=73)
    dh = ops.geglu_bwd(h, du, p, seed).float()
    dh_ref = ops.geglu_bwd(h, (du.float() * keep / (1 - p)).bfloat16()).float()
    assert rel(dh, dh_ref) < BF16_TOL
    assert torch.equal(dh[:, :2048][~keep], torch.zeros_like(dh[:, :2048][~keep]))


# ---------------------------------------------------------------- loss heads -------------------------------------
@pytest.mark.parametrize("A,G,V,extra", [(4, 2, 320, 0), (2, 2, 320, 5), (2, 2, 640, 1), (4, 8, 1024, 3)])
def test_audio_ce_matches_reference_indexing(ops, A, G, V, extra):
    """lightning.py:168-171: F.cross_entropy(logits.reshape(-1,V), audio_tokens[:, :T*A].flatten())."""
    B, T = 3, 29
    logits = randn(B * T, A * G * V, seed=80, dtype=torch.float32) * 2
    g = torch.Generator(device="cuda").manual_seed(81)
    tokens = torch.randint(0, V, (B, T * A + extra, G), device="cuda", generator=g)
    lr = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr.reshape(B, T, A * G, V).reshape(-1, V), tokens[:, : T * A].flatten())
    loss, dl, bad = ops.audio_ce(logits, tokens, T, A, G, V, dscale=1.0 / (B * T * A * G))
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    (gl,) = torch.autograd.grad(ref, lr)
    assert rel(dl, gl) < BF16_TOL
    assert bad.item() == 0
    # the target of every row is reproduced bit-exactly: gradient is negative exactly at the reference's target class
    picked = (dl.float().view(-1, V) < 0).float().argmax(-1)
    assert torch.equal(picked, tokens[:, : T * A].flatten())


def test_audio_ce_flags_out_of_range_tokens(ops):
    logits = randn(29, 2560, seed=82, dtype=torch.float32)
    tokens = torch.zeros(1, 116, 2, dtype=torch.int64, device="cuda")
    tokens[0, 5, 1] = 320
    _, _, bad = ops.audio_ce(logits, tokens, 29, 4, 2, 320)
    assert bad.item() == 1


@pytest.mark.parametrize("smoothing", [0.0, 0.1])
@pytest.mark.parametrize("soft", [False, True])
def test_category_ce(ops, smoothing, soft):
    B, C, ld = 37, 500, 512
    logits = torch.zeros(B, ld, device="cuda")
    logits[:, :C] = randn(B, C, seed=90, dtype=torch.float32) * 3
    g = torch.Generator(device="cuda").manual_seed(91)
    hard = torch.randint(0, C, (B,), device="cuda", generator=g)
    if soft:  # CutMix-style mix of two one-hots (augment.py)
        other = torch.randint(0, C, (B,), device="cuda", generator=g)
        lam = torch.rand(B, 1, device="cuda", generator=g) * 0.4 + 0.6
        labels = F.one_hot(hard, C).float() * lam + F.one_hot(other, C).float() * (1 - lam)
    else:
        labels = hard
    lr = logits[:, :C].clone().requires_grad_(True)
    ref = F.cross_entropy(lr, labels, label_smoothing=smoothing)
    loss, top1, top5, dl = ops.category_ce(logits, labels, C, smoothing, dscale=1.0 / B)
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    (gl,) = torch.autograd.grad(ref, lr)
    assert rel(dl[:, :C], gl) < BF16_TOL
    tgt = labels.argmax(-1) if soft else labels
    corrects = lr.topk(5, dim=1)[1] == tgt.unsqueeze(1)
    assert abs(top1.item() - corrects[:, 0].float().mean().item()) < 1e-6
    assert abs(top5.item() - corrects.float().amax(1).mean().item()) < 1e-6
