"""Kernel-level parity: every C-ABI entry point against plain PyTorch fp32 ops on the same (bf16-rounded) inputs.

Tolerances are stated per test. GEMM operands are bf16 (exactly representable inputs are fed to both sides), the
accumulation is fp32 on both sides, so the only difference is summation order: |err| <= 2e-3 * max|ref|.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def ops():
    from tempo_vae_b200 import ops as o
    return o


@pytest.fixture(params=["cta_pair", "single_cta"])
def conv_sched(request):
    """Run a conv test under both schedules of tvae_conv_gemm (cta_group::2 CTA pairs / one CTA per SM)."""
    from tempo_vae_b200._lib import lib
    prev = lib.tvae_conv_set_cta_pair(1 if request.param == "cta_pair" else 0)
    yield request.param
    lib.tvae_conv_set_cta_pair(prev)


@pytest.fixture(params=["cta_pair", "single_cta"])
def wgrad_sched(request):
    """Both schedules of tvae_wgrad_gemm (CTA pairs over M tiles + single-CTA remainder launch / one CTA per SM)."""
    from tempo_vae_b200._lib import lib
    prev = lib.tvae_wgrad_set_cta_pair(1 if request.param == "cta_pair" else 0)
    yield request.param
    lib.tvae_wgrad_set_cta_pair(prev)


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def nhwc_bf16(x_nchw, pitch):
    N, C, H, W = x_nchw.shape
    out = torch.zeros((N, H, W, pitch), dtype=torch.bfloat16, device=x_nchw.device)
    out[..., :C] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


CONV_CASES = [
    # N, H, W, Cin, Cout, R
    (2, 16, 16, 64, 64, 3),
    (2, 16, 16, 128, 128, 1),
    (3, 32, 32, 256, 256, 3),
    (2, 64, 64, 128, 512, 3),
    (1, 64, 64, 1028, 512, 3),
    (1, 64, 64, 512, 1028, 3),
    (2, 16, 16, 32, 128, 3),
    (2, 16, 16, 128, 64, 3),
    (2, 16, 16, 512, 4, 1),
    (5, 8, 8, 64, 48, 3),
    (3, 4, 4, 32, 32, 3),
    (1, 8, 256, 64, 64, 3),
]


@pytest.mark.usefixtures("conv_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout,R", CONV_CASES)
def test_conv_fwd(N, H, W, Cin, Cout, R):
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    w = bf16_round(torch.randn((Cout, Cin, R, R), device="cuda", generator=g) / math.sqrt(Cin * R * R))
    b = torch.randn((Cout,), device="cuda", generator=g)
    res = torch.randn((N, Cout, H, W), device="cuda", generator=g)
    ref = F.conv2d(x, w, b, padding=R // 2) + res
    xp = nhwc_bf16(x, o.round_up(Cin, 8))
    wp = o.pack_weight(w, "fwd")
    res_nhwc = torch.zeros((N, H, W, o.round_up(Cout, 4)), device="cuda")
    res_nhwc[..., :Cout] = res.permute(0, 2, 3, 1)
    of, ob = o.conv_gemm(xp, Cin, wp, kind=0, R=R, Cout=Cout, bias=b, residual=res_nhwc, want_f32=True, want_bf16=True)
    torch.cuda.synchronize()
    got = of[..., :Cout].permute(0, 3, 1, 2)
    assert rel_err(got, ref) < 2e-3
    gotb = ob[..., :Cout].float().permute(0, 3, 1, 2)
    assert rel_err(gotb, ref) < 1e-2
    # bf16-only output WITH a bias and nothing else: the LEAN epilogue instantiation on a forward layer (inference paths
    # whose geometry rules out fused statistics take it; a dropped bias there once went unnoticed by this file)
    _, ob2 = o.conv_gemm(xp, Cin, wp, kind=0, R=R, Cout=Cout, bias=b, want_f32=False, want_bf16=True)
    torch.cuda.synchronize()
    assert rel_err(ob2[..., :Cout].float().permute(0, 3, 1, 2), ref - res) < 1e-2


@pytest.mark.usefixtures("conv_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout,loss_type", [(2, 16, 16, 64, 132, 0), (3, 16, 16, 64, 132, 1), (2, 64, 64, 128, 1028, 0),
                                                       (3, 8, 8, 64, 48, 0)])
def test_conv_fused_reconstruction_loss(N, H, W, Cin, Cout, loss_type):
    """tvae_conv_args.nll_*: decoder.conv_out with the reconstruction loss in its epilogue (sums + the gradient wrt the
    reconstruction as the bf16 output, pad lanes zeroed) against conv -> fp32 -> tvae_nll_fwd on the same operands. The
    gradient must be IDENTICAL (same fp32 d = conv + bias - target in both paths); the sums differ only by summation order.
    Ragged channel counts (132, 1028 = 4 x 208 + 196, 48) and a ragged last pixel tile are covered."""
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    w = bf16_round(torch.randn((Cout, Cin, 3, 3), device="cuda", generator=g) / math.sqrt(Cin * 9))
    b = torch.randn((Cout,), device="cuda", generator=g)
    xp = nhwc_bf16(x, o.round_up(Cin, 8))
    wp = o.pack_weight(w, "fwd")
    pitch = o.round_up(Cout, 8)
    target = torch.randn((N, H, W, pitch), device="cuda", generator=g).to(torch.bfloat16)
    logvar = torch.tensor([0.3], device="cuda")
    of, _ = o.conv_gemm(xp, Cin, wp, kind=0, R=3, Cout=Cout, bias=b, want_f32=True)
    sums_ref, dx_ref = o.nll_fwd(target[..., :Cout], of, Cout, loss_type, logvar, N, True)
    nll = {"x": target[..., :Cout], "loss_type": loss_type, "logvar": logvar, "batch": N}
    out = torch.full((N, H, W, pitch), float("nan"), device="cuda", dtype=torch.bfloat16)
    _, dx = o.conv_gemm(xp, Cin, wp, kind=0, R=3, Cout=Cout, bias=b, want_f32=False, want_bf16=True, out_bf16=out, nll=nll)
    torch.cuda.synchronize()
    assert dx.shape == dx_ref.shape and torch.equal(dx[..., :Cout], dx_ref[..., :Cout])
    assert (dx[..., Cout:] == 0).all()
    rel = ((nll["sums"] - sums_ref).abs() / sums_ref.abs().clamp_min(1e-30)).max().item()
    assert rel < 1e-6, (nll["sums"], sums_ref)
    assert float(nll["sums"][2]) == 0.0


@pytest.mark.usefixtures("conv_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout,R", [(2, 16, 16, 64, 128, 3), (2, 32, 32, 256, 512, 3), (1, 64, 64, 1028, 512, 3),
                                               (2, 16, 16, 128, 128, 1)])
def test_conv_dgrad(N, H, W, Cin, Cout, R):
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(2)
    w = bf16_round(torch.randn((Cout, Cin, R, R), device="cuda", generator=g) / math.sqrt(Cout * R * R))
    dy = bf16_round(torch.randn((N, Cout, H, W), device="cuda", generator=g))
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w, dy, padding=R // 2)
    dyp = nhwc_bf16(dy, o.round_up(Cout, 8))
    wp = o.pack_weight(w, "dgrad")
    of, _ = o.conv_gemm(dyp, Cout, wp, kind=0, R=R, Cout=Cin, flip=True)
    torch.cuda.synchronize()
    assert rel_err(of[..., :Cin].permute(0, 3, 1, 2), ref) < 2e-3


@pytest.mark.usefixtures("conv_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (2, 64, 64, 512, 512), (3, 32, 32, 256, 256), (2, 8, 8, 32, 16)])
def test_conv_down_and_dgrad(N, H, W, Cin, Cout):
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    w = bf16_round(torch.randn((Cout, Cin, 2, 2), device="cuda", generator=g) / math.sqrt(Cin * 4))
    b = torch.randn((Cout,), device="cuda", generator=g)
    ref = F.conv2d(x, w, b, stride=2)
    xp = nhwc_bf16(x, o.round_up(Cin, 8))
    of, _ = o.conv_gemm(xp, Cin, o.pack_weight(w, "fwd"), kind=1, R=2, Cout=Cout, bias=b)
    torch.cuda.synchronize()
    assert rel_err(of[..., :Cout].permute(0, 3, 1, 2), ref) < 2e-3
    # dgrad of the down conv == transposed-conv GEMM with scatter
    dy = bf16_round(torch.randn((N, Cout, H // 2, W // 2), device="cuda", generator=g))
    refd = torch.nn.grad.conv2d_input((N, Cin, H, W), w, dy, stride=2)
    dyp = nhwc_bf16(dy, o.round_up(Cout, 8))
    od, _ = o.conv_gemm(dyp, Cout, o.pack_weight(w, "down_dgrad"), kind=2, R=2, Cout=Cin)
    torch.cuda.synchronize()
    assert rel_err(od[..., :Cin].permute(0, 3, 1, 2), refd) < 2e-3


@pytest.mark.usefixtures("conv_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 16, 16, 128, 256), (2, 32, 32, 256, 512), (2, 8, 8, 32, 16)])
def test_convT_up_and_dgrad(N, H, W, Cin, Cout):
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(4)
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    w = bf16_round(torch.randn((Cin, Cout, 2, 2), device="cuda", generator=g) / math.sqrt(Cin))
    b = torch.randn((Cout,), device="cuda", generator=g)
    ref = F.conv_transpose2d(x, w, b, stride=2)
    xp = nhwc_bf16(x, o.round_up(Cin, 8))
    of, _ = o.conv_gemm(xp, Cin, o.pack_weight(w, "up_fwd"), kind=2, R=2, Cout=Cout, bias=b)
    torch.cuda.synchronize()
    assert rel_err(of[..., :Cout].permute(0, 3, 1, 2), ref) < 2e-3
    dy = bf16_round(torch.randn((N, Cout, 2 * H, 2 * W), device="cuda", generator=g))
    refd = F.conv2d(dy, w, stride=2)  # d/dx of conv_transpose2d
    dyp = nhwc_bf16(dy, o.round_up(Cout, 8))
    od, _ = o.conv_gemm(dyp, Cout, o.pack_weight(w, "up_dgrad"), kind=1, R=2, Cout=Cin)
    torch.cuda.synchronize()
    assert rel_err(od[..., :Cin].permute(0, 3, 1, 2), refd) < 2e-3


@pytest.mark.usefixtures("wgrad_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout,R,splits", [
    (2, 16, 16, 64, 64, 3, 1), (2, 16, 16, 128, 128, 1, 0), (4, 32, 32, 256, 256, 3, 0), (2, 64, 64, 512, 512, 3, 0),
    (1, 64, 64, 1028, 512, 3, 0), (1, 64, 64, 512, 1028, 3, 3), (2, 16, 16, 128, 64, 3, 2), (2, 16, 16, 512, 4, 1, 0),
    (6, 4, 4, 32, 32, 3, 0), (2, 16, 16, 32, 128, 3, 40), (2, 16, 16, 64, 384, 3, 0), (2, 16, 16, 96, 200, 3, 2),
    (3, 8, 8, 48, 648, 1, 0)])
def test_wgrad(N, H, W, Cin, Cout, R, splits):
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    dy = bf16_round(torch.randn((N, Cout, H, W), device="cuda", generator=g))
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, R, R), dy, padding=R // 2)
    grad = torch.full((Cout, Cin, R, R), float("nan"), device="cuda")
    o.wgrad_gemm(nhwc_bf16(dy, o.round_up(Cout, 8)), Cout, nhwc_bf16(x, o.round_up(Cin, 8)), Cin, kind=0, R=R,
                 grad=grad, splits=splits)
    torch.cuda.synchronize()
    assert rel_err(grad, ref) < 2e-3
    o.wgrad_gemm(nhwc_bf16(dy, o.round_up(Cout, 8)), Cout, nhwc_bf16(x, o.round_up(Cin, 8)), Cin, kind=0, R=R,
                 grad=grad, splits=splits, accumulate=True)
    torch.cuda.synchronize()
    assert rel_err(grad, 2 * ref) < 2e-3


@pytest.mark.usefixtures("wgrad_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (2, 64, 64, 512, 512), (2, 8, 8, 32, 16)])
def test_wgrad_strided(N, H, W, Cin, Cout):
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(6)
    # 2x2 s2 Conv2d: P = dY (coarse), Q = x (fine)
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    dy = bf16_round(torch.randn((N, Cout, H // 2, W // 2), device="cuda", generator=g))
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, 2, 2), dy, stride=2)
    grad = torch.empty((Cout, Cin, 2, 2), device="cuda")
    o.wgrad_gemm(nhwc_bf16(dy, o.round_up(Cout, 8)), Cout, nhwc_bf16(x, o.round_up(Cin, 8)), Cin, kind=1, R=2, grad=grad)
    torch.cuda.synchronize()
    assert rel_err(grad, ref) < 2e-3
    # ConvTranspose2d [Cin][Cout][2][2]: P = x (coarse), Q = dY (fine)
    xc = bf16_round(torch.randn((N, Cin, H // 2, W // 2), device="cuda", generator=g))
    dyf = bf16_round(torch.randn((N, Cout, H, W), device="cuda", generator=g))
    wt = torch.zeros((Cin, Cout, 2, 2), device="cuda", requires_grad=True)
    F.conv_transpose2d(xc, wt, stride=2).backward(dyf)
    gradt = torch.empty((Cin, Cout, 2, 2), device="cuda")
    o.wgrad_gemm(nhwc_bf16(xc, o.round_up(Cin, 8)), Cin, nhwc_bf16(dyf, o.round_up(Cout, 8)), Cout, kind=1, R=2, grad=gradt)
    torch.cuda.synchronize()
    assert rel_err(gradt, wt.grad) < 2e-3


@pytest.mark.parametrize("rows,C,in_pitch", [(4096, 1028, 1028), (1000, 20, 24), (257, 7, 7), (64, 512, 640)])
def test_channels_last_cast_and_radiance_normalisation(rows, C, in_pitch):
    o = ops()
    from tempo_vae_b200._lib import lib
    g = torch.Generator(device="cuda").manual_seed(23)
    buf = torch.randn((rows, in_pitch), device="cuda", generator=g)
    pitch = o.round_up(C, 8)
    out = torch.full((rows, pitch), float("nan"), device="cuda", dtype=torch.bfloat16)
    o.check(lib.tvae_nhwc_f32_to_nhwc_bf16(buf.data_ptr(), in_pitch, rows, C, out.data_ptr(), pitch, None, None),
            "tvae_nhwc_f32_to_nhwc_bf16")
    torch.cuda.synchronize()
    assert torch.equal(out[:, :C], buf[:, :C].to(torch.bfloat16)) and (out[:, C:] == 0).all()
    # z-scored log radiance, fp32 and bf16-operand outputs from one pass
    rad = torch.exp(torch.randn((rows, C), device="cuda", generator=g) * 2.0 + 1.0)
    mean = torch.randn((C,), device="cuda", generator=g) + 1.0
    std = torch.rand((C,), device="cuda", generator=g) + 0.5
    zf, zb = o.normalize_radiance(rad, mean, std, 1.0, -3.0, 3.0, want_f32=True, want_bf16=True)
    ref = torch.clamp((torch.log(torch.clamp(rad, 1.0, float("inf"))) - mean) / (std + 1e-8), -3.0, 3.0)
    torch.cuda.synchronize()
    assert (zf - ref).abs().max() < 1e-5
    assert torch.equal(zb[:, :C], zf.to(torch.bfloat16)) and (zb[:, C:] == 0).all()
    assert float(zf.max()) <= 3.0 and float(zf.min()) >= -3.0


def test_layout_roundtrip():
    o = ops()
    x = torch.randn((3, 1028, 16, 16), device="cuda")
    y = o.nchw_to_nhwc_bf16(x, 1032)
    assert torch.equal(y[..., :1028], x.permute(0, 2, 3, 1).to(torch.bfloat16))
    assert (y[..., 1028:] == 0).all()
    z = o.nhwc_to_nchw_f32(y, 1028)
    assert torch.equal(z, x.to(torch.bfloat16).float())
    f = torch.randn((3, 16, 16, 1028), device="cuda")
    assert torch.equal(o.nhwc_to_nchw_f32(f, 1028), f.permute(0, 3, 1, 2))
    assert torch.equal(o.f32_to_bf16(f), f.to(torch.bfloat16))
    # vectorised and scalar transpose paths, partial tiles, pad lanes
    for (N, C, H, W) in [(1, 1028, 64, 64), (2, 20, 16, 16), (2, 70, 12, 12), (3, 7, 5, 5), (2, 130, 4, 8)]:
        x = torch.randn((N, C, H, W), device="cuda")
        pitch = o.round_up(C, 8)
        y = o.nchw_to_nhwc_bf16(x)
        assert y.shape == (N, H, W, pitch)
        assert torch.equal(y[..., :C], x.permute(0, 2, 3, 1).to(torch.bfloat16)), (N, C, H, W)
        assert (y[..., C:] == 0).all()


@pytest.mark.parametrize("N,H,W,C,G,act,eps", [(3, 16, 16, 128, 8, 1, 1e-6), (2, 64, 64, 512, 8, 1, 1e-6), (2, 16, 16, 128, 8, 0, 1e-6),
                                                (2, 8, 8, 32, 8, 1, 1e-5), (2, 4, 4, 16, 8, 1, 1e-6)])
def test_groupnorm_gelu_fwd_bwd(N, H, W, C, G, act, eps):
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.randn((N, C, H, W), device="cuda", generator=g) * 1.7 + 0.3).requires_grad_(True)
    gamma = (torch.randn((C,), device="cuda", generator=g) * 0.5 + 1.0).requires_grad_(True)
    beta = (torch.randn((C,), device="cuda", generator=g) * 0.2).requires_grad_(True)
    y = F.group_norm(x, G, gamma, beta, eps)
    if act:
        y = F.gelu(y)
    da = bf16_round(torch.randn((N, C, H, W), device="cuda", generator=g))
    gres = bf16_round(torch.randn((N, C, H, W), device="cuda", generator=g))
    y.backward(da)
    xn = x.detach().permute(0, 2, 3, 1).contiguous()
    stats = o.gn_stats(xn, C, G, eps)
    a = o.gn_act_fwd(xn, stats, gamma.detach(), beta.detach(), G, act)
    torch.cuda.synchronize()
    ref_stats_mean = x.detach().reshape(N, G, -1).mean(-1)
    assert torch.allclose(stats[..., 0], ref_stats_mean, atol=1e-5)
    assert rel_err(a.float().permute(0, 3, 1, 2), y.detach()) < 1e-2  # bf16 output rounding
    dgamma = torch.empty((C,), device="cuda"); dbeta = torch.empty((C,), device="cuda")
    colsum = torch.full((C,), float("nan"), device="cuda")
    dx = o.gn_act_bwd(xn, stats, gamma.detach(), beta.detach(), da.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16),
                      gres.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), G, act, dgamma, dbeta, colsum)
    torch.cuda.synchronize()
    assert rel_err(dx.float().permute(0, 3, 1, 2), x.grad + gres) < 1e-2
    assert rel_err(dgamma, gamma.grad) < 1e-3
    assert rel_err(dbeta, beta.grad) < 1e-3
    # fused column sums of dx (bias gradient of the producing conv), accumulated before dx is rounded to bf16:
    # against the fp32 reference, and within bf16 rounding noise of the sums of the stored values
    ref_cs = (x.grad + gres).double().sum(dim=(0, 2, 3))
    scale = (x.grad + gres).double().abs().sum(dim=(0, 2, 3)).max()
    tol = 1e-4 if o.gn_fast_ok(C, G) else 4e-3        # generic geometries sum the stored (bf16-rounded) dx instead
    assert (colsum.double() - ref_cs).abs().max() <= tol * scale + 1e-6
    assert (colsum.double() - dx.double().sum(dim=(0, 1, 2))).abs().max() <= 4e-3 * scale
    # and the entry point still works without it
    dx2 = o.gn_act_bwd(xn, stats, gamma.detach(), beta.detach(), da.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16),
                       gres.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), G, act, dgamma, dbeta)
    torch.cuda.synchronize()
    assert torch.equal(dx, dx2)


@pytest.mark.parametrize("N,H,W,C,G,act", [(3, 16, 16, 128, 8, 1), (2, 64, 64, 512, 8, 1), (2, 8, 8, 256, 8, 0)])
def test_groupnorm_bf16_input(N, H, W, C, G, act):
    """The same entry points with a bf16 input tensor (a conv output that only feeds a norm): identical results to the
    fp32 path fed the same (bf16-representable) values."""
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(17)
    xn = bf16_round(torch.randn((N, H, W, C), device="cuda", generator=g) * 1.7 + 0.3)
    gamma = torch.randn((C,), device="cuda", generator=g) * 0.5 + 1.0
    beta = torch.randn((C,), device="cuda", generator=g) * 0.2
    da = torch.randn((N, H, W, C), device="cuda", generator=g).to(torch.bfloat16)
    gres = torch.randn((N, H, W, C), device="cuda", generator=g).to(torch.bfloat16)
    stats = o.gn_stats(xn, C, G, 1e-6)
    xb = xn.to(torch.bfloat16)
    a32 = o.gn_act_fwd(xn, stats, gamma, beta, G, act)
    a16 = o.gn_act_fwd(xb, stats, gamma, beta, G, act)
    outs = []
    for xin in (xn, xb):
        dg, db, cs = (torch.empty((C,), device="cuda") for _ in range(3))
        dx = o.gn_act_bwd(xin, stats, gamma, beta, da, gres, G, act, dg, db, cs)
        outs.append((dx, dg, db, cs))
    torch.cuda.synchronize()
    assert torch.equal(a32, a16)
    for u, v in zip(outs[0], outs[1]):
        assert torch.equal(u, v)
    assert o.gn_fast_ok(C, G) and not o.gn_fast_ok(4, 2)
    with pytest.raises(Exception):      # geometry outside the vectorised kernels: bf16 input is refused, not converted
        o.gn_act_fwd(torch.zeros((1, 4, 4, 4), device="cuda", dtype=torch.bfloat16),
                     torch.zeros((1, 2, 2), device="cuda"), torch.ones(4, device="cuda"), torch.zeros(4, device="cuda"), 2, 1)


def test_wgrad_sub_block_split_matches_single_gemm():
    """tvae_wgrad_args.grad_ld / grad_off: a 132-channel weight gradient computed as a 128-channel GEMM (whole tile) plus a
    4-channel skinny GEMM with the leftover channels on the N side, on the input-channel side (conv_in-like, written with
    an inner-dimension offset) and on the output-channel side (conv_out-like, a contiguous row block), against the single
    padded GEMM."""
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    N, H, W, Cw, Cn = 2, 16, 16, 132, 128
    xw = torch.randn((N, H, W, 136), device="cuda", generator=g).to(torch.bfloat16)        # 132 channels, pitch 136
    dy = torch.randn((N, H, W, Cn), device="cuda", generator=g).to(torch.bfloat16)
    # conv_in-like: weight [Cout=128][Cin=132][3][3]
    ref = torch.empty((Cn, Cw, 3, 3), device="cuda")
    o.wgrad_gemm(xw[..., :Cw], Cw, dy, Cn, kind=0, R=3, grad=ref, flip=True)
    got = torch.full((Cn, Cw, 3, 3), float("nan"), device="cuda")
    o.wgrad_gemm(xw[..., :128], 128, dy, Cn, kind=0, R=3, grad=got, flip=True, grad_ld=Cw, grad_off=0)
    o.wgrad_gemm(dy, Cn, xw[..., 128:Cw], Cw - 128, kind=0, R=3, grad=got, grad_ld=Cw, grad_off=128)
    torch.cuda.synchronize()
    assert torch.isfinite(got).all() and rel_err(got, ref) < 1e-5
    # conv_out-like: weight [Cout=132][Cin=128][3][3], dY has the 132 channels
    ref2 = torch.empty((Cw, Cn, 3, 3), device="cuda")
    o.wgrad_gemm(xw[..., :Cw], Cw, dy, Cn, kind=0, R=3, grad=ref2)
    got2 = torch.full((Cw, Cn, 3, 3), float("nan"), device="cuda")
    flat = got2.view(-1)
    o.wgrad_gemm(xw[..., :128], 128, dy, Cn, kind=0, R=3, grad=flat[:128 * Cn * 9])
    o.wgrad_gemm(dy, Cn, xw[..., 128:Cw], Cw - 128, kind=0, R=3, grad=flat[128 * Cn * 9:], flip=True)
    torch.cuda.synchronize()
    assert torch.isfinite(got2).all() and rel_err(got2, ref2) < 1e-5


@pytest.mark.parametrize("N,H,W,Cn,tail,accumulate", [(2, 16, 16, 128, 4, False), (3, 8, 24, 64, 3, True), (1, 64, 64, 512, 4, False),
                                                      (2, 5, 7, 192, 1, False), (148, 16, 16, 256, 4, True)])
def test_wgrad_skinny_tail_channels(N, H, W, Cn, tail, accumulate):
    """tvae_wgrad_skinny (taps on the M side, wide operand read once) for the leftover channels of a 128 + tail channel
    count, on the input-channel side (conv_in-like: inner-dimension offset) and the output-channel side (conv_out-like:
    contiguous row block): against fp32 autograd of F.conv2d on the same bf16 operands, run-to-run bit-identical, and with
    everything around the written block untouched. Ragged pixel counts (not a multiple of the 64-pixel chunk), non-square
    images and more chunks than CTAs are covered."""
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(N * 100 + tail)
    Cw = 128 + tail
    pitch = 136
    xw = torch.randn((N, H, W, pitch), device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn((N, H, W, Cn), device="cuda", generator=g).to(torch.bfloat16)

    def conv_wgrad(x_nhwc, dy_nhwc):      # weight gradient [Cout][Cin][3][3] of a 3x3 / pad 1 convolution, fp32
        xx = x_nhwc.float().permute(0, 3, 1, 2)
        w = torch.zeros((dy_nhwc.shape[-1], xx.shape[1], 3, 3), device="cuda", requires_grad=True)
        torch.nn.functional.conv2d(xx, w, padding=1).backward(dy_nhwc.float().permute(0, 3, 1, 2))
        return w.grad
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        ref_in = conv_wgrad(xw[..., :Cw], dy)                 # [Cn][Cw][3][3]
        ref_out = conv_wgrad(dy, xw[..., :Cw])                # [Cw][Cn][3][3]: xw plays dY here, dy plays x
    base = 0.5 if accumulate else float("nan")
    # conv_in-like: the tail input channels of dW[n][128 + c][tap]
    got = torch.full((Cn, Cw, 3, 3), base, device="cuda")
    runs = []
    for _ in range(2):
        got.fill_(base)
        o.wgrad_skinny(dy, Cn, xw[..., 128:Cw], tail, sign=+1, grad=got.view(-1)[128 * 9:], stride_c=9, stride_n=9 * Cw,
                       accumulate=accumulate)
        torch.cuda.synchronize()
        runs.append(got.clone())
    assert torch.equal(runs[0][:, 128:], runs[1][:, 128:])
    want = ref_in[:, 128:] + (base if accumulate else 0.0)
    assert rel_err(got[:, 128:], want) < 1e-4      # fp32 sums over up to 38 K pixels in two different orders
    untouched = got[:, :128]
    assert (untouched == base).all() if accumulate else torch.isnan(untouched).all()
    # conv_out-like: the tail output channels dW[128 + c][n][tap]
    got2 = torch.full((Cw, Cn, 3, 3), base, device="cuda")
    o.wgrad_skinny(dy, Cn, xw[..., 128:Cw], tail, sign=-1, grad=got2.view(-1)[128 * Cn * 9:], stride_c=9 * Cn, stride_n=9,
                   accumulate=accumulate)
    torch.cuda.synchronize()
    want2 = ref_out[128:] + (base if accumulate else 0.0)
    assert rel_err(got2[128:], want2) < 1e-4
    untouched = got2[:128]
    assert (untouched == base).all() if accumulate else torch.isnan(untouched).all()


def test_new_entry_points_refuse_bad_arguments():
    """tvae_wgrad_skinny / the fused-loss mode of tvae_conv_gemm / tvae_attn_*_tc fail with a message instead of launching
    when their preconditions do not hold (error behaviour of the C ABI: negative return code + tvae_last_error)."""
    o = ops()
    from tempo_vae_b200._lib import TvaeError
    wide = torch.zeros((1, 8, 8, 96), device="cuda", dtype=torch.bfloat16)
    skinny = torch.zeros((1, 8, 8, 8), device="cuda", dtype=torch.bfloat16)
    grad = torch.zeros((96 * 9 * 8,), device="cuda")
    with pytest.raises(TvaeError, match="multiple of 64"):
        o.wgrad_skinny(wide, 96, skinny[..., :4], 4, sign=+1, grad=grad, stride_c=9, stride_n=9 * 8)
    wide = torch.zeros((1, 8, 8, 64), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(TvaeError, match="1..4 skinny"):
        o.wgrad_skinny(wide, 64, skinny[..., :5], 5, sign=+1, grad=grad, stride_c=9, stride_n=9 * 8)
    with pytest.raises(TvaeError, match="shift_sign"):
        o.wgrad_skinny(wide, 64, skinny[..., :4], 4, sign=0, grad=grad, stride_c=9, stride_n=9 * 8)
    # fused loss: only a stride-1 forward conv with the bf16 output alone may carry it
    x = torch.zeros((1, 8, 8, 64), device="cuda", dtype=torch.bfloat16)
    w = torch.zeros((64, 64, 2, 2), device="cuda")
    nll = {"x": x, "loss_type": 0, "logvar": torch.zeros(1, device="cuda"), "batch": 1}
    with pytest.raises((TvaeError, AssertionError)):
        o.conv_gemm(x, 64, o.pack_weight(w, "fwd"), kind=1, R=2, Cout=64, want_f32=False, want_bf16=True, nll=nll)
    # attention: the tensor-core entry points are for head dimension 32 only
    q = torch.zeros((16, 3 * 64), device="cuda")
    out = torch.zeros((16, 64), device="cuda")
    lse = torch.zeros((1, 4, 16), device="cuda")
    rc = o.lib.tvae_attn_fwd_tc(q.data_ptr(), q.data_ptr() + 256, q.data_ptr() + 512, 192, 1, 16, 64, 4, 0, out.data_ptr(),
                                lse.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc != 0 and b"head dimension must be 32" in o.lib.tvae_last_error()


@pytest.mark.parametrize("N,H,W,C,act,x_bf16", [(3, 16, 16, 128, 1, False), (2, 64, 64, 512, 1, True), (2, 8, 8, 256, 3, False),
                                                 (2, 16, 16, 128, 2, False)])
def test_groupnorm_saved_activation_gradient(N, H, W, C, act, x_bf16):
    """tvae_gn_act_fwd2 / _bwd2: the forward also stores act'(y) (bf16) and the backward starts from it instead of
    re-evaluating the activation. The stored derivative is checked against torch, the backward against the recomputing
    path (same inputs) and against fp32 autograd."""
    o = ops()
    G = 8
    g = torch.Generator(device="cuda").manual_seed(31)
    x = bf16_round(torch.randn((N, H, W, C), device="cuda", generator=g) * 1.7 + 0.3)
    gamma = torch.randn((C,), device="cuda", generator=g) * 0.5 + 1.0
    beta = torch.randn((C,), device="cuda", generator=g) * 0.2
    da = torch.randn((N, H, W, C), device="cuda", generator=g).to(torch.bfloat16)
    gres = torch.randn((N, H, W, C), device="cuda", generator=g).to(torch.bfloat16)
    stats = o.gn_stats(x, C, G, 1e-6)
    xin = x.to(torch.bfloat16) if x_bf16 else x
    a_plain = o.gn_act_fwd(xin, stats, gamma, beta, G, act)
    a, gp = o.gn_act_fwd(xin, stats, gamma, beta, G, act, want_act_grad="force")     # (off by default in the model)
    assert gp is not None and gp.dtype == torch.bfloat16 and gp.shape == a.shape
    assert torch.equal(a, a_plain)
    fn = {1: F.gelu, 2: F.relu, 3: F.silu}[act]
    y = F.group_norm(x.permute(0, 3, 1, 2), G, gamma, beta, 1e-6).detach().requires_grad_(True)
    fn(y).sum().backward()
    ref_gp = y.grad.permute(0, 2, 3, 1)
    if act == 2:     # ReLU': ignore the elements whose pre-activation rounds across 0
        keep = y.detach().permute(0, 2, 3, 1).abs() > 1e-3
        assert torch.equal(gp.float()[keep], ref_gp[keep])
    else:
        assert float((gp.float() - ref_gp).abs().max()) < 8e-3          # bf16 storage of a value in [-0.2, 1.2]
    outs = []
    for use in (None, gp):
        dg, db, cs = (torch.full((C,), float("nan"), device="cuda") for _ in range(3))
        dx = o.gn_act_bwd(xin, stats, gamma, beta, da, gres, G, act, dg, db, cs, use)
        outs.append((dx, dg, db, cs))
    torch.cuda.synchronize()
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    fn(F.group_norm(xr, G, gr, br, 1e-6)).backward(da.float().permute(0, 3, 1, 2))
    ref_dx = xr.grad.permute(0, 2, 3, 1) + gres.float()
    for dx, dg, db, cs in outs:
        assert rel_err(dx.float(), ref_dx) < 1e-2
        assert rel_err(dg, gr.grad) < 3e-3 and rel_err(db, br.grad) < 3e-3
    assert rel_err(outs[1][0].float(), outs[0][0].float()) < 8e-3


@pytest.mark.parametrize("N,H,W,C,x_bf16,with_gres,group_mb", [(7, 64, 64, 512, False, True, 30), (5, 64, 64, 512, True, False, 20),
                                                               (37, 32, 32, 256, False, True, 8), (3, 16, 16, 128, False, True, 1)])
def test_groupnorm_bwd_single_pass_matches_two_pass(N, H, W, C, x_bf16, with_gres, group_mb):
    """The persistent single-pass GroupNorm backward (L2-resident groups, ticket + flag hand-over between its two phases)
    against the two-pass kernels on the same inputs and against fp32 autograd; group sizes chosen so that the last group
    is partial; run twice: bit-reproducible."""
    from tempo_vae_b200 import _lib
    o = ops()
    G, act = 8, 1
    g = torch.Generator(device="cuda").manual_seed(23)
    x = bf16_round(torch.randn((N, H, W, C), device="cuda", generator=g) * 1.7 + 0.3)
    gamma = torch.randn((C,), device="cuda", generator=g) * 0.5 + 1.0
    beta = torch.randn((C,), device="cuda", generator=g) * 0.2
    da = torch.randn((N, H, W, C), device="cuda", generator=g).to(torch.bfloat16)
    gres = torch.randn((N, H, W, C), device="cuda", generator=g).to(torch.bfloat16) if with_gres else None
    stats = o.gn_stats(x, C, G, 1e-6)
    xin = x.to(torch.bfloat16) if x_bf16 else x

    def run():
        dg, db, cs = (torch.full((C,), float("nan"), device="cuda") for _ in range(3))
        dx = o.gn_act_bwd(xin, stats, gamma, beta, da, gres, G, act, dg, db, cs)
        torch.cuda.synchronize()
        return dx, dg, db, cs
    try:
        _lib.lib.tvae_gn_set_bwd_fused(0, 0)
        two = run()
        _lib.lib.tvae_gn_set_bwd_fused(2, group_mb)
        one = run()
        again = run()
    finally:
        _lib.lib.tvae_gn_set_bwd_fused(0, 24)
    for u, v in zip(one, again):
        assert torch.equal(u, v)
    # fp32 autograd reference
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.gelu(F.group_norm(xr, G, gr, br, 1e-6)).backward(da.float().permute(0, 3, 1, 2))
    ref_dx = xr.grad.permute(0, 2, 3, 1) + (gres.float() if with_gres else 0)
    for got in (one, two):
        assert rel_err(got[0].float(), ref_dx) < 1e-2
        assert rel_err(got[1], gr.grad) < 1e-3 and rel_err(got[2], br.grad) < 1e-3
    scale = ref_dx.double().abs().sum(dim=(0, 1, 2)).max()
    assert (one[3].double() - ref_dx.double().sum(dim=(0, 1, 2))).abs().max() <= 1e-4 * scale + 1e-6
    # the two schedules differ in the summation order of the row sums and in ONE bf16 rounding: the two-pass kernels hand
    # dy = da * act'(y) from the row-sum pass to the apply pass as bf16 (in the dx buffer), the single pass keeps it in
    # fp32 -- a 2^-9 change of dy flips the final bf16 rounding of dx for a fraction of the elements (one ulp = 2^-8)
    assert rel_err(one[0].float(), two[0].float()) < 8e-3
    assert rel_err(one[1], two[1]) < 1e-5 and rel_err(one[2], two[2]) < 1e-5


@pytest.mark.parametrize("rows,C,pitch", [(4096, 512, 512), (8192, 1028, 1032), (512, 4, 8), (1000, 64, 64)])
def test_colsum(rows, C, pitch):
    o = ops()
    x = torch.randn((rows, pitch), device="cuda").to(torch.bfloat16)
    out = torch.empty((C,), device="cuda")
    o.colsum_bf16(x, C, out)
    torch.cuda.synchronize()
    assert rel_err(out, x[:, :C].float().sum(0)) < 1e-4


@pytest.mark.parametrize("B,T,C,heads,tc", [(3, 256, 128, 4, True), (2, 64, 16, 4, False), (1, 1024, 128, 4, True),
                                            (2, 100, 32, 4, False), (2, 100, 128, 4, True), (1, 700, 64, 2, True),
                                            (3, 256, 128, 4, False)])
def test_attention_fwd_bwd(B, T, C, heads, tc):
    """tc=True: TF32 tensor-core kernels (head dim 32; tcgen05 kind::tf32 by default, then the mma.sync kernels behind
    tvae_attn_set_tcgen05(0)), tolerance 3e-3 of the max; tc=False: exact fp32 kernels, 1e-4."""
    o = ops()
    o.ATTN_TENSOR_CORES[0] = tc
    assert o.attn_uses_tensor_cores(C, heads) == (tc and C == 32 * heads)
    try:
        assert o.lib.tvae_attn_set_tcgen05(-1) == 1          # the tcgen05 kernels are the default
        _attention_case(o, B, T, C, heads, 3e-3 if tc else 1e-4)
        if tc:
            o.lib.tvae_attn_set_tcgen05(0)
            _attention_case(o, B, T, C, heads, 3e-3)
    finally:
        o.ATTN_TENSOR_CORES[0] = True
        o.lib.tvae_attn_set_tcgen05(1)


@pytest.mark.parametrize("B,T,heads", [(2, 256, 4), (1, 200, 4), (2, 64, 1), (1, 2048, 2), (1, 1, 4), (1, 129, 3)])
def test_attention_tcgen05_matches_mma_sync(B, T, heads):
    """The two TF32 implementations (tcgen05 + TMEM vs mma.sync) agree to TF32 rounding on every output, including the
    saved log-sum-exp, on ragged / tiny / multi-tile token counts; the tcgen05 kernels are run-to-run bit-identical."""
    o = ops()
    C = 32 * heads
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = torch.randn((B * T, 3 * C), device="cuda", generator=g) * 1.5
    d_out = torch.randn((B * T, C), device="cuda", generator=g)
    res = {}
    try:
        for impl in (1, 0, 1):
            o.lib.tvae_attn_set_tcgen05(impl)
            ob, of, lse = o.attn_fwd(qkv, C, heads, B, T)
            dqkv = o.attn_bwd(qkv, of, d_out, lse, C, heads, B, T)
            torch.cuda.synchronize()
            res.setdefault(impl, []).append((ob.clone(), of.clone(), lse.clone(), dqkv.clone()))
    finally:
        o.lib.tvae_attn_set_tcgen05(1)
    a, a2, b = res[1][0], res[1][1], res[0][0]
    for x, y in zip(a, a2):
        assert torch.equal(x, y)
    assert rel_err(a[1], b[1]) < 2e-3 and rel_err(a[0].float(), b[0].float()) < 1e-2
    assert (a[2] - b[2]).abs().max().item() < 2e-3
    assert rel_err(a[3].float(), b[3].float()) < 1e-2
    assert all(torch.isfinite(t.float()).all() for t in a)


def _attention_case(o, B, T, C, heads, tol):
    g = torch.Generator(device="cuda").manual_seed(8)
    qkv = torch.randn((B * T, 3 * C), device="cuda", generator=g).requires_grad_(True)
    hd = C // heads

    def ref_attn(qkv):
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        # channel c = d*heads + h  (reference: reshape(b, c_, n_heads, hw))
        def split(t):
            return t.reshape(B, T, hd, heads).permute(0, 3, 1, 2)  # [B, heads, T, hd]
        q, k, v = split(q), split(k), split(v)
        w = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1)
        out = w @ v  # [B, heads, T, hd]
        return out.permute(0, 2, 3, 1).reshape(B * T, C)

    ref = ref_attn(qkv)
    d_out = torch.randn((B * T, C), device="cuda", generator=g)
    ref.backward(d_out)
    ob, of, lse = o.attn_fwd(qkv.detach(), C, heads, B, T)
    torch.cuda.synchronize()
    assert rel_err(of, ref.detach()) < tol
    assert rel_err(ob.float(), ref.detach()) < 1e-2
    dqkv = o.attn_bwd(qkv.detach(), of, d_out, lse, C, heads, B, T)
    torch.cuda.synchronize()
    assert rel_err(dqkv.float(), qkv.grad) < 1e-2  # bf16 output


@pytest.mark.parametrize("B,T,heads", [(2, 200, 4), (1, 129, 3), (3, 64, 1)])
def test_attention_tcgen05_writes_stay_in_bounds(B, T, heads):
    """Every output of the tcgen05 attention kernels (o_bf16, o_f32, lse, dqkv, the D workspace) is allocated inside a
    NaN-filled arena at ragged token counts; the bands around each of them must stay untouched and the payload finite."""
    o = ops()
    C = 32 * heads
    g = torch.Generator(device="cuda").manual_seed(17)
    qkv = torch.randn((B * T, 3 * C), device="cuda", generator=g)
    d_out = torch.randn((B * T, C), device="cuda", generator=g)
    band = 4096

    def arena(numel, dtype):
        a = torch.full((numel + 2 * band,), float("nan"), device="cuda", dtype=dtype)
        return a, a[band:band + numel]
    a_ob, ob = arena(B * T * C, torch.bfloat16)
    a_of, of = arena(B * T * C, torch.float32)
    a_lse, lse = arena(B * heads * T, torch.float32)
    a_dq, dqkv = arena(B * T * 3 * C, torch.bfloat16)
    a_ws, ws = arena(B * heads * T, torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    base = qkv.data_ptr()
    assert o.lib.tvae_attn_set_tcgen05(-1) == 1
    rc = o.lib.tvae_attn_fwd_tc(base, base + 4 * C, base + 8 * C, 3 * C, B, T, C, heads, ob.data_ptr(), of.data_ptr(),
                                lse.data_ptr(), st)
    assert rc == 0, o.lib.tvae_last_error()
    rc = o.lib.tvae_attn_bwd_tc(base, base + 4 * C, base + 8 * C, 3 * C, of.data_ptr(), d_out.data_ptr(), lse.data_ptr(), B, T,
                                C, heads, dqkv.data_ptr(), ws.data_ptr(), st)
    assert rc == 0, o.lib.tvae_last_error()
    torch.cuda.synchronize()
    for a, payload in ((a_ob, ob), (a_of, of), (a_lse, lse), (a_dq, dqkv), (a_ws, ws)):
        assert torch.isnan(a[:band].float()).all() and torch.isnan(a[-band:].float()).all()
        assert torch.isfinite(payload.float()).all()


def test_reparam_kl_fwd_bwd():
    o = ops()
    B, h, w, Z = 3, 16, 16, 32
    g = torch.Generator(device="cuda").manual_seed(9)
    mom = (torch.randn((B, 2 * Z, h, w), device="cuda", generator=g) * 3).requires_grad_(True)
    with torch.no_grad():
        mom[0, Z, 0, 0] = 25.0  # exercises the clamp
        mom[0, Z + 1, 0, 0] = -40.0
    eps = torch.randn((B, Z, h, w), device="cuda", generator=g)
    mean, logvar = torch.chunk(mom, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    z = mean + torch.exp(0.5 * logvar) * eps
    kl = 0.5 * torch.sum(mean ** 2 + torch.exp(logvar) - 1.0 - logvar, dim=[1, 2, 3])
    dz = torch.randn((B, Z, h, w), device="cuda", generator=g)
    kl_scale = 0.37
    (torch.sum(z * dz) + kl_scale * kl.sum()).backward()
    mn = mom.detach().permute(0, 2, 3, 1).contiguous()
    zb, zn, eps_used, klo = o.reparam_fwd(mn, Z, eps=eps, want_z_nchw=True)
    torch.cuda.synchronize()
    assert rel_err(zn, z.detach()) < 1e-5
    assert rel_err(klo, kl.detach()) < 1e-5
    dm = o.reparam_bwd(mn, Z, dz.permute(0, 2, 3, 1).contiguous(), eps, None, None, kl_scale)
    torch.cuda.synchronize()
    assert rel_err(dm.float().permute(0, 3, 1, 2), mom.grad) < 1e-2
    # Philox mode: N(0,1) moments, reproducible, offset-keyed
    _, _, e1, _ = o.reparam_fwd(mn, Z, seed=123, sample_offset=0)
    _, _, e2, _ = o.reparam_fwd(mn, Z, seed=123, sample_offset=0)
    _, _, e3, _ = o.reparam_fwd(mn[1:], Z, seed=123, sample_offset=1)
    torch.cuda.synchronize()
    assert torch.equal(e1, e2) and torch.equal(e1[1:], e3)
    assert abs(e1.mean().item()) < 0.02 and abs(e1.std().item() - 1.0) < 0.02


def test_few_large_samples_take_the_wide_reparam_and_statistics_kernels():
    """Whole-granule inference is ONE sample with 524 K latent elements and 2,048 tile slots per GroupNorm group. Both
    per-sample kernels have a variant for that shape (reparam: a cluster of 8 blocks per sample, KL partials combined
    through distributed shared memory in rank order; statistics finalize: a block per (sample, group)); they must
    agree with the per-sample kernels on the same data -- draws and samples bit for bit -- and with torch."""
    o = ops()
    from tempo_vae_b200._lib import lib
    g = torch.Generator(device="cuda").manual_seed(21)
    Z, h, w = 32, 64, 32                                        # 65,536 latent elements per sample
    many = torch.randn((65, h, w, 2 * Z), device="cuda", generator=g) * 2       # 65 samples: the per-sample kernel
    few = many[:3].contiguous()                                                  # <= 64 samples: the clustered one
    eps = torch.randn((65, Z, h, w), device="cuda", generator=g)
    _, z_many, _, kl_many = o.reparam_fwd(many, Z, eps=eps, want_z_nchw=True)
    zb_few, z_few, _, kl_few = o.reparam_fwd(few, Z, eps=eps[:3].contiguous(), want_z_nchw=True)
    mean, logvar = few[..., :Z].permute(0, 3, 1, 2), few[..., Z:].permute(0, 3, 1, 2).clamp(-30.0, 20.0)
    assert torch.equal(z_few, z_many[:3]) and rel_err(z_few, mean + torch.exp(0.5 * logvar) * eps[:3]) < 1e-6
    assert rel_err(zb_few.float().permute(0, 3, 1, 2), z_few) < 4e-3
    kl_ref = 0.5 * torch.sum(mean.double() ** 2 + torch.exp(logvar.double()) - 1.0 - logvar.double(), dim=[1, 2, 3])
    assert rel_err(kl_few.double(), kl_ref) < 1e-6 and rel_err(kl_few, kl_many[:3]) < 1e-6
    _, _, e_many, _ = o.reparam_fwd(many, Z, seed=77, sample_offset=5)
    _, _, e_few, _ = o.reparam_fwd(few, Z, seed=77, sample_offset=5)
    assert torch.equal(e_few, e_many[:3]) and abs(e_few.mean().item()) < 0.01 and abs(e_few.std().item() - 1.0) < 0.01
    # statistics finalize: 300 tile slots per sample; 2 samples (wide kernel) against the same rows inside 129 samples
    spi, G, count, eps_gn = 300, 8, 300.0 * 128 * 16, 1e-6
    part = torch.randn((129, spi, G, 2), device="cuda", generator=g)
    part[..., 1] = part[..., 1].abs() * 40 + 30                  # sums of squares: positive, variance > 0
    st_many = torch.empty((129, G, 2), device="cuda")
    st_few = torch.empty((2, G, 2), device="cuda")
    assert lib.tvae_gn_stats_finalize(part.data_ptr(), spi, 129, G, count, eps_gn, st_many.data_ptr(), None) == 0
    assert lib.tvae_gn_stats_finalize(part.data_ptr(), spi, 2, G, count, eps_gn, st_few.data_ptr(), None) == 0
    s = part[:2].double().sum(dim=1)
    m = s[..., 0] / count
    want = torch.stack([m, 1.0 / torch.sqrt(s[..., 1] / count - m * m + eps_gn)], dim=-1)
    assert rel_err(st_few.double(), want) < 1e-6 and rel_err(st_few, st_many[:2]) < 1e-6


@pytest.mark.parametrize("HW,C", [(16, 1028), (32, 1028), (32, 20), (40, 7)])
@pytest.mark.parametrize("loss_type", [0, 1])
def test_nll(loss_type, HW, C):
    """Small pixel counts / odd channel counts take the generic kernels, [2,32,32,1028] the row-structured kernel that
    also produces the column sums of the gradient (bias gradient of the last decoder conv)."""
    o = ops()
    N, H, W = 2, HW, HW
    pitch = o.round_up(C, 8)
    g = torch.Generator(device="cuda").manual_seed(10)
    x = bf16_round(torch.randn((N, C, H, W), device="cuda", generator=g))
    xh = torch.randn((N, C, H, W), device="cuda", generator=g).requires_grad_(True)
    logvar = torch.tensor(0.7, device="cuda", requires_grad=True)
    rec = (x - xh).abs() if loss_type == 0 else (x - xh) ** 2
    nll = torch.sum(rec / torch.exp(logvar) + logvar) / N
    nll.backward()
    xb = nhwc_bf16(x, pitch)
    xhn = xh.detach().permute(0, 2, 3, 1).contiguous()
    sums, dx = o.nll_fwd(xb, xhn, C, loss_type, logvar.detach(), N, True)
    torch.cuda.synchronize()
    assert abs(sums[0].item() - rec.sum().item()) / rec.sum().item() < 1e-6
    assert abs(sums[1].item() - ((x - xh) ** 2).sum().item()) / ((x - xh) ** 2).sum().item() < 1e-6
    assert rel_err(dx[..., :C].float().permute(0, 3, 1, 2), xh.grad) < 1e-2
    assert dx.shape[-1] == pitch
    if C % 4 == 0 and N * H * W >= 148 * 8:          # row-structured kernel: pad lanes are zeroed as well
        assert (dx[..., C:] == 0).all()
    ref_cs = xh.grad.double().sum(dim=(0, 2, 3))
    scale = xh.grad.double().abs().sum(dim=(0, 2, 3)).max()
    assert (dx.tvae_colsum.double() - ref_cs).abs().max() <= 4e-3 * scale
    sums2, none = o.nll_fwd(xb, xhn, C, loss_type, logvar.detach(), N, False)       # forward only (validation)
    torch.cuda.synchronize()
    assert none is None and torch.equal(sums2[:2], sums[:2])


def test_l2head_loss():
    o = ops()
    B, h, w = 3, 16, 16
    g = torch.Generator(device="cuda").manual_seed(11)
    pred = torch.randn((B, h, w, 4), device="cuda", generator=g).requires_grad_(True)
    targets = []
    for p in range(4):
        t = torch.randn((B, 4 * h, 4 * w), device="cuda", generator=g)
        t[torch.rand((B, 4 * h, 4 * w), device="cuda", generator=g) < 0.02] = float("nan")
        targets.append(t)
    targets[3][:] = float("nan")  # a product with no valid pixel is skipped
    weights = torch.tensor([0.1, 0.2, 0.3, 0.4], device="cuda")
    total = 0.0
    ref_losses = []
    for p in range(4):
        td = F.avg_pool2d(targets[p].unsqueeze(1), 4)
        pr = pred[..., p].unsqueeze(1)
        m = ~torch.isnan(td)
        if m.sum() > 0:
            l = F.mse_loss(pr[m], td[m])
            total = total + weights[p] * l
            ref_losses.append(l.item())
        else:
            ref_losses.append(None)
    total.backward()
    sums = o.l2head_loss_fwd(pred.detach(), targets, B, h, w)
    dp = o.l2head_loss_bwd(pred.detach(), targets, B, h, w, sums, weights, 1.0)
    torch.cuda.synchronize()
    for p in range(4):
        if ref_losses[p] is None:
            assert sums[p, 1].item() == 0
        else:
            assert abs(sums[p, 0].item() / sums[p, 1].item() - ref_losses[p]) < 1e-5 * max(1.0, ref_losses[p])
    assert rel_err(dp[..., :4].float(), pred.grad) < 1e-2
    assert (dp[..., 4:] == 0).all()


def test_adamw_and_clip():
    o = ops()
    n = 100003
    g = torch.Generator(device="cuda").manual_seed(12)
    p0 = torch.randn((n,), device="cuda", generator=g)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p_ref], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    p = p0.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    ss = torch.empty((1,), dtype=torch.float64, device="cuda")
    for step in range(1, 4):
        grad = torch.randn((n,), device="cuda", generator=g) * (10.0 if step == 1 else 1e-3)
        p_ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        o.sumsq(grad, ss)
        o.adamw(p, grad, m, v, lr=1e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.05, step=step, sumsq_buf=ss,
                max_norm=1.0)
        torch.cuda.synchronize()
        assert abs(ss.item() - grad.double().pow(2).sum().item()) / ss.item() < 1e-7
        assert torch.allclose(p, p_ref.detach(), rtol=1e-5, atol=1e-7)


@pytest.mark.usefixtures("conv_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout,kind", [(2, 64, 64, 128, 512, 0), (3, 16, 16, 64, 128, 0), (2, 32, 32, 256, 256, 1),
                                                 (2, 16, 16, 128, 256, 2), (2, 8, 8, 64, 128, 0)])
def test_conv_fused_groupnorm_stats(N, H, W, Cin, Cout, kind):
    """The conv epilogue's (mean, rstd) of its own output == a separate statistics pass over that output."""
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(13)
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    b = torch.randn((Cout,), device="cuda", generator=g)
    G, eps = 8, 1e-6
    if kind == 2:
        w = bf16_round(torch.randn((Cin, Cout, 2, 2), device="cuda", generator=g) / math.sqrt(Cin))
        wp, R, oH, oW = o.pack_weight(w, "up_fwd"), 2, 2 * H, 2 * W
    elif kind == 1:
        w = bf16_round(torch.randn((Cout, Cin, 2, 2), device="cuda", generator=g) / math.sqrt(4 * Cin))
        wp, R, oH, oW = o.pack_weight(w, "fwd"), 2, H // 2, W // 2
    else:
        w = bf16_round(torch.randn((Cout, Cin, 3, 3), device="cuda", generator=g) / math.sqrt(9 * Cin))
        wp, R, oH, oW = o.pack_weight(w, "fwd"), 3, H, W
    res = torch.randn((N, oH, oW, Cout), device="cuda", generator=g)
    of, _, st = o.conv_gemm(nhwc_bf16(x, o.round_up(Cin, 8)), Cin, wp, kind=kind, R=R, Cout=Cout, bias=b, residual=res,
                            stats=(G, eps))
    torch.cuda.synchronize()
    if oH * oW < 128 and kind != 2 or (kind == 2 and H * W < 128):
        assert st is None          # images smaller than one 128-pixel tile: the caller falls back to gn_stats
        return
    assert st is not None
    ref = o.gn_stats(of, Cout, G, eps)
    torch.cuda.synchronize()
    assert torch.allclose(st[..., 0], ref[..., 0], atol=1e-5, rtol=1e-5)
    assert torch.allclose(st[..., 1], ref[..., 1], rtol=1e-5)


@pytest.mark.usefixtures("conv_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout,R", [(2, 16, 16, 128, 128, 3), (1, 64, 64, 1028, 512, 3), (2, 16, 16, 64, 64, 1)])
def test_conv_split_bf16_fp32_mode(N, H, W, Cin, Cout, R):
    """Split-bf16 operands (hi + lo): the conv matches an fp32 convolution of the UNROUNDED inputs to ~2^-16."""
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(14)
    x = torch.randn((N, Cin, H, W), device="cuda", generator=g)
    w = torch.randn((Cout, Cin, R, R), device="cuda", generator=g) / math.sqrt(Cin * R * R)
    b = torch.randn((Cout,), device="cuda", generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=R // 2).float()
    o.SPLIT_BF16[0] = True
    try:
        xp = o.nchw_to_nhwc_bf16(x, o.round_up(Cin, 8))
        assert isinstance(xp, o.Pair)
        rec = (xp.hi.float() + xp.lo.float())[..., :Cin].permute(0, 3, 1, 2)
        assert rel_err(rec, x) < 2e-5
        wp = o.pack_weight(w, "fwd")
        of, ob = o.conv_gemm(xp, Cin, wp, kind=0, R=R, Cout=Cout, bias=b, want_bf16=True)
        torch.cuda.synchronize()
        assert rel_err(of[..., :Cout].permute(0, 3, 1, 2), ref) < 5e-5
        assert isinstance(ob, o.Pair)
        assert rel_err((ob.hi.float() + ob.lo.float())[..., :Cout].permute(0, 3, 1, 2), ref) < 5e-5
    finally:
        o.SPLIT_BF16[0] = False
    # plain bf16 operands on the same unrounded inputs are ~100x less accurate: the mode really is in effect
    of2, _ = o.conv_gemm(o.nchw_to_nhwc_bf16(x, o.round_up(Cin, 8)), Cin, o.pack_weight(w, "fwd"), kind=0, R=R, Cout=Cout,
                         bias=b)
    torch.cuda.synchronize()
    assert rel_err(of2[..., :Cout].permute(0, 3, 1, 2), ref) > 5e-4


@pytest.mark.usefixtures("conv_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout,kind", [(3, 8, 8, 64, 1028, 0), (5, 4, 4, 32, 48, 0), (1, 64, 64, 512, 1028, 0),
                                                 (3, 8, 8, 32, 16, 2), (3, 8, 8, 32, 48, 1)])
def test_conv_epilogue_writes_stay_in_bounds(N, H, W, Cin, Cout, kind):
    """compute-sanitizer is closed on this pool, so the masked epilogue (partial N tiles, rows past the last pixel,
    pixel-shuffle scatter) is checked with guard bands: outputs are views into the middle of sentinel-filled
    buffers and the sentinels must survive."""
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(15)
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    if kind == 2:
        w = bf16_round(torch.randn((Cin, Cout, 2, 2), device="cuda", generator=g))
        wp, R, oH, oW = o.pack_weight(w, "up_fwd"), 2, 2 * H, 2 * W
    elif kind == 1:
        w = bf16_round(torch.randn((Cout, Cin, 2, 2), device="cuda", generator=g))
        wp, R, oH, oW = o.pack_weight(w, "fwd"), 2, H // 2, W // 2
    else:
        w = bf16_round(torch.randn((Cout, Cin, 3, 3), device="cuda", generator=g))
        wp, R, oH, oW = o.pack_weight(w, "fwd"), 3, H, W
    pf, pb = o.round_up(Cout, 4), o.round_up(Cout, 8)
    guard = 4096
    n32, n16 = N * oH * oW * pf, N * oH * oW * pb
    buf32 = torch.full((n32 + 2 * guard,), 12345.0, device="cuda")
    buf16 = torch.full((n16 + 2 * guard,), 77.0, device="cuda", dtype=torch.bfloat16)
    of = buf32[guard:guard + n32].view(N, oH, oW, pf)
    ob = buf16[guard:guard + n16].view(N, oH, oW, pb)
    o.conv_gemm(nhwc_bf16(x, o.round_up(Cin, 8)), Cin, wp, kind=kind, R=R, Cout=Cout, want_f32=True, want_bf16=True,
                out_f32=of, out_bf16=ob)
    torch.cuda.synchronize()
    assert (buf32[:guard] == 12345.0).all() and (buf32[guard + n32:] == 12345.0).all()
    assert (buf16[:guard] == 77.0).all() and (buf16[guard + n16:] == 77.0).all()
    assert (of[..., Cout:] == 12345.0).all() and (ob[..., Cout:] == 77.0).all()      # pad lanes are never written
    assert torch.isfinite(of[..., :Cout]).all() and (of[..., :Cout] != 12345.0).any()


@pytest.mark.usefixtures("wgrad_sched")
def test_wgrad_writes_stay_in_bounds():
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(16)
    N, H, W, Cin, Cout = 3, 8, 8, 20, 36
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    dy = bf16_round(torch.randn((N, Cout, H, W), device="cuda", generator=g))
    n = Cout * Cin * 9
    buf = torch.full((n + 2048,), 555.0, device="cuda")
    grad = buf[1024:1024 + n].view(Cout, Cin, 3, 3)
    o.wgrad_gemm(nhwc_bf16(dy, o.round_up(Cout, 8)), Cout, nhwc_bf16(x, o.round_up(Cin, 8)), Cin, kind=0, R=3, grad=grad)
    torch.cuda.synchronize()
    assert (buf[:1024] == 555.0).all() and (buf[1024 + n:] == 555.0).all()
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, 3, 3), dy, padding=1)
    assert rel_err(grad, ref) < 2e-3


@pytest.mark.usefixtures("wgrad_sched")
@pytest.mark.parametrize("N,H,W,Cin,Cout,R", [(2, 16, 16, 300, 128, 3), (1, 64, 64, 1028, 512, 3), (2, 16, 16, 260, 128, 1)])
def test_wgrad_exchanged_operand_roles(N, H, W, Cin, Cout, R):
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(17)
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    dy = bf16_round(torch.randn((N, Cout, H, W), device="cuda", generator=g))
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, R, R), dy, padding=R // 2)
    grad = torch.full((Cout, Cin, R, R), float("nan"), device="cuda")
    o.wgrad_gemm(nhwc_bf16(x, o.round_up(Cin, 8)), Cin, nhwc_bf16(dy, o.round_up(Cout, 8)), Cout, kind=0, R=R, grad=grad,
                 flip=True)
    torch.cuda.synchronize()
    assert rel_err(grad, ref) < 2e-3


@pytest.mark.parametrize("N,H,W,Cin,Cout,R", [(3, 32, 32, 256, 256, 3), (5, 8, 8, 64, 48, 3), (1, 64, 64, 512, 1028, 3)])
def test_conv_cta_pair_is_bit_identical_to_single_cta(N, H, W, Cin, Cout, R):
    """Both schedules accumulate the same K blocks in the same order into fp32 TMEM: outputs must match bit for bit
    (includes an odd number of M tiles, where the peer CTA of the last pair works on zero-filled rows)."""
    from tempo_vae_b200._lib import lib
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(11)
    x = bf16_round(torch.randn((N, Cin, H, W), device="cuda", generator=g))
    w = bf16_round(torch.randn((Cout, Cin, R, R), device="cuda", generator=g) / math.sqrt(Cin * R * R))
    b = torch.randn((Cout,), device="cuda", generator=g)
    xp = nhwc_bf16(x, o.round_up(Cin, 8))
    wp = o.pack_weight(w, "fwd")
    outs = []
    prev = lib.tvae_conv_set_cta_pair(1)
    try:
        for mode in (1, 0):
            lib.tvae_conv_set_cta_pair(mode)
            of, ob = o.conv_gemm(xp, Cin, wp, kind=0, R=R, Cout=Cout, bias=b, want_f32=True, want_bf16=True)
            torch.cuda.synchronize()
            outs.append((of[..., :Cout].clone(), ob[..., :Cout].clone()))
    finally:
        lib.tvae_conv_set_cta_pair(prev)
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("Cin,Cout,splits", [(256, 256, 2), (128, 512, 3)])
def test_wgrad_cta_pair_is_bit_identical_to_single_cta(Cin, Cout, splits):
    """Even M tile count and the same split-K factor: both schedules add the same K blocks in the same order."""
    from tempo_vae_b200._lib import lib
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(13)
    x = nhwc_bf16(bf16_round(torch.randn((3, Cin, 32, 32), device="cuda", generator=g)), Cin)
    dy = nhwc_bf16(bf16_round(torch.randn((3, Cout, 32, 32), device="cuda", generator=g)), Cout)
    outs = []
    prev = lib.tvae_wgrad_set_cta_pair(1)
    try:
        for mode in (1, 0):
            lib.tvae_wgrad_set_cta_pair(mode)
            grad = torch.full((Cout, Cin, 3, 3), float("nan"), device="cuda")
            o.wgrad_gemm(dy, Cout, x, Cin, kind=0, R=3, grad=grad, splits=splits)
            torch.cuda.synchronize()
            outs.append(grad)
    finally:
        lib.tvae_wgrad_set_cta_pair(prev)
    assert torch.equal(outs[0], outs[1])


def test_batched_weight_packing_matches_single_packs():
    """tvae_pack_weights_batched (one launch for all packs of a step) against tvae_pack_weight, every pack mode."""
    o = ops()
    g = torch.Generator(device="cuda").manual_seed(29)
    specs = [((512, 1028, 3, 3), "fwd"), ((512, 1028, 3, 3), "dgrad"), ((64, 128, 1, 1), "fwd"), ((48, 20, 3, 3), "dgrad"),
             ((256, 256, 2, 2), "fwd"), ((256, 256, 2, 2), "down_dgrad"), ((128, 256, 2, 2), "up_fwd"),
             ((128, 256, 2, 2), "up_dgrad"), ((4, 512, 1, 1), "fwd")]
    items, refs = [], []
    for shape, mode in specs:
        w = torch.randn(shape, device="cuda", generator=g)
        ref = o.pack_weight(w, mode)
        ent = o.pack_weight(torch.zeros_like(w), mode)        # same geometry, different content
        items.append((w, mode, ent))
        refs.append(ref)
    o.pack_weights_batched(items)
    o.pack_weights_batched(items)                              # second call reuses the cached descriptor table
    torch.cuda.synchronize()
    for (w, mode, ent), ref in zip(items, refs):
        assert torch.equal(ent.data, ref.data), (tuple(w.shape), mode)
