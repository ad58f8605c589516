"""GPU parity: FeatUp JBU stack pieces and the AdaptiveConv kernels vs the oracle
(oracle/jbu.py; parity UNPINNED upstream, see its header).  fp32 mode: <= 1e-3 max
relative error (north_star)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import jbu as ojbu
from oracle import synth
from tests.gpu_util import DEV, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def isp():
    import isegprobe_b200
    return isegprobe_b200


def _call(name, *a):
    from isegprobe_b200 import _lib
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 32, 64), (1, 19, 37, 128), (2, 64, 64, 384), (1, 8, 16, 512)])
def test_adaptive_conv_nhwc(isp, B, H, W, C):
    g = torch.Generator().manual_seed(B * H + C)
    x = torch.randn(B, C, H + 6, W + 6, generator=g)
    f = torch.randn(B, H, W, 7, 7, generator=g) * 0.2
    want = ojbu.adaptive_conv(x, f)
    xin = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.empty(B, H, W, C, device=DEV)
    _call("isp_adaptive_conv_fwd", xin, f.reshape(B, H, W, 49).contiguous().to(DEV), out, B, H, W, C, 49)
    f56 = torch.zeros(B, H, W, 7, 8)
    f56[..., :7] = f
    out56 = torch.empty(B, H, W, C, device=DEV)
    _call("isp_adaptive_conv_fwd", xin, f56.reshape(B, H, W, 56).contiguous().to(DEV), out56, B, H, W, C, 56)
    assert torch.equal(out56, out)  # padded-filter (TMA) path computes the same thing
    assert relerr(out.permute(0, 3, 1, 2), want) < 1e-5


def test_adaptive_conv_properties_full_size(isp):
    """Size-independent properties at BASELINE size (512x512x384): a centre-delta
    filter is the identity; linearity in the input."""
    B, H, W, C = 1, 512, 512, 384
    x = torch.randn(B, H + 6, W + 6, C, device=DEV)
    f = torch.zeros(B, H, W, 49, device=DEV)
    f[..., 24] = 1
    out = torch.empty(B, H, W, C, device=DEV)
    _call("isp_adaptive_conv_fwd", x, f, out, B, H, W, C, 49)
    assert torch.equal(out, x[:, 3:-3, 3:-3])
    f = torch.rand(B, H, W, 49, device=DEV)
    y = torch.randn_like(x)
    o1, o2, o3 = (torch.empty(B, H, W, C, device=DEV) for _ in range(3))
    _call("isp_adaptive_conv_fwd", x, f, o1, B, H, W, C, 49)
    _call("isp_adaptive_conv_fwd", y, f, o2, B, H, W, C, 49)
    _call("isp_adaptive_conv_fwd", x + 2 * y, f, o3, B, H, W, C, 49)
    assert relerr(o3, o1 + 2 * o2) < 1e-5


def test_adaptive_conv_grad_input(isp):
    B, H, W, C = 2, 11, 13, 64
    g = torch.Generator().manual_seed(7)
    go = torch.randn(B, C, H, W, generator=g)
    f = torch.randn(B, H, W, 7, 7, generator=g)
    want = ojbu.adaptive_conv_grad_input(go, f)
    gi = torch.empty(B, H + 6, W + 6, C, device=DEV)
    _call("isp_adaptive_conv_grad_input", go.permute(0, 2, 3, 1).contiguous().to(DEV),
          f.reshape(B, H, W, 49).contiguous().to(DEV), gi, B, H, W, C)
    assert relerr(gi.permute(0, 3, 1, 2), want) < 1e-5


@pytest.mark.parametrize("H,W,OH,OW", [(448, 448, 64, 64), (448, 448, 128, 128), (448, 448, 512, 512), (50, 70, 16, 24)])
def test_pool_guidance(isp, H, W, OH, OW):
    gd = synth.image_batch(2, H, W, seed=1)
    out = torch.empty(2, OH, OW, 4, device=DEV)
    d = gd.to(DEV)
    _call("isp_jbu_pool_guidance", d, out, 2, H, W, OH, OW, *d.stride())
    want = F.adaptive_avg_pool2d(gd, (OH, OW)).permute(0, 2, 3, 1)
    assert relerr(out[..., :3], want) < 1e-6 and float(out[..., 3].abs().max()) == 0


def test_bicubic_reflectpad(isp):
    src = synth.lr_features(2, 64, 9, 12, seed=2)
    out = torch.empty(2, 24, 30, 64, device=DEV)
    _call("isp_jbu_bicubic2x_reflectpad", src.permute(0, 2, 3, 1).contiguous().to(DEV), out, 2, 9, 12, 64)
    want = F.pad(F.interpolate(src, size=(18, 24), mode="bicubic", align_corners=False), [3] * 4, mode="reflect")
    assert relerr(out.permute(0, 3, 1, 2), want) < 1e-5


def test_filters_and_range_proj(isp):
    sd = ojbu.init_state_dict(8, seed=3)
    sd["up2.range_temp"] = torch.tensor(0.7)
    sd["up2.sigma_spatial"] = torch.tensor(0.8)
    g = F.adaptive_avg_pool2d((synth.image_batch(2, 60, 44, seed=1) - 0.45) / 0.225, (30, 22))
    want = ojbu.combined_kernel(sd, "up2", g).reshape(2, 30, 22, 49)
    g4 = torch.cat([g, torch.zeros(2, 1, 30, 22)], 1).permute(0, 2, 3, 1).contiguous().to(DEV)
    proj = torch.empty(2, 30, 22, 32, device=DEV)
    p = "up2"
    w = {k: sd[f"{p}.{k}"].reshape(sd[f"{p}.{k}"].shape[0], -1).contiguous().to(DEV) if sd[f"{p}.{k}"].dim() > 1
         else sd[f"{p}.{k}"].to(DEV) for k in ("range_proj.0.weight", "range_proj.0.bias", "range_proj.3.weight",
                                                "range_proj.3.bias", "fixup_proj.0.weight", "fixup_proj.0.bias",
                                                "fixup_proj.3.weight", "fixup_proj.3.bias")}
    _call("isp_jbu_range_proj", g4, proj, 2 * 30 * 22, w["range_proj.0.weight"], w["range_proj.0.bias"],
          w["range_proj.3.weight"], w["range_proj.3.bias"])
    pw = F.conv2d(F.gelu(F.conv2d(g, sd[p + ".range_proj.0.weight"], sd[p + ".range_proj.0.bias"])),
                  sd[p + ".range_proj.3.weight"], sd[p + ".range_proj.3.bias"])
    assert relerr(proj.permute(0, 3, 1, 2), pw) < 1e-5
    filt = torch.empty(2, 30, 22, 49, device=DEV)
    import math
    _call("isp_jbu_filters", proj, g4, filt, 2, 30, 22, math.exp(0.7), 0.8, w["fixup_proj.0.weight"],
          w["fixup_proj.0.bias"], w["fixup_proj.3.weight"], w["fixup_proj.3.bias"], 49)
    assert relerr(filt, want) < 1e-4
    filt56 = torch.empty(2, 30, 22, 7, 8, device=DEV)
    _call("isp_jbu_filters_simt", proj, g4, filt56, 2, 30, 22, math.exp(0.7), 0.8, w["fixup_proj.0.weight"],
          w["fixup_proj.0.bias"], w["fixup_proj.3.weight"], w["fixup_proj.3.bias"], 56)
    assert torch.equal(filt56[..., :7].reshape(2, 30, 22, 49), filt) and float(filt56[..., 7].abs().max()) == 0
    # padded layout, fix-up MLP on the tensor cores (kind::tf32, hi / lo split activations, weights rounded to tf32): 1320 pixels
    # = ten 128-pixel tiles and a ragged one; the pad slots carry the guidance values inside the kernel and must come out zero
    tc56 = torch.full((2, 30, 22, 7, 8), float("nan"), device=DEV)
    _call("isp_jbu_filters", proj, g4, tc56, 2, 30, 22, math.exp(0.7), 0.8, w["fixup_proj.0.weight"],
          w["fixup_proj.0.bias"], w["fixup_proj.3.weight"], w["fixup_proj.3.bias"], 56)
    assert float(tc56[..., 7].abs().max()) == 0
    assert relerr(tc56[..., :7].reshape(2, 30, 22, 49), want) < 1e-4
    assert relerr(tc56[..., :7].reshape(2, 30, 22, 49), filt) < 3e-5, relerr(tc56[..., :7].reshape(2, 30, 22, 49), filt)


@pytest.mark.parametrize("B,h,w,H,W", [(2, 4, 4, 64, 64), (1, 6, 9, 96, 144), (1, 8, 8, 112, 112)])
def test_jbu_stack_module(isp, B, h, w, H, W):
    sd = ojbu.init_state_dict(384, seed=0)
    up = isp.JBUFeatUpUpsampler("dinov2").to(DEV).eval()
    up.upsampler.load_state_dict(sd)
    src = synth.lr_features(B, 384, h, w, seed=2)
    gd = (synth.image_batch(B, H, W, seed=1) - 0.45) / 0.225
    with torch.no_grad():
        out = up(source=src.to(DEV), guidance=gd.to(DEV))
        want = ojbu.jbu_stack_forward(sd, src, gd)
    assert tuple(out.shape) == (B, 384, 16 * h, 16 * w)
    assert relerr(out, want) < TOL


def test_jbu_forward_resized(isp):
    """What the pipeline runs (channel map on the source, stack, resize + bias) equals the reference order, stack with its
    final 1x1 then bilinear align_corners resize (iseg_probe_model.py:120-129), to fp32-mode tolerance."""
    sd = ojbu.init_state_dict(384, seed=0)
    up = isp.JBUFeatUpUpsampler("dinov2").to(DEV).eval()
    up.upsampler.load_state_dict(sd)
    src = synth.lr_features(2, 384, 8, 8, seed=2)
    gd = (synth.image_batch(2, 112, 112, seed=1) - 0.45) / 0.225
    with torch.no_grad():
        out = up.forward_resized(src.to(DEV), gd.to(DEV), (112, 112))
        want = F.interpolate(ojbu.jbu_stack_forward(sd, src, gd), size=(112, 112), mode="bilinear", align_corners=True)
    assert tuple(out.shape) == (2, 384, 112, 112)
    assert relerr(out, want) < TOL


@pytest.mark.parametrize("B,C,H,W,OH,OW,with_bias", [(2, 128, 40, 56, 35, 49, True), (1, 384, 64, 64, 56, 56, True),
                                                       (2, 256, 20, 24, 33, 37, False), (1, 128, 16, 16, 16, 16, True),
                                                       (1, 96, 12, 12, 9, 30, True)])
def test_bilinear_resize_with_bias(isp, B, C, H, W, OH, OW, with_bias):
    """isp_bilinear_ac_nhwc_bias (the JBU stack's last pass: strip-marching kernel for C % 128 == 0, the per-pixel kernel
    otherwise) against ATen's align_corners=True resize plus a per-channel constant."""
    g = torch.Generator().manual_seed(C + OH)
    x, bias = torch.randn(B, C, H, W, generator=g), torch.randn(C, generator=g)
    want = F.interpolate(x, size=(OH, OW), mode="bilinear", align_corners=True) + (bias.view(1, C, 1, 1) if with_bias else 0.0)
    out = torch.empty(B, OH, OW, C, device=DEV)
    xin = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    _call("isp_bilinear_ac_nhwc_bias", xin, out, 0, bias.to(DEV) if with_bias else None, B, C, H, W, OH, OW)
    assert relerr(out.permute(0, 3, 1, 2), want) < 1e-5
    out_bf = torch.empty(B, OH, OW, C, device=DEV, dtype=torch.bfloat16)  # the head's input format: the same values, rounded
    _call("isp_bilinear_ac_nhwc_bias", xin, out_bf, 1, bias.to(DEV) if with_bias else None, B, C, H, W, OH, OW)
    assert torch.equal(out_bf, out.to(torch.bfloat16))


def test_jbu_channel_map_commutes(isp):
    """The stack's final `fixup_proj(x) * 0.1 + x` applied to the SOURCE (upsamplers._mix_channels) instead of the 16x
    output: forward and adjoint against plain fp32 matrix arithmetic (split-bf16 products: ~1e-5 of the 0.1-scaled term)."""
    up = isp.JBUFeatUpUpsampler("dinov2").to(DEV).eval()
    up.upsampler.load_state_dict(ojbu.init_state_dict(384, seed=0))
    Wm = up.upsampler.fixup_proj[1].weight.detach().view(384, 384).double()
    x = torch.randn(2, 5, 7, 384, generator=torch.Generator().manual_seed(4)).to(DEV)
    m = (torch.rand(2, 384, generator=torch.Generator().manual_seed(5)) > 0.2).float().to(DEV) / 0.8
    for mask in (None, m):
        xm = x.double() if mask is None else x.double() * mask[:, None, None, :].double()
        assert relerr(up._mix_channels(x, False, mask), x.double() + 0.1 * xm @ Wm.T) < 1e-5
        gW = x.double() @ Wm
        want_t = x.double() + 0.1 * (gW if mask is None else gW * mask[:, None, None, :].double())
        assert relerr(up._mix_channels(x, True, mask), want_t) < 1e-5


def test_jbu_reference_shape_contract(isp):
    """JBUFeatUp.py:36-45: [1,384,14,14] + [1,3,224,224] -> [1,384,224,224]."""
    up = isp.JBUFeatUpUpsampler(backbone_type="dinov2").to(DEV).eval()
    out = up(torch.rand(1, 384, 14, 14, device=DEV), torch.rand(1, 3, 224, 224, device=DEV))
    assert tuple(out.shape) == (1, 384, 224, 224) and bool(torch.isfinite(out).all())
    with pytest.raises(AssertionError):
        isp.JBUFeatUpUpsampler(backbone_type="nope")


def test_layout_bilinear_gemm(isp):
    x = synth.lr_features(2, 96, 10, 14, seed=5)
    xin = x.to(DEV)
    nhwc = torch.empty(2, 10, 14, 96, device=DEV)
    _call("isp_nchw_to_nhwc_f32", xin, nhwc, 2, 96, 10, 14, *xin.stride())
    assert torch.equal(nhwc.cpu(), x.permute(0, 2, 3, 1))
    from isegprobe_b200.upsamplers import bilinear_align_corners_nhwc
    out = bilinear_align_corners_nhwc(nhwc, (23, 31))
    want = F.interpolate(x, size=(23, 31), mode="bilinear", align_corners=True)
    assert relerr(out.permute(0, 3, 1, 2), want) < 1e-5
    g = torch.Generator().manual_seed(1)
    A, Wt, b = torch.randn(300, 70, generator=g), torch.randn(50, 70, generator=g), torch.randn(50, generator=g)
    R = torch.randn(300, 50, generator=g)
    C = torch.empty(300, 50, device=DEV)
    _call("isp_gemm_f32_simt", A.to(DEV), Wt.to(DEV), b.to(DEV), R.to(DEV), 0.1, C, 300, 50, 70)
    assert relerr(C, 0.1 * (A @ Wt.T + b) + R) < 1e-5


def test_jbu_train_mode_dropout_fwd_bwd():
    """train() (the reference's trainer puts the frozen stack in train(), core/training/trainer.py:213-214): FeatUp's three
    Dropout2d sites with explicit masks, forward and the input gradient vs torch autograd through the oracle with the same
    masks; and fresh draws differ from eval()."""
    import isegprobe_b200 as isp
    from oracle import jbu as ojbu
    up = isp.JBUFeatUpUpsampler("dinov2").to(DEV)
    sd = ojbu.init_state_dict(384, seed=0)
    up.upsampler.load_state_dict(sd)
    B = 2
    src = synth.lr_features(B, 384, 4, 6, seed=2)
    gd = (synth.image_batch(B, 64, 96, seed=1) - 0.45) / 0.225
    masks = ojbu.dropout2d_masks(B, 384, seed=7)
    up.train()
    up.dropout_masks = masks
    s_d = src.to(DEV).requires_grad_(True)
    out = up(s_d, gd.to(DEV))
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    (out * gout.to(DEV)).sum().backward()
    s_o = src.clone().requires_grad_(True)
    want = ojbu.jbu_stack_forward(sd, s_o, gd, masks)
    (want * gout).sum().backward()
    assert relerr(out, want) < 1e-3, relerr(out, want)
    assert relerr(s_d.grad, s_o.grad) < 5e-3, relerr(s_d.grad, s_o.grad)
    # eval() ignores the masks; train() with fresh draws is stochastic
    up.eval()
    with torch.no_grad():
        ev = up(src.to(DEV), gd.to(DEV))
        assert relerr(ev, ojbu.jbu_stack_forward(sd, src, gd)) < 1e-3
        up.train()
        up.dropout_masks = None
        a, b = up(src.to(DEV), gd.to(DEV)), up(src.to(DEV), gd.to(DEV))
    assert relerr(a, ev) > 1e-2 and relerr(a, b) > 1e-3
