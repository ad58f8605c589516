"""GPU parity at BASELINE.json's full sizes (448x448 images, 32x32x384 low-res features, 512x512 JBU stage):
the oracle where it still finishes in seconds (one image), and size-independent properties elsewhere
(identity filters, linearity, softmax rows summing to one, uniform attention = mean of V, sampled-pixel checks)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import jbu as ojbu
from oracle import loftup as oloft
from oracle import synth
from tests.gpu_util import DEV, cosine, relerr

pytestmark = pytest.mark.gpu


def _call(name, *a):
    from isegprobe_b200 import _lib
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


def test_adaptive_conv_padded_filters_full_size():
    """The kernel the JBU stack runs (row-padded [7][8] filters, two output rows per half-warp) at 512x512x384:
    centre-delta filter = identity (bit-exact), linearity, and equality with the dense-filter kernel."""
    B, H, W, C = 1, 512, 512, 384
    x = torch.randn(B, H + 6, W + 6, C, device=DEV)
    f49 = torch.zeros(B, H, W, 49, device=DEV)
    f49[..., 24] = 1

    def pad(f):  # [.,49] -> [.,7,8]
        return F.pad(f.view(B, H, W, 7, 7), (0, 1)).reshape(B, H, W, 56).contiguous()

    out = torch.empty(B, H, W, C, device=DEV)
    _call("isp_adaptive_conv_fwd", x, pad(f49), out, B, H, W, C, 56)
    assert torch.equal(out, x[:, 3:-3, 3:-3])
    f49 = torch.rand(B, H, W, 49, device=DEV)
    y = torch.randn_like(x)
    o1, o2, o3, o4 = (torch.empty(B, H, W, C, device=DEV) for _ in range(4))
    _call("isp_adaptive_conv_fwd", x, pad(f49), o1, B, H, W, C, 56)
    _call("isp_adaptive_conv_fwd", y, pad(f49), o2, B, H, W, C, 56)
    _call("isp_adaptive_conv_fwd", x + 2 * y, pad(f49), o3, B, H, W, C, 56)
    assert relerr(o3, o1 + 2 * o2) < 1e-5
    _call("isp_adaptive_conv_fwd", x, f49, o4, B, H, W, C, 49)
    assert relerr(o1, o4) < 1e-6
    # sampled pixels against the definition (loftup-free fp64 check)
    g = torch.Generator().manual_seed(0)
    ys, xs = torch.randint(0, H, (32,), generator=g), torch.randint(0, W, (32,), generator=g)
    xc, fc, oc = x.cpu().double(), f49.cpu().double(), o1.cpu().double()
    for yy, xx in zip(ys.tolist(), xs.tolist()):
        want = (xc[0, yy:yy + 7, xx:xx + 7, :] * fc[0, yy, xx].view(7, 7, 1)).sum((0, 1))
        assert float((oc[0, yy, xx] - want).abs().max()) < 1e-4


def test_attention_properties_loftup_size():
    """200 704 pixel queries x 1024 keys, 4 heads x 101 (LoftUp at 448^2): with V = 1 every output is the row sum of
    the softmax (= 1); with Q = 0 the attention is uniform and the output is the mean of V over the keys."""
    M, T, nh, hd, HP, KP = 448 * 448, 1024, 4, 101, 112, 128
    bf = torch.bfloat16
    g = torch.Generator(device=DEV).manual_seed(1)
    Q = torch.zeros(M, nh * HP, dtype=bf, device=DEV)
    Q.view(M, nh, HP)[:, :, :hd] = (torch.randn(M, nh, hd, device=DEV, generator=g) * 0.4).to(bf)
    K = torch.zeros(1, nh, T, KP, dtype=bf, device=DEV)
    K[..., :hd] = (torch.randn(1, nh, T, hd, device=DEV, generator=g) * 0.4).to(bf)
    Vt = torch.zeros(1, nh, HP, T, dtype=bf, device=DEV)
    Vt[:, :, :hd] = 1
    O = torch.empty(M, nh * HP, dtype=bf, device=DEV)
    _call("isp_attention_bf16_tc", Q, nh * HP, HP, K, Vt, O, nh * HP, HP, 1, M, nh, T, 1)
    Or = O.view(M, nh, HP)[:, :, :hd].float()
    assert float((Or - 1).abs().max()) < 1e-2, float((Or - 1).abs().max())  # bf16 P rounding
    assert float(O.view(M, nh, HP)[:, :, hd:].float().abs().max()) == 0     # padded columns stay zero
    Vt[:, :, :hd] = torch.randn(1, nh, hd, T, device=DEV, generator=g).to(bf)
    Q.zero_()
    _call("isp_attention_bf16_tc", Q, nh * HP, HP, K, Vt, O, nh * HP, HP, 1, M, nh, T, 1)
    want = Vt[0, :, :hd].float().mean(-1)  # [nh, hd]
    got = O.view(M, nh, HP)[:, :, :hd].float()
    assert float((got - want[None]).abs().max()) < 2e-3
    # a slice of rows against fp32 softmax on the same bf16 operands
    Q.view(M, nh, HP)[:, :, :hd] = (torch.randn(M, nh, hd, device=DEV, generator=g) * 0.4).to(bf)
    _call("isp_attention_bf16_tc", Q, nh * HP, HP, K, Vt, O, nh * HP, HP, 1, M, nh, T, 1)
    rows = torch.tensor([0, 1, 127, 128, 100000, M - 129, M - 1], device=DEV)
    q = Q.view(M, nh, HP)[rows].float().permute(1, 0, 2)               # [nh, r, HP]
    s = q @ F.pad(K[0].float(), (0, 0))[:, :, :HP].transpose(1, 2)     # [nh, r, T]
    ref = (torch.softmax(s, -1) @ Vt[0].float().transpose(1, 2)).permute(1, 0, 2)[:, :, :hd]
    got = O.view(M, nh, HP)[rows][:, :, :hd].float()
    assert relerr(got, ref) < 2e-2 and cosine(got, ref) > 0.9995


def test_conv3x3_head_size_sampled_pixels():
    """ConvSegHead layer at full size (1 x 448 x 448 x 384 -> 384, the paired-tile kernel): 48 sampled output pixels
    against the fp32 definition on the same bf16 operands, plus linearity in the input."""
    from isegprobe_b200 import tc
    H = W = 448
    C = 384
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, H, W, C, generator=g).to(torch.bfloat16)
    w = (torch.randn(C, C, 3, 3, generator=g) * (9 * C) ** -0.5).to(torch.bfloat16)
    b = torch.randn(C, generator=g)
    wp = tc.pack_conv3x3_weight(w).to(DEV)
    y = tc.conv3x3(x.to(DEV), wp, b.to(DEV), C, C, act=None, out_dtype=torch.float32).cpu()
    xf = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1))
    ys, xs = torch.randint(0, H, (48,), generator=g), torch.randint(0, W, (48,), generator=g)
    ys[:4], xs[:4] = torch.tensor([0, 0, H - 1, H - 1]), torch.tensor([0, W - 1, 0, W - 1])  # corners: zero padding
    for yy, xx in zip(ys.tolist(), xs.tolist()):
        patch = xf[0, :, yy:yy + 3, xx:xx + 3]
        want = (w.float() * patch[None]).sum((1, 2, 3)) + b
        assert float((y[0, yy, xx] - want).abs().max()) < 2e-3 * float(want.abs().max())
    x2 = torch.randn(1, H, W, C, generator=g).to(torch.bfloat16)
    zero_b = torch.zeros(C, device=DEV)
    y1 = tc.conv3x3(x.to(DEV), wp, zero_b, C, C, act=None, out_dtype=torch.float32)
    y2 = tc.conv3x3(x2.to(DEV), wp, zero_b, C, C, act=None, out_dtype=torch.float32)
    xs_ = (x.float() + x2.float()).to(torch.bfloat16)
    y3 = tc.conv3x3(xs_.to(DEV), wp, zero_b, C, C, act=None, out_dtype=torch.float32)
    assert relerr(y3, y1 + y2) < 2e-2  # bf16 rounding of the summed input


def test_loftup_full_size_vs_oracle():
    """BASELINE config 2 geometry on one image: 448x448 guidance, 32x32x384 DINOv2 features -> [1,384,448,448];
    bf16 tensor-core mode against the fp32 oracle: cosine >= 0.999 (north_star)."""
    from isegprobe_b200.loftup import LoftUpUpsampler
    m = LoftUpUpsampler(None, n_dim=384)
    sd, cn = synth.loftup_state_dict(384, seed=0), synth.channelnorm_state_dict(384, seed=1)
    m.upsampler.upsampler.load_state_dict(sd, strict=True)
    m.upsampler.channelnorm.load_state_dict(cn, strict=True)
    m = m.to(DEV).eval()
    img = (synth.image_batch(1, 448, 448, seed=1) - 0.45) / 0.225
    lr = synth.lr_features(1, 384, 32, 32, seed=2)
    with torch.no_grad():
        out = m(source=lr.to(DEV), guidance=img.to(DEV)).cpu().float()
        want = oloft.loftup_forward(sd, lr, img, cn["norm.weight"], cn["norm.bias"])
    assert tuple(out.shape) == (1, 384, 448, 448)
    c = cosine(out, want)
    assert c >= 0.999, c
    per_pixel = F.cosine_similarity(out, want, dim=1)
    assert float(per_pixel.min()) > 0.99, float(per_pixel.min())


def test_jbu_full_size_vs_oracle():
    """BASELINE config 1 geometry on one image: 32x32x384 -> 512x512x384 through the four JBU stages (fp32 mode):
    1e-3 max relative error against the oracle (north_star)."""
    import isegprobe_b200 as isp
    up = isp.JBUFeatUpUpsampler("dinov2").to(DEV).eval()
    sd = ojbu.init_state_dict(384, seed=0)
    up.upsampler.load_state_dict(sd)
    src = synth.lr_features(1, 384, 32, 32, seed=2)
    gd = (synth.image_batch(1, 448, 448, seed=1) - 0.45) / 0.225
    with torch.no_grad():
        out = up(src.to(DEV), gd.to(DEV)).cpu()
        want = ojbu.jbu_stack_forward(sd, src, gd)
    assert tuple(out.shape) == (1, 384, 512, 512)
    err = float((out - want).abs().max() / want.abs().max())
    assert err < 1e-3, err


def test_flash_attention_backward_full_size_properties():
    """isp_attention_bwd_bf16_tc at the LoftUp shape (200 704 queries x 1024 keys, 4 heads x 101 in 112 columns), where no
    reference can hold the 822 M scores per head-set: size-independent identities of the softmax backward.
      * rows of P sum to one      =>  sum over keys of dV  =  sum over queries of dO          (per head, per channel)
      * rows of dS sum to zero    =>  sum over keys of dK  =  0   and  dQ . 1-direction: sum_k dS K uses the same dS
      * linearity in dO           =>  gradients for 2*dO are exactly twice those for dO (powers of two are exact in bf16)
    and equality with torch autograd on a sampled block of 256 queries (all keys)."""
    from isegprobe_b200 import _lib
    B, nh, rows, T, hd, HP = 1, 4, 448 * 448, 1024, 101, 112
    g = torch.Generator(device=DEV).manual_seed(1)

    def rnd(*shape, s=1.0):
        t = torch.zeros(*shape, device=DEV)
        t[..., :hd] = torch.randn(*shape[:-1], hd, generator=g, device=DEV) * s
        return t.to(torch.bfloat16)

    Q = rnd(B, rows, nh, HP, s=0.3).reshape(B * rows, nh * HP)
    dO = rnd(B, rows, nh, HP, s=0.5).reshape(B * rows, nh * HP)
    K, V = rnd(B, nh, T, HP), rnd(B, nh, T, HP)
    Kp = torch.zeros(B, nh, T, 128, dtype=torch.bfloat16, device=DEV)
    Kp[..., :HP] = K
    Vt = V.transpose(2, 3).contiguous()
    O = torch.empty_like(Q)
    lse = torch.zeros(B * nh * rows + 64, device=DEV)
    dvec = torch.zeros(B * nh * rows + 64, device=DEV)
    _call("isp_attention_bf16_tc_lse", Q, nh * HP, HP, Kp, Vt, O, nh * HP, HP, B, rows, nh, T, 1, lse)

    def backward(dOx):
        _call("isp_attention_rowdot_heads", dOx, nh * HP, O, nh * HP, dvec, B, rows, nh, HP)
        dK, dV = torch.zeros(B, nh, T, HP, device=DEV), torch.zeros(B, nh, T, HP, device=DEV)
        dQ = torch.zeros(B * rows, nh * HP, device=DEV)
        _call("isp_attention_bwd_bf16_tc", Q, nh * HP, dOx, nh * HP, K, V, lse, dvec, dK, dV, dQ, nh * HP, B, rows, nh, T, HP)
        return dK, dV, dQ

    dK, dV, dQ = backward(dO)
    want = dO.float().view(B, rows, nh, HP).sum(1)                      # [B, nh, HP]
    got = dV.sum(2)
    assert float((got - want).abs().max()) < 2e-3 * float(want.abs().max()) + 0.5, float((got - want).abs().max())
    # sum_k dK[k] = sum_q (sum_k dS[q,k]) Q[q] = 0 up to the bf16 rounding of dS (independent errors of 1024 keys add up):
    # the signed sum must vanish against the sum of magnitudes
    cancel = float((dK.sum(2).abs() / dK.abs().sum(2).clamp(min=1e-6))[..., :hd].max())
    assert cancel < 3e-3, cancel
    dK2, dV2, dQ2 = backward((dO.float() * 2).to(torch.bfloat16))
    # dO doubles exactly; P is identical; dS = P (dP - D) doubles up to the fp32 rounding of D = rowsum(dO.O)
    assert relerr(dV2, 2 * dV) < 1e-5
    assert relerr(dK2, 2 * dK) < 1e-2 and relerr(dQ2, 2 * dQ) < 1e-2
    # sampled query block against autograd (all keys): dQ rows are local to the block
    r0 = 137 * 128
    q = Q[r0:r0 + 256].float().view(256, nh, HP).permute(1, 0, 2).requires_grad_(True)
    s = q @ K[0].float().transpose(-1, -2)
    o = torch.softmax(s, -1) @ V[0].float()
    (o * dO[r0:r0 + 256].float().view(256, nh, HP).permute(1, 0, 2)).sum().backward()
    got_q = dQ[r0:r0 + 256].view(256, nh, HP).permute(1, 0, 2)
    assert cosine(got_q, q.grad) > 0.999 and relerr(got_q, q.grad) < 3e-2, (cosine(got_q, q.grad), relerr(got_q, q.grad))
