"""LiFT x2 conv upsampler -- same plugin surface as the reference
`LiFTUpsampler(lift_path, n_dim=384, patch=14)` (core/model/upsamplers/LiFT.py:139-146) and the
same state-dict layout as `LiFT(n_dim, patch)` (LiFT.py:47-90).  eval(): BatchNorm folded into the convs; train() (how the
reference's trainer runs the frozen upsampler, core/training/trainer.py:213-214): batch statistics, running statistics updated,
and the BatchNorm backward of the two double-conv layers in the source gradient.

  image --conv3x3 s2 (3->32) --conv3x3 s2 (32->32)--> adaptive max-pool to (2h,2w) = imgs_1
  imgs_1 --conv3x3 s2 (32->32)--> imgs_2 ; cat[source, imgs_2] (C+32)
        --ConvTranspose2d k2 s2 (one GEMM to 4*(C+32)/2 columns + pixel shuffle)--> cat with imgs_1
        --conv3x3+BN+ReLU x2 (tcgen05 implicit GEMM)--> conv1x1 --> [B,C,2h,2w]
"""
import torch
import torch.nn as nn

from . import _lib, tc
from .upsamplers import BaseUpsampler


def _call(name, *a):
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


class _DoubleConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.double_conv = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout),
                                         nn.ReLU(inplace=True), nn.Conv2d(cout, cout, 3, padding=1, bias=False),
                                         nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cin // 2, kernel_size=2, stride=2)
        self.conv_1 = _DoubleConv(cin // 2 + 32, cout // 2)


class _LiFTParams(nn.Module):
    """Keys/shapes of reference LiFT(in_channels, patch_size).state_dict()."""

    def __init__(self, C, patch):
        super().__init__()
        if patch not in (8, 14, 16):
            raise ValueError("ERROR: patch size %i not currently supported" % patch)  # LiFT.py:83-85
        self.up1 = _Up(C + 32, C)
        self.outc = nn.Conv2d(C // 2, C, kernel_size=1)
        self.image_convs_1 = nn.Sequential(nn.Conv2d(3, 32, 3, padding=1, stride=2), nn.BatchNorm2d(32),
                                           nn.ReLU(inplace=True), nn.Conv2d(32, 32, 3, padding=1, stride=2),
                                           nn.BatchNorm2d(32), nn.ReLU(inplace=True))
        self.image_convs_2 = nn.Sequential(nn.Conv2d(32, 32, 3, padding=1, stride=2), nn.BatchNorm2d(32),
                                           nn.ReLU(inplace=True))


def _fold(conv, bn):
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
    w = conv.weight.detach().float() * s[:, None, None, None]
    b0 = conv.bias.detach().float() if conv.bias is not None else torch.zeros_like(s)
    return w, (b0 - bn.running_mean.float()) * s + bn.bias.detach().float()


class LiFTUpsampler(BaseUpsampler):
    """`lift_path=None` (extension) keeps the random initialisation (reference always loads)."""

    def __init__(self, lift_path: str = None, n_dim: int = 384, patch: int = 14):
        super().__init__()
        self.n_dim = n_dim
        self.lift = _LiFTParams(n_dim, patch)
        if lift_path is not None:
            sd = torch.load(lift_path, map_location="cpu")
            sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}  # LiFT.py:129-132
            self.lift.load_state_dict(sd)
        self._packed = None

    def _version(self):
        return sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers())

    def _pack(self, dev):
        key = (str(dev), self._version())
        if self._packed is not None and self._packed["key"] == key:
            return self._packed
        L, C = self.lift, self.n_dim
        f32 = lambda t: t.detach().float().contiguous().to(dev)
        P = {"key": key}
        for name, conv, bn in (("ic1a", L.image_convs_1[0], L.image_convs_1[1]),
                               ("ic1b", L.image_convs_1[3], L.image_convs_1[4]),
                               ("ic2", L.image_convs_2[0], L.image_convs_2[1])):
            w, b = _fold(conv, bn)
            P[name] = (f32(w), f32(b))
        up = L.up1.up  # weight [Cin, Cout, 2, 2]
        Cin, Co = up.weight.shape[0], up.weight.shape[1]
        wt = up.weight.detach().float().permute(2, 3, 1, 0).reshape(4 * Co, Cin)  # row (dy*2+dx)*Co + co
        P["up_w"] = tc.pack_linear_weight(wt).to(dev)
        P["up_b"] = f32(up.bias.detach().float().repeat(4))
        dc = L.up1.conv_1.double_conv
        w, b = _fold(dc[0], dc[1])
        P["dc1"] = (tc.pack_conv3x3_weight(w).to(dev), f32(b))
        w, b = _fold(dc[3], dc[4])
        P["dc2"] = (tc.pack_conv3x3_weight(w).to(dev), f32(b))
        P["out_w"] = tc.pack_linear_weight(L.outc.weight.detach().float().reshape(C, C // 2)).to(dev)
        P["out_b"] = f32(L.outc.bias)
        # train() mode: the un-folded convolutions (BatchNorm then uses the statistics of their raw outputs)
        zero32 = torch.zeros(32, device=dev)
        for name, conv in (("ic1a", L.image_convs_1[0]), ("ic1b", L.image_convs_1[3]), ("ic2", L.image_convs_2[0])):
            P[name + "_raw"] = (f32(conv.weight), f32(conv.bias) if conv.bias is not None else zero32)
        P["dc1_raw"] = tc.pack_conv3x3_weight(dc[0].weight.detach().float()).to(dev)
        P["dc2_raw"] = tc.pack_conv3x3_weight(dc[3].weight.detach().float()).to(dev)
        P["zero_bias"] = torch.zeros(C, device=dev)
        self._packed = P
        return P

    @staticmethod
    def _bn_update(bn, mean, var, count):
        """nn.BatchNorm2d's train() side effects: running statistics (unbiased variance) with momentum, batch counter."""
        with torch.no_grad():
            m = bn.momentum if bn.momentum is not None else 0.1
            bn.running_mean.mul_(1 - m).add_(mean.to(bn.running_mean.dtype), alpha=m)
            bn.running_var.mul_(1 - m).add_((var * (count / max(count - 1, 1))).to(bn.running_var.dtype), alpha=m)
            bn.num_batches_tracked += 1

    def _bn_relu_f32(self, x, bn):
        """Training-mode BatchNorm2d + ReLU of a small fp32 NHWC image-branch tensor [B,H,W,32], in place (LiFT.py:69-90)."""
        flat = x.view(-1, x.shape[-1])
        var, mean = torch.var_mean(flat, dim=0, unbiased=False)
        self._bn_update(bn, mean, var, flat.shape[0])
        scale = bn.weight.detach().float() / torch.sqrt(var + bn.eps)
        flat.mul_(scale).add_(bn.bias.detach().float() - mean * scale).clamp_(min=0)
        return x

    def _bn_relu_bf16(self, z, bn, C):
        """Training-mode BatchNorm2d + ReLU of a raw tcgen05 conv output (NHWC bf16, ld = z.shape[-1]): per-channel batch
        statistics from deterministic slab partials (isp_col_moments_bf16, summed in fp64), y = max(z * scale + shift, 0)
        written to a NEW tensor (the backward needs z).  Returns (y, scale, mean, rstd)."""
        dev, ld = z.device, z.shape[-1]
        M = z.numel() // ld
        part = torch.empty((M + 1023) // 1024, 2, C, dtype=torch.float32, device=dev)  # ISP_COL_MOMENTS_SLAB_ROWS
        _call("isp_col_moments_bf16", z, ld, M, C, part, C)
        tot = part.sum(0, dtype=torch.float64)
        mean = tot[0] / M
        var = (tot[1] / M - mean * mean).clamp_(min=0.0)
        self._bn_update(bn, mean, var, M)
        rstd = 1.0 / torch.sqrt(var + bn.eps)
        scale = bn.weight.detach().double() * rstd
        shift = bn.bias.detach().double() - mean * scale
        y = z.clone()
        _call("isp_bn_relu_rows_bf16", y, ld, scale.float().contiguous(), shift.float().contiguous(), M, C, None)
        return y, scale.float(), mean.float(), rstd.float()

    def forward(self, source: torch.Tensor, guidance: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled() and source.requires_grad:  # frozen weights, but the features' gradient flows through
            return _LiFTFn.apply(self, source, guidance)
        return self._forward_impl(source, guidance, None)

    def _forward_impl(self, source: torch.Tensor, guidance: torch.Tensor, saved) -> torch.Tensor:
        dev = guidance.device
        P = self._pack(dev)
        img, src = guidance.detach().float(), source.detach().float()
        B, _, H, W = img.shape
        C, h, w = src.shape[1], src.shape[2], src.shape[3]
        bf = torch.bfloat16
        H1, W1 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        train = self.training
        L = self.lift
        a = torch.empty(B, H1, W1, 32, device=dev)
        H2, W2 = (H1 - 1) // 2 + 1, (W1 - 1) // 2 + 1
        i1 = torch.empty(B, H2, W2, 32, device=dev)
        if train:
            _call("isp_conv3x3_s2_c32_raw", img, *img.stride(), P["ic1a_raw"][0], P["ic1a_raw"][1], a, B, 3, H, W)
            self._bn_relu_f32(a, L.image_convs_1[1])
            _call("isp_conv3x3_s2_c32_raw", a, H1 * W1 * 32, 1, W1 * 32, 32, P["ic1b_raw"][0], P["ic1b_raw"][1], i1, B, 32, H1, W1)
            self._bn_relu_f32(i1, L.image_convs_1[4])
        else:
            _call("isp_conv3x3_s2_c32", img, *img.stride(), P["ic1a"][0], P["ic1a"][1], a, B, 3, H, W)
            _call("isp_conv3x3_s2_c32", a, H1 * W1 * 32, 1, W1 * 32, 32, P["ic1b"][0], P["ic1b"][1], i1, B, 32, H1, W1)
        GH, GW = 2 * h, 2 * w
        i1p = torch.empty(B, GH, GW, 32, device=dev)
        _call("isp_adaptive_maxpool_nhwc", i1, i1p, B, 32, H2, W2, GH, GW)
        i2 = torch.empty(B, h, w, 32, device=dev)
        if train:
            _call("isp_conv3x3_s2_c32_raw", i1p, GH * GW * 32, 1, GW * 32, 32, P["ic2_raw"][0], P["ic2_raw"][1], i2, B, 32, GH, GW)
            self._bn_relu_f32(i2, L.image_convs_2[1])
        else:
            _call("isp_conv3x3_s2_c32", i1p, GH * GW * 32, 1, GW * 32, 32, P["ic2"][0], P["ic2"][1], i2, B, 32, GH, GW)
        # cat[source, imgs_2] -> bf16 [B*h*w, C+32]
        Cc = C + 32
        cat1 = torch.empty(B, h, w, Cc, dtype=bf, device=dev)
        _call("isp_copy_channels", src, 0, *src.stride(), cat1, 1, h * w * Cc, w * Cc, Cc, B, C, h, w)
        _call("isp_copy_channels", i2, 0, h * w * 32, 1, w * 32, 32, cat1[..., C:], 1, h * w * Cc, w * Cc, Cc, B, 32, h, w)
        Co = Cc // 2
        up = tc.gemm(cat1.view(B * h * w, Cc), P["up_w"], bias=P["up_b"], out_dtype=bf, N=4 * Co, K=Cc)
        # pixel shuffle into cat[up, imgs_1] -> bf16 [B,2h,2w,Co+32]
        C2 = Co + 32
        cat2 = torch.empty(B, GH, GW, C2, dtype=bf, device=dev)
        for dy in range(2):
            for dx in range(2):
                s = up[:, (dy * 2 + dx) * Co:]
                d = cat2[:, dy::2, dx::2]
                _call("isp_copy_channels", s, 1, h * w * 4 * Co, 1, w * 4 * Co, 4 * Co, d, 1, GH * GW * C2,
                      2 * GW * C2, 2 * C2, B, Co, h, w)
        _call("isp_copy_channels", i1p, 0, GH * GW * 32, 1, GW * 32, 32, cat2[..., Co:], 1, GH * GW * C2, GW * C2, C2,
              B, 32, GH, GW)
        Ch = C // 2
        ldh = tc.round_up(Ch, 8)
        if train:
            dc = L.up1.conv_1.double_conv
            z1 = tc.conv3x3(cat2, P["dc1_raw"], P["zero_bias"], C2, Ch, act=None, ldy=ldh)
            x1, s1, m1, r1 = self._bn_relu_bf16(z1, dc[1], Ch)
            z2 = tc.conv3x3(x1, P["dc2_raw"], P["zero_bias"], Ch, Ch, act=None, ldy=ldh)
            x, s2, m2, r2 = self._bn_relu_bf16(z2, dc[4], Ch)
            self._packed = None  # the running statistics moved: the eval-mode fold is stale
            if saved is not None:
                saved["bn"] = ((z1, s1, m1, r1), (z2, s2, m2, r2))
        else:
            x1 = tc.conv3x3(cat2, P["dc1"][0], P["dc1"][1], C2, Ch, act="relu", ldy=ldh)
            x = tc.conv3x3(x1, P["dc2"][0], P["dc2"][1], Ch, Ch, act="relu", ldy=ldh)
        if saved is not None:  # ReLU outputs (masks of the two dgrads) and the shapes
            saved.update({"x1": x1, "x2": x, "B": B, "h": h, "w": w, "C": C})
        out = tc.gemm(x.view(B * GH * GW, -1), P["out_w"], bias=P["out_b"], out_dtype=torch.float32, N=C, K=Ch)
        return out.view(B, GH, GW, C).permute(0, 3, 1, 2)


    # ---------------------------------------------------------------- activation backward (d loss / d source)
    def _pack_bwd(self, dev):
        P = self._pack(dev)
        if "bwd" not in P:
            L, C = self.lift, self.n_dim
            dc = L.up1.conv_1.double_conv
            w1, _ = _fold(dc[0], dc[1])
            w2, _ = _fold(dc[3], dc[4])
            flipT = lambda w: tc.pack_conv3x3_weight(w.detach().float().flip(2, 3).transpose(0, 1).contiguous()).to(dev)
            up = L.up1.up
            Cin, Co = up.weight.shape[0], up.weight.shape[1]
            wt = up.weight.detach().float().permute(2, 3, 1, 0).reshape(4 * Co, Cin)
            P["bwd"] = {"dc1T": flipT(w1), "dc2T": flipT(w2), "dc1T_raw": flipT(dc[0].weight), "dc2T_raw": flipT(dc[3].weight),
                        "outT": tc.pack_linear_weight(L.outc.weight.detach().float().reshape(C, C // 2).t().contiguous()).to(dev),
                        "upT": tc.pack_linear_weight(wt.t().contiguous()).to(dev)}
        return P, P["bwd"]

    def _backward_impl(self, saved, grad_out):
        """Chain of dgrads back to the source channels: 1x1 conv, two 3x3 convs (ReLU masks), inverse pixel shuffle,
        transposed conv (LiFT.py:106-122; the image branch carries no gradient)."""
        dev, bf = grad_out.device, torch.bfloat16
        P, PB = self._pack_bwd(dev)
        B, h, w, C = saved["B"], saved["h"], saved["w"], saved["C"]
        GH, GW = 2 * h, 2 * w
        Cc, Ch = C + 32, C // 2
        Co = Cc // 2
        C2 = Co + 32
        x1, x2 = saved["x1"], saved["x2"]
        ldh = x2.shape[3]
        g = grad_out.detach().float().permute(0, 2, 3, 1).contiguous().view(B * GH * GW, C).to(bf)
        dz = tc.gemm(g, PB["outT"], out_dtype=bf, N=Ch, K=C, ldd=ldh).view(B, GH, GW, ldh)
        dz = dz * (x2 > 0)  # ReLU of the second 3x3 conv
        d1 = torch.empty(B, GH, GW, ldh, dtype=bf, device=dev)
        if ldh > Ch:
            d1.zero_()
        ld2 = tc.round_up(C2, 8)
        dcat2 = torch.empty(B, GH, GW, ld2, dtype=bf, device=dev)
        if "bn" in saved:
            # train(): the two BatchNorm layers normalised with batch statistics, which depend on the input themselves:
            # dz = scale * (dy - mean(dy) - zhat * mean(dy * zhat)) per channel over all pixels (zhat = (z - mean) * rstd),
            # then the raw (un-folded) convolutions' data gradients
            def bn_bwd(dy, zs):
                z, scale, mean, rstd = zs
                dyf, zhat = dy.reshape(-1, ldh)[:, :Ch].float(), (z.reshape(-1, ldh)[:, :Ch].float() - mean) * rstd
                out = torch.zeros(dyf.shape[0], ldh, dtype=bf, device=dev) if ldh > Ch else torch.empty(dyf.shape[0], ldh, dtype=bf, device=dev)
                out[:, :Ch] = (scale * (dyf - dyf.mean(0) - zhat * (dyf * zhat).mean(0))).to(bf)
                return out.view(B, GH, GW, ldh)
            bn1, bn2 = saved["bn"]
            _call("isp_conv3x3_dgrad_bf16_tc", bn_bwd(dz, bn2), PB["dc2T_raw"], None, 0, d1, 1, B, GH, GW, Ch, ldh, Ch, ldh)
            d1 = d1 * (x1 > 0)  # ReLU of the first 3x3 conv
            _call("isp_conv3x3_dgrad_bf16_tc", bn_bwd(d1, bn1), PB["dc1T_raw"], None, 0, dcat2, 1, B, GH, GW, Ch, ldh, C2, ld2)
        else:
            _call("isp_conv3x3_dgrad_bf16_tc", dz.contiguous(), PB["dc2T"], x1, ldh, d1, 1, B, GH, GW, Ch, ldh, Ch, ldh)
            _call("isp_conv3x3_dgrad_bf16_tc", d1, PB["dc1T"], None, 0, dcat2, 1, B, GH, GW, Ch, ldh, C2, ld2)
        dup = torch.empty(B, h, w, 4 * Co, dtype=bf, device=dev)  # inverse of the pixel shuffle
        for dy in range(2):
            for dx in range(2):
                dup[..., (dy * 2 + dx) * Co:(dy * 2 + dx + 1) * Co] = dcat2[:, dy::2, dx::2, :Co]
        dcat1 = tc.gemm(dup.view(B * h * w, 4 * Co), PB["upT"], out_dtype=torch.float32, N=Cc, K=4 * Co)
        return dcat1[:, :C].reshape(B, h, w, C).permute(0, 3, 1, 2)


class _LiFTFn(torch.autograd.Function):
    """Frozen LiFT with an input gradient for `source`."""

    @staticmethod
    def forward(ctx, mod, source, guidance):
        saved = {}
        out = mod._forward_impl(source, guidance, saved)
        ctx.mod, ctx.saved = mod, saved
        return out

    @staticmethod
    def backward(ctx, grad_out):
        d = ctx.mod._backward_impl(ctx.saved, grad_out)
        ctx.saved = None
        return None, d, None
