"""Feature upsamplers -- the reference's plugin API
(core/model/upsamplers/__init__.py:6-33): `forward(source, guidance) -> hr_feats`,
same registry keys, same constructor kwargs as models/*.py pass them.

Everything internal is channels-last: a module returns a tensor whose logical
shape is the reference's [B, C, H', W'] and whose memory is NHWC (a
`.permute(0, 3, 1, 2)` view), so the next stage (resize, head) reads it without a
transpose.  Compute is libisp_b200 only; there is no fallback.
"""
import math
from abc import ABC, abstractmethod

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


KERNEL_TIMING = False  # bench.py: record CUDA events around the dominant kernels (roofline)
KERNEL_TIMERS = {}


class timed_kernel:
    """with timed_kernel("name"): <launch>   -- CUDA events on the launching (current) stream."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if KERNEL_TIMING:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if KERNEL_TIMING:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            KERNEL_TIMERS.setdefault(self.name, []).append((self.e0, e1))
        return False


class BaseUpsampler(nn.Module, ABC):
    """core/model/upsamplers/__init__.py:6-11"""

    @abstractmethod
    def forward(self, source, guidance):
        pass


def to_nhwc_f32(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] float32 CUDA tensor (any strides) -> dense [B,H,W,C].  Free when the
    tensor already is channels-last (the reference's ViT adapters return permuted
    [B,h,w,C] views, DINOv2.py:545)."""
    if x.dtype != torch.float32:
        x = x.float()
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    if torch.is_grad_enabled() and x.requires_grad:
        return v.contiguous()  # autograd-tracked copy (training path only)
    B, C, H, W = x.shape
    out = torch.empty(B, H, W, C, dtype=torch.float32, device=x.device)
    sb, sc, sh, sw = x.stride()
    _lib.call("isp_nchw_to_nhwc_f32", _lib.dptr(x), _lib.dptr(out), B, C, H, W, sb, sc, sh, sw, _lib.stream_ptr())
    return out


def _bilinear_fwd(x_nhwc, size):
    B, H, W, C = x_nhwc.shape
    if C % 4:  # the kernel moves float4 channel groups: pad narrow maps (the 1-channel logits) to 4 channels
        return _bilinear_fwd(F.pad(x_nhwc, (0, 4 - C % 4)), size)[..., :C].contiguous()
    out = torch.empty(B, size[0], size[1], C, dtype=torch.float32, device=x_nhwc.device)
    _lib.call("isp_bilinear_ac_nhwc", _lib.dptr(x_nhwc), _lib.dptr(out), B, C, H, W, size[0], size[1], 0, C,
              _lib.stream_ptr())
    return out


class _BilinearFn(torch.autograd.Function):
    """align_corners=True resize with its adjoint as backward (isp_bilinear_ac_nhwc_bwd)."""

    @staticmethod
    def forward(ctx, x_nhwc, size):
        ctx.in_hw = (x_nhwc.shape[1], x_nhwc.shape[2])
        return _bilinear_fwd(x_nhwc.contiguous(), size)

    @staticmethod
    def backward(ctx, g):
        g = g.detach().float().contiguous()
        B, OH, OW, C = g.shape
        H, W = ctx.in_hw
        gin = torch.empty(B, H, W, C, dtype=torch.float32, device=g.device)
        _lib.call("isp_bilinear_ac_nhwc_bwd", _lib.dptr(g), _lib.dptr(gin), B, C, H, W, OH, OW, _lib.stream_ptr())
        return gin, None


def bilinear_align_corners_nhwc(x_nhwc: torch.Tensor, size) -> torch.Tensor:
    """F.interpolate(mode='bilinear', align_corners=True) on a dense NHWC fp32 tensor; differentiable w.r.t. x."""
    if torch.is_grad_enabled() and x_nhwc.requires_grad:
        return _BilinearFn.apply(x_nhwc, (int(size[0]), int(size[1])))
    return _bilinear_fwd(x_nhwc, size)


class IdentityUpsampler(BaseUpsampler):
    def forward(self, source, guidance):
        return source


class NearestUpsampler(BaseUpsampler):
    def forward(self, source, guidance):
        return F.interpolate(source, guidance.shape[2:], mode="nearest")


class BilinearUpsampler(BaseUpsampler):
    """basic_upsamplers.py:26-33 on the library's own resize kernel."""

    def forward(self, source, guidance):
        out = bilinear_align_corners_nhwc(to_nhwc_f32(source), tuple(guidance.shape[2:]))
        return out.permute(0, 3, 1, 2)


class BicubicUpsampler(BaseUpsampler):
    def forward(self, source, guidance):
        return F.interpolate(source, guidance.shape[2:], mode="bicubic")


# --------------------------------------------------------------------------- JBU
_JBU_FEAT_DIM = {"dinov2": 384, "dino16": 384, "vit": 384, "maskclip": 512, "clip": 512, "resnet50": 2048}


class _JBULearnedRange(nn.Module):
    """Parameter container with upstream FeatUp's names/shapes (featup/upsamplers.py
    JBULearnedRange: guidance_dim=3, key_dim=32, radius=3)."""

    def __init__(self):
        super().__init__()
        self.range_temp = nn.Parameter(torch.tensor(0.0))
        self.range_proj = nn.Sequential(nn.Conv2d(3, 32, 1), nn.GELU(), nn.Dropout2d(0.1), nn.Conv2d(32, 32, 1))
        self.fixup_proj = nn.Sequential(nn.Conv2d(52, 49, 1), nn.GELU(), nn.Dropout2d(0.1), nn.Conv2d(49, 49, 1))
        self.sigma_spatial = nn.Parameter(torch.tensor(1.0))


class _JBUStack(nn.Module):
    def __init__(self, feat_dim):
        super().__init__()
        self.up1, self.up2, self.up3, self.up4 = (_JBULearnedRange() for _ in range(4))
        self.fixup_proj = nn.Sequential(nn.Dropout2d(0.2), nn.Conv2d(feat_dim, feat_dim, kernel_size=1))


class JBUFeatUpUpsampler(BaseUpsampler):
    """FeatUp's learned JBU stack (x16), same constructor as the reference wrapper
    (core/model/upsamplers/JBUFeatUp.py:9-32).  The reference pulls weights and code
    from torch.hub (network); here the stack is built locally with upstream's
    parameter layout (so a FeatUp `upsampler` state dict loads with
    `self.upsampler.load_state_dict`) and random-initialised otherwise.
    `weights` (extension): path to such a state dict."""

    def __init__(self, backbone_type: str = None, use_norm: bool = True, weights: str = None) -> None:
        super().__init__()
        self.backbone_type = backbone_type
        self.use_norm = use_norm
        assert self.backbone_type in _JBU_FEAT_DIM, f"Invalid model type: {self.backbone_type}"
        self.feat_dim = _JBU_FEAT_DIM[self.backbone_type]
        self.upsampler = _JBUStack(self.feat_dim)
        if weights is not None:
            self.upsampler.load_state_dict(torch.load(weights, map_location="cpu"))

    @staticmethod
    def _flat(p):
        return p.detach().reshape(p.shape[0], -1).float().contiguous()

    def _scalar(self, p: torch.Tensor) -> float:
        """Host value of a 0-d parameter, cached per parameter version (avoids a
        device sync on every forward of the frozen stack)."""
        cache = self.__dict__.setdefault("_scalar_cache", {})
        key = id(p)
        hit = cache.get(key)
        if hit is None or hit[0] != p._version or hit[1] != p.data_ptr():
            hit = (p._version, p.data_ptr(), float(p.detach()))
            cache[key] = hit
        return hit[2]

    def _filters(self, up: _JBULearnedRange, guidance: torch.Tensor, GH: int, GW: int, ld: int, masks=None) -> torch.Tensor:
        """Combined per-pixel 7x7 kernels of one stage from the guidance image: [B,GH,GW,ld], ld = 56 (rows padded
        to 8, the layout isp_adaptive_conv_fwd fetches with one TMA box) or 49 (dense, for the input-gradient kernel).
        masks = (range [B,32], fixup [B,49]): train() mode -- the two Dropout2d layers' multiplicative masks.  A mask on the
        hidden channels of `conv -> GELU -> Dropout2d -> conv` is a per-sample column scaling of the second conv's weight, so
        the batch is processed one sample at a time with its own folded weights."""
        if masks is not None:
            B = guidance.shape[0]
            filt = torch.empty(B, GH, GW, ld, dtype=torch.float32, device=guidance.device)
            for b in range(B):
                filt[b:b + 1] = self._filters_batch(up, guidance[b:b + 1], GH, GW, ld, (masks[0][b], masks[1][b]))
            return filt
        return self._filters_batch(up, guidance, GH, GW, ld)

    def _filters_batch(self, up, guidance, GH, GW, ld, fold=None):
        B = guidance.shape[0]
        dev, st = guidance.device, _lib.stream_ptr()
        g = torch.empty(B, GH, GW, 4, dtype=torch.float32, device=dev)
        sb, sc, sh, sw = guidance.stride()
        _lib.call("isp_jbu_pool_guidance", _lib.dptr(guidance), _lib.dptr(g), B, guidance.shape[2], guidance.shape[3],
                  GH, GW, sb, sc, sh, sw, st)
        proj = torch.empty(B, GH, GW, 32, dtype=torch.float32, device=dev)
        rp, fp = up.range_proj, up.fixup_proj
        w0, b0, w1, b1 = self._flat(rp[0].weight), rp[0].bias.detach().float(), self._flat(rp[3].weight), rp[3].bias.detach().float()
        if fold is not None:
            w1 = (w1 * fold[0].to(w1)[None, :]).contiguous()
        _lib.call("isp_jbu_range_proj", _lib.dptr(g), _lib.dptr(proj), B * GH * GW, _lib.dptr(w0), _lib.dptr(b0),
                  _lib.dptr(w1), _lib.dptr(b1), st)
        filt = torch.empty(B, GH, GW, ld, dtype=torch.float32, device=dev)
        temp = min(max(math.exp(self._scalar(up.range_temp)), 1e-4), 1e4)
        f0, fb0, f1, fb1 = self._flat(fp[0].weight), fp[0].bias.detach().float(), self._flat(fp[3].weight), fp[3].bias.detach().float()
        if fold is not None:
            f1 = (f1 * fold[1].to(f1)[None, :]).contiguous()
        _lib.call("isp_jbu_filters", _lib.dptr(proj), _lib.dptr(g), _lib.dptr(filt), B, GH, GW, float(temp),
                  self._scalar(up.sigma_spatial), _lib.dptr(f0), _lib.dptr(fb0), _lib.dptr(f1), _lib.dptr(fb1), ld, st)
        return filt

    def _stage(self, up: _JBULearnedRange, src: torch.Tensor, guidance: torch.Tensor, masks=None) -> torch.Tensor:
        B, h, w, C = src.shape
        GH, GW = 2 * h, 2 * w
        dev, st = src.device, _lib.stream_ptr()
        filt = self._filters(up, guidance, GH, GW, 56, masks)
        hr = torch.empty(B, GH + 6, GW + 6, C, dtype=torch.float32, device=dev)
        _lib.call("isp_jbu_bicubic2x_reflectpad", _lib.dptr(src), _lib.dptr(hr), B, h, w, C, st)
        out = torch.empty(B, GH, GW, C, dtype=torch.float32, device=dev)
        with timed_kernel(f"adaptive_conv_{GH}"):
            _lib.call("isp_adaptive_conv_fwd", _lib.dptr(hr), _lib.dptr(filt), _lib.dptr(out), B, GH, GW, C, 56, st)
        return out

    def forward(self, source: torch.Tensor, guidance: torch.Tensor) -> torch.Tensor:
        return self.forward_resized(source, guidance, None)

    # dtype of the returned features; ISegPipeline switches to bf16 when a head consumes them (the head rounds its input to
    # bf16 anyway: same bits, one 7.4 GB / 16-image conversion pass less)
    out_dtype = torch.float32
    dropout_masks = None  # tests: a dict of explicit masks (keys as oracle.jbu.dropout2d_masks) instead of fresh draws

    def _draw_masks(self, B: int, dev):
        """train() mode (the reference's trainer puts the frozen stack in train(), core/training/trainer.py:213-214): the
        multiplicative masks of FeatUp's Dropout2d layers -- range_proj.2 and fixup_proj.2 of every JBULearnedRange (p = 0.1)
        and JBUStack.fixup_proj.0 (p = 0.2) -- one Bernoulli draw per (sample, channel), kept values scaled by 1 / (1 - p)."""
        if not self.training:
            return None
        if self.dropout_masks is not None:
            return {k: v.to(dev, torch.float32) for k, v in self.dropout_masks.items()}

        def m(c, p):
            return torch.bernoulli(torch.full((B, c), 1.0 - p, device=dev)) / (1.0 - p)

        out = {"final": m(self.feat_dim, 0.2)}
        for k in range(1, 5):
            out[f"up{k}.range"], out[f"up{k}.fixup"] = m(32, 0.1), m(49, 0.1)
        return out

    @staticmethod
    def _split_rows(x2d: torch.Tensor) -> torch.Tensor:
        """fp32 [M,C] -> bf16 [M,3C] = [hi | lo | hi] with x = hi + lo to 2^-17."""
        hi = x2d.to(torch.bfloat16)
        lo = (x2d - hi.float()).to(torch.bfloat16)
        return torch.cat((hi, lo, hi), 1)

    @staticmethod
    def _split_weight(Wm: torch.Tensor) -> torch.Tensor:
        """fp32 [N,K] -> bf16 [N,3K] = [hi | hi | lo]: with _split_rows the products hi*hi + lo*hi + hi*lo."""
        hi = Wm.to(torch.bfloat16)
        lo = (Wm - hi.float()).to(torch.bfloat16)
        return torch.cat((hi, hi, lo), 1).contiguous()

    def _mix_channels(self, x: torch.Tensor, transpose: bool, mask=None) -> torch.Tensor:
        """y = x + 0.1 (x . m) W^T  (forward, m = the Dropout2d(0.2) mask of JBUStack.fixup_proj under train(), else 1) or its
        adjoint  g + 0.1 m . (g W)  (transpose=True), on fp32 NHWC [B,h,w,C] at SOURCE resolution.  tcgen05 GEMM on
        split-bf16 operands (three products per element, fp32 accumulate, fp32 residual): error ~1e-5 of the 0.1-scaled
        term.  Channel counts that are not a multiple of 8: the fp32 SIMT GEMM."""
        from . import tc
        B, h, w, C = x.shape
        conv = self.upsampler.fixup_proj[1]
        Wf = self._flat(conv.weight).to(x.device)
        x2 = x.reshape(B * h * w, C)
        if C % 8:
            if mask is not None:
                raise NotImplementedError("JBUFeatUpUpsampler.train() needs a channel count that is a multiple of 8")
            out = torch.empty_like(x2)
            Wm = Wf.t().contiguous() if transpose else Wf
            _lib.call("isp_gemm_f32_simt", _lib.dptr(x2), _lib.dptr(Wm), _lib.dptr(None), _lib.dptr(x2), 0.1, _lib.dptr(out), B * h * w,
                      C, C, _lib.stream_ptr())
            return out.view(B, h, w, C)
        if mask is None:
            key = (conv.weight._version, str(x.device), transpose)
            cache = self.__dict__.setdefault("_mix_w", {})
            if cache.get("key" + str(transpose)) != key:
                cache["key" + str(transpose)] = key
                cache[transpose] = self._split_weight(Wf.t().contiguous() if transpose else Wf)
            return tc.gemm(self._split_rows(x2), cache[transpose], resid=x2, alpha=0.1, out_dtype=torch.float32, N=C,
                           K=3 * C).view(B, h, w, C)
        out = torch.empty_like(x)
        for b in range(B):  # per-sample column (forward) / row (adjoint) scaling of the weight
            Wb = (Wf.t() * mask[b][:, None]) if transpose else (Wf * mask[b][None, :])
            xb = x[b].reshape(h * w, C)
            tc.gemm(self._split_rows(xb), self._split_weight(Wb.contiguous()), resid=xb, alpha=0.1, out_dtype=torch.float32,
                    N=C, K=3 * C, out=out[b].view(h * w, C))
        return out

    def forward_resized(self, source: torch.Tensor, guidance: torch.Tensor, size=None) -> torch.Tensor:
        """`forward` followed by the bilinear (align_corners=True) resize to `size` that the reference
        applies right after the upsampler (core/model/iseg_probe_model.py:120-129).  The final
        `fixup_proj(x) * 0.1 + x` of JBUStack.forward is the per-pixel channel map  x -> (I + 0.1 W) x + 0.1 b.  Every
        other step of the stack -- bicubic x2, reflect pad, AdaptiveConv with its guidance-only 7x7 kernels, the resize --
        is a spatial linear map applied to each channel alike, so the channel map commutes with all of them (the bias
        because the resize weights sum to 1).  It is therefore applied ONCE TO THE 32x32 SOURCE (`_mix_channels`:
        16 384 instead of 3.2 M pixels at B = 16, in split-bf16 arithmetic that is exact to ~2^-17), and 0.1 b is added by
        the resize pass: the 448^2 1x1-conv GEMM and its bf16 operand copy (2.7 ms of a 19.7 ms step) do not exist.
        size=None keeps the reference's 16x output."""
        masks = self._draw_masks(source.shape[0], source.device)
        if torch.is_grad_enabled() and source.requires_grad:  # frozen stack, but the features' gradient flows through
            return _JBUFn.apply(self, source, guidance, size, masks)
        return self._forward_impl(source, guidance, size, masks)

    def _backward_impl(self, grad_out: torch.Tensor, guidance: torch.Tensor, src_hw, size, masks=None):
        """d(loss)/d(source) of forward_resized: the stack is linear in `source` (the 7x7 kernels depend on the guidance
        image only), so the backward is the chain of adjoints -- 1x1 conv dgrad (+ identity), resize adjoint, and per
        stage isp_adaptive_conv_grad_input followed by the bicubic / reflect-pad adjoint."""
        dx = grad_out.detach().float().permute(0, 2, 3, 1).contiguous()  # [B,OH,OW,C]
        B, OH, OW, C = dx.shape
        dev, st = dx.device, _lib.stream_ptr()
        guidance = guidance.detach().float()
        GH, GW = 16 * src_hw[0], 16 * src_hw[1]
        if (OH, OW) != (GH, GW):
            d = torch.empty(B, GH, GW, C, dtype=torch.float32, device=dev)
            _lib.call("isp_bilinear_ac_nhwc_bwd", _lib.dptr(dx), _lib.dptr(d), B, C, GH, GW, OH, OW, st)
            dx = d
        for k, up in zip((4, 3, 2, 1), (self.upsampler.up4, self.upsampler.up3, self.upsampler.up2, self.upsampler.up1)):
            filt = self._filters(up, guidance, GH, GW, 49,
                                 None if masks is None else (masks[f"up{k}.range"], masks[f"up{k}.fixup"]))
            dpad = torch.empty(B, GH + 6, GW + 6, C, dtype=torch.float32, device=dev)
            _lib.call("isp_adaptive_conv_grad_input", _lib.dptr(dx), _lib.dptr(filt), _lib.dptr(dpad), B, GH, GW, C, st)
            GH, GW = GH // 2, GW // 2
            dx = torch.empty(B, GH, GW, C, dtype=torch.float32, device=dev)
            _lib.call("isp_jbu_bicubic2x_reflectpad_bwd", _lib.dptr(dpad), _lib.dptr(dx), B, GH, GW, C, st)
        dx = self._mix_channels(dx, True, None if masks is None else masks["final"])  # adjoint of the channel map
        return dx.permute(0, 3, 1, 2)

    def _forward_impl(self, source: torch.Tensor, guidance: torch.Tensor, size=None, masks=None) -> torch.Tensor:
        x = to_nhwc_f32(source.detach())
        guidance = guidance.detach().float()
        # fixup_proj(x) * 0.1 + x of JBUStack.forward, commuted to the source (see forward_resized)
        x = self._mix_channels(x, False, None if masks is None else masks["final"])
        for k, up in enumerate((self.upsampler.up1, self.upsampler.up2, self.upsampler.up3, self.upsampler.up4), 1):
            x = self._stage(up, x, guidance, None if masks is None else (masks[f"up{k}.range"], masks[f"up{k}.fixup"]))
        B, H, W, C = x.shape
        OH, OW = (H, W) if size is None else (int(size[0]), int(size[1]))
        conv = self.upsampler.fixup_proj[1]
        key = (conv.bias._version, str(x.device))
        if getattr(self, "_fix_key", None) != key:
            self._fix_b = (0.1 * conv.bias.detach().float()).contiguous().to(x.device)
            self._fix_key = key
        if C % 4:
            if (OH, OW) != (H, W):
                x = bilinear_align_corners_nhwc(x, (OH, OW))
            return (x + self._fix_b).permute(0, 3, 1, 2)
        bf = self.out_dtype == torch.bfloat16 and C % 8 == 0
        out = torch.empty(B, OH, OW, C, dtype=torch.bfloat16 if bf else torch.float32, device=x.device)
        with timed_kernel(f"resize_bias_{OH}"):
            _lib.call("isp_bilinear_ac_nhwc_bias", _lib.dptr(x), _lib.dptr(out), int(bf), _lib.dptr(self._fix_b), B, C, H, W,
                      OH, OW, _lib.stream_ptr())
        return out.permute(0, 3, 1, 2)


class _JBUFn(torch.autograd.Function):
    """Frozen JBU stack with an input gradient for `source` (the backbone features)."""

    @staticmethod
    def forward(ctx, mod, source, guidance, size, masks=None):
        ctx.mod, ctx.guidance, ctx.size, ctx.masks = mod, guidance, size, masks
        ctx.src_hw = (source.shape[2], source.shape[3])
        return mod._forward_impl(source, guidance, size, masks)

    @staticmethod
    def backward(ctx, grad_out):
        return None, ctx.mod._backward_impl(grad_out, ctx.guidance, ctx.src_hw, ctx.size, ctx.masks), None, None, None


UPSAMPLER_REGISTRY = {
    "identity": IdentityUpsampler,
    "nearest": NearestUpsampler,
    "bilinear": BilinearUpsampler,
    "bicubic": BicubicUpsampler,
    "jbu_featup": JBUFeatUpUpsampler,
}
