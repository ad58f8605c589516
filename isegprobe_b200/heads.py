"""IS heads -- the reference's HEAD_REGISTRY surface (core/model/heads/__init__.py:11-15,
conv_heads.py:10-73, base_head.py:8-18) on the tcgen05 implicit-GEMM conv.

Parameter names follow mmcv's ConvModule (`convs.{i}.conv.weight/bias`, `classifier.*`) so
reference IS checkpoints load with load_state_dict (SURVEY 8b "names that leak")."""
import torch
import torch.nn as nn

from . import _lib, tc
from .upsamplers import to_nhwc_f32


def _call(name, *a):
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


class _ConvModule(nn.Module):
    """mmcv.cnn.ConvModule defaults: Conv2d(bias=True) under `.conv`, ReLU under `.activate`."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
        self.activate = nn.ReLU(inplace=True)


class BaseClassifierHead(nn.Module):
    def __init__(self, in_channels: int, num_classes: int) -> None:
        super().__init__()
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.classifier = nn.Conv2d(in_channels, num_classes, kernel_size=1)
        self._packed = None

    def _version(self):
        return sum(p._version for p in self.parameters())

    def _features_bf16(self, x):
        """[B,C,H,W] float (any layout) or bf16 channels-last -> dense NHWC bf16 [B,H,W,C8]."""
        B, C, H, W = x.shape
        if x.dtype == torch.bfloat16 and x.permute(0, 2, 3, 1).is_contiguous() and C % 8 == 0:
            return x.permute(0, 2, 3, 1)
        xn = to_nhwc_f32(x.detach())
        Cp = tc.round_up(C, 8)
        out = torch.empty(B, H, W, Cp, dtype=torch.bfloat16, device=x.device)
        if C % 4 == 0:
            # same-size align_corners bilinear == exact copy; reused as the f32 -> bf16 NHWC converter
            _call("isp_bilinear_ac_nhwc", xn, out, B, C, H, W, H, W, 1, Cp)
        else:
            out.zero_()
            out[..., :C] = xn.to(torch.bfloat16)
        return out

    def _classify(self, feats_bf16, C):
        B, H, W, ld = feats_bf16.shape
        K = self.num_classes
        P = self._packed
        out = torch.empty(B, H, W, K, dtype=torch.float32, device=feats_bf16.device)
        _call("isp_rowdot", feats_bf16, 1, ld, P["wc"], P["bc"], out, B * H * W, C, K)
        return out.permute(0, 3, 1, 2)


class _ConvHeadFn(torch.autograd.Function):
    """ConvSegHead forward + backward on libisp_b200 (the head is the trainable part of the IS model,
    core/training/trainer.py:213-221).  Forward keeps the bf16 NHWC activations of every layer;
    backward: fused classifier backward + ReLU mask -> per layer { wgrad (tcgen05, pixel-major GEMM),
    bias column sums, dgrad (the forward kernel on flipped weights with a ReLU-mask epilogue) }.
    Parameter gradients are fp32; activation gradients travel in bf16."""

    @staticmethod
    def forward(ctx, head, x, *params):
        C, L = head.in_channels, head.num_layers
        P = head._pack(x.device)
        f = head._features_bf16(x)
        acts = [f]
        for w, b in zip(P["w"], P["b"]):
            f = tc.conv3x3(f, w, b, C, C, act="relu", ldy=tc.round_up(C, 8))
            acts.append(f)
        ctx.head, ctx.acts, ctx.in_shape = head, acts, tuple(x.shape)
        ctx.need_dx = x.requires_grad
        return head._classify(f, C)

    @staticmethod
    def backward(ctx, dlogits):
        head, acts = ctx.head, ctx.acts
        C, L = head.in_channels, head.num_layers
        P = head._pack(dlogits.device)
        B, H, W, ld = acts[-1].shape
        M = B * H * W
        dev = dlogits.device
        dlog = dlogits.detach().float().contiguous().view(M)
        dz = torch.empty(B, H, W, ld, dtype=torch.bfloat16, device=dev)
        if ld > C:
            dz.zero_()
        dwc = torch.zeros(C, dtype=torch.float32, device=dev)
        dbc = torch.zeros(1, dtype=torch.float32, device=dev)
        dbias = [torch.zeros(C, dtype=torch.float32, device=dev) for _ in range(L)]
        _call("isp_head_classifier_bwd", acts[-1], ld, dlog, P["wc"], dz, ld, dwc, dbc, dbias[L - 1], M, C)
        dweights = [None] * L
        dx = None
        for i in range(L - 1, -1, -1):
            xin = acts[i]  # input of layer i (bf16 NHWC); for i > 0 also the ReLU output masking dgrad
            dW = torch.zeros(C, 9, C, dtype=torch.float32, device=dev)
            _call("isp_conv3x3_wgrad_bf16_tc", xin, xin.shape[3], dz, ld, dW, B, H, W, C, C)
            dweights[i] = dW.view(C, 3, 3, C).permute(0, 3, 1, 2)
            if i < L - 1:
                _call("isp_colsum_bf16", dz, ld, dbias[i], M, C)
            if i > 0:
                dprev = torch.empty(B, H, W, ld, dtype=torch.bfloat16, device=dev)
                _call("isp_conv3x3_dgrad_bf16_tc", dz, P["wT"][i], xin, xin.shape[3], dprev, 1, B, H, W, C, ld, C, ld)
                dz = dprev
            elif ctx.need_dx:
                dxn = torch.empty(B, H, W, C, dtype=torch.float32, device=dev)
                _call("isp_conv3x3_dgrad_bf16_tc", dz, P["wT"][0], None, 0, dxn, 0, B, H, W, C, ld, C, C)
                dx = dxn.permute(0, 3, 1, 2)
        grads = []
        for i in range(L):
            grads += [dweights[i], dbias[i]]
        grads += [dwc.view(1, C, 1, 1), dbc]
        return (None, dx, *grads)


class ConvSegHead(BaseClassifierHead):
    """Several 3x3 conv+ReLU layers, then a 1x1 classifier (conv_heads.py:48-73).  Differentiable
    (parameters and input) for the 3x3 / one-class configuration every shipped model uses."""

    KERNEL = 3

    def __init__(self, in_channels: int, num_layers: int, num_classes: int) -> None:
        super().__init__(in_channels, num_classes)
        self.num_layers = num_layers
        k = self.KERNEL
        self.convs = nn.Sequential(*[_ConvModule(in_channels, in_channels, k, 1, k // 2) for _ in range(num_layers)])

    def _pack(self, dev):
        key = (str(dev), self._version())
        if self._packed is None or self._packed["key"] != key:
            C = self.in_channels
            P = {"key": key, "w": [], "b": []}
            for m in self.convs:
                w = m.conv.weight.detach().float()
                if self.KERNEL == 3:
                    P["w"].append(tc.pack_conv3x3_weight(w).to(dev))
                    # dgrad operand: W'[ci][tap][co] = W[co][ci][8 - tap]
                    P.setdefault("wT", []).append(tc.pack_conv3x3_weight(w.flip(2, 3).transpose(0, 1)).to(dev))
                else:
                    P["w"].append(tc.pack_linear_weight(w.reshape(C, C)).to(dev))
                P["b"].append(m.conv.bias.detach().float().contiguous().to(dev))
            P["wc"] = self.classifier.weight.detach().float().reshape(self.num_classes, C).contiguous().to(dev)
            P["bc"] = self.classifier.bias.detach().float().contiguous().to(dev)
            self._packed = P
        return self._packed

    def _param_list(self):
        ps = []
        for m in self.convs:
            ps += [m.conv.weight, m.conv.bias]
        return ps + [self.classifier.weight, self.classifier.bias]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            if self.KERNEL != 3 or self.num_classes != 1 or self.in_channels % 64 != 0:
                raise NotImplementedError("head backward is implemented for 3x3 layers, one class and "
                                          "in_channels % 64 == 0 (the shipped ConvSegHead configurations)")
            return _ConvHeadFn.apply(self, x, *self._param_list())
        C = self.in_channels
        P = self._pack(x.device)
        f = self._features_bf16(x)
        B, H, W, ld = f.shape
        for w, b in zip(P["w"], P["b"]):
            if self.KERNEL == 3:
                f = tc.conv3x3(f, w, b, C, C, act="relu", ldy=tc.round_up(C, 8))
            else:
                f = tc.gemm(f.view(B * H * W, -1), w, bias=b, act="relu", N=C, K=C,
                            ldd=tc.round_up(C, 8)).view(B, H, W, -1)
        return self._classify(f, C)


class SimpleConvSegHead(ConvSegHead):
    """Several 1x1 conv+ReLU layers, then the classifier (conv_heads.py:22-45)."""

    KERNEL = 1


class SimpleClassifierHead(BaseClassifierHead):
    """Single 1x1 conv (conv_heads.py:10-19)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dev = x.device
        key = (str(dev), self._version())
        if self._packed is None or self._packed["key"] != key:
            C = self.in_channels
            self._packed = {"key": key,
                            "wc": self.classifier.weight.detach().float().reshape(self.num_classes, C).contiguous().to(dev),
                            "bc": self.classifier.bias.detach().float().contiguous().to(dev)}
        return self._classify(self._features_bf16(x), self.in_channels)


HEAD_REGISTRY = {
    "linear": SimpleClassifierHead,
    "simple_conv": SimpleConvSegHead,
    "convhead": ConvSegHead,
}
