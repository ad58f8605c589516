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


class ConvSegHead(BaseClassifierHead):
    """Several 3x3 conv+ReLU layers, then a 1x1 classifier (conv_heads.py:48-73)."""

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
                else:
                    P["w"].append(tc.pack_linear_weight(w.reshape(C, C)).to(dev))
                P["b"].append(m.conv.bias.detach().float().contiguous().to(dev))
            P["wc"] = self.classifier.weight.detach().float().reshape(self.num_classes, C).contiguous().to(dev)
            P["bc"] = self.classifier.bias.detach().float().contiguous().to(dev)
            self._packed = P
        return self._packed

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("head backward is not implemented yet (forward/inference only)")
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
