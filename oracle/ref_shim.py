"""Import shim that makes the reference's hot-path modules importable in the
authoring container (no omegaconf / mmcv / timm here).  Used ONLY by
`oracle/make_golden.py` and by not-gpu tests that skip when /root/reference is
absent.  Nothing on the GPU box may import this (the reference does not travel).
"""
import os
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = os.environ.get("ISEGPROBE_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core", "model"))


def install() -> None:
    """Register namespace stand-ins so `core.model.<x>` imports skip the package
    __init__ files (they pull omegaconf/timm/mmcv)."""
    if "core" in sys.modules and getattr(sys.modules["core"], "_isp_shim", False):
        return
    os.environ.setdefault("XFORMERS_DISABLED", "1")
    for name, rel in [
        ("core", "core"),
        ("core.model", "core/model"),
        ("core.utils", "core/utils"),
        ("core.model.featurizers", "core/model/featurizers"),
    ]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
        m._isp_shim = True
        sys.modules[name] = m
    oc = types.ModuleType("omegaconf")
    oc.DictConfig = dict
    oc.OmegaConf = object
    sys.modules.setdefault("omegaconf", oc)

    class ConvModule(nn.Module):  # mmcv==1.6.2 defaults: conv(bias) + ReLU(inplace)
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
            super().__init__()
            self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
            self.activate = nn.ReLU(inplace=True)

        def forward(self, x):
            return self.activate(self.conv(x))

    mm, mmc = types.ModuleType("mmcv"), types.ModuleType("mmcv.cnn")
    mmc.ConvModule = ConvModule
    mm.cnn = mmc
    sys.modules.setdefault("mmcv", mm)
    sys.modules.setdefault("mmcv.cnn", mmc)
