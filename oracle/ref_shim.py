"""Import shim that makes the reference's hot-path modules importable in the
authoring container (no omegaconf / mmcv / timm here).  Used ONLY by
`oracle/make_golden.py`, by tests (they skip when no reference tree is found) and by
bench.py's `--impl reference` / context legs.  The tree is /root/reference in the authoring
container; on the GPU box it is the UNMODIFIED copy `__graft_entry__.build()` stages under the
git-ignored `baseline/_ref/` (never committed).  Test infrastructure, not product.
"""
import os
import sys
import types

import torch.nn as nn

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")  # git-ignored copy made by __graft_entry__.build() (travels to the GPU box)


def _pick_root() -> str:
    env = os.environ.get("ISEGPROBE_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/core/model"):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _pick_root()


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
