"""One isp_jbu_filters + isp_jbu_bicubic2x_reflectpad + isp_jbu_range_proj launch at stage-512 size (B=4) for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isegprobe_b200 import _lib
dev, B, GH, C = "cuda:0", int(os.environ.get("B", 4)), 512, 384
def call(name, *a):
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())
g = torch.rand(B, GH, GH, 4, device=dev)
proj = torch.randn(B, GH, GH, 32, device=dev)
filt56 = torch.empty(B, GH, GH, 56, device=dev)
f0, fb0 = torch.randn(49, 52, device=dev) * 0.1, torch.randn(49, device=dev)
f1, fb1 = torch.randn(49, 49, device=dev) * 0.1, torch.randn(49, device=dev)
w0, b0 = torch.randn(32, 3, device=dev), torch.randn(32, device=dev)
w1, b1 = torch.randn(32, 32, device=dev) * 0.2, torch.randn(32, device=dev)
src = torch.randn(B, GH // 2, GH // 2, C, device=dev)
hr = torch.empty(B, GH + 6, GH + 6, C, device=dev)
for _ in range(2):
    call("isp_jbu_range_proj", g, proj, B * GH * GH, w0, b0, w1, b1)
    call("isp_jbu_filters", proj, g, filt56, B, GH, GH, 1.0, 1.0, f0, fb0, f1, fb1, 56)
    call("isp_jbu_bicubic2x_reflectpad", src, hr, B, GH // 2, GH // 2, C)
torch.cuda.synchronize()
print("ok")
