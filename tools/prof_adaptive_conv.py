"""One AdaptiveConv launch at BASELINE stage-512 size (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isegprobe_b200 import _lib
B, GH, C = int(os.environ.get("B", 4)), 512, 384
hr = torch.randn(B, GH + 6, GH + 6, C, device="cuda")
filt = torch.rand(B, GH, GH, 56, device="cuda")
out = torch.empty(B, GH, GH, C, device="cuda")
for _ in range(3):
    _lib.call("isp_adaptive_conv_fwd", hr.data_ptr(), filt.data_ptr(), out.data_ptr(), B, GH, GH, C, 56, _lib.stream_ptr())
torch.cuda.synchronize()
print("ok")
