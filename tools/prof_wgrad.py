"""One isp_conv3x3_wgrad_bf16_tc launch at head size (B images 448x448, C=384) for ncu / timing."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isegprobe_b200 import _lib
dev, B, H, W, C = "cuda:0", int(os.environ.get("B", 4)), 448, 448, 384
x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
dy = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
dW = torch.zeros(C, 9, C, device=dev)
def run():
    _lib.call("isp_conv3x3_wgrad_bf16_tc", _lib.dptr(x), C, _lib.dptr(dy), C, _lib.dptr(dW), B, H, W, C, C, _lib.stream_ptr())
for _ in range(2):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"wgrad B={B}: {ms:.3f} ms, {2*9*C*C*B*H*W/ms/1e9:.1f} TFLOP/s")
