import sys, os, math, torch
sys.path.insert(0, os.getcwd())
from isegprobe_b200 import _lib
dev = "cuda"
B, H, W = 16, 512, 512
torch.manual_seed(0)
proj = torch.randn(B, H, W, 32, device=dev) * 0.3
g4 = torch.randn(B, H, W, 4, device=dev)
fw0 = (torch.randn(49, 52, device=dev) * 0.14).contiguous(); fb0 = torch.randn(49, device=dev) * 0.1
fw1 = (torch.randn(49, 49, device=dev) * 0.14).contiguous(); fb1 = torch.randn(49, device=dev) * 0.1
def run(name, out):
    _lib.call(name, _lib.dptr(proj), _lib.dptr(g4), _lib.dptr(out), B, H, W, 1.0, 1.0, _lib.dptr(fw0), _lib.dptr(fb0), _lib.dptr(fw1), _lib.dptr(fb1), 56, _lib.stream_ptr())
a = torch.empty(B, H, W, 56, device=dev); b = torch.empty_like(a)
for name, out in (("isp_jbu_filters_simt", a), ("isp_jbu_filters", b)):
    for _ in range(2): run(name, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run(name, out)
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 5, "ms (range + fixup)")
print("max abs diff", float((a - b).abs().max()), "rel", float((a - b).abs().max() / a.abs().max()))
# range_proj (3 -> 32 -> 32 per pixel) alone
w0 = torch.randn(32, 3, device=dev).contiguous(); b0 = torch.randn(32, device=dev); w1 = (torch.randn(32, 32, device=dev) * 0.2).contiguous(); b1 = torch.randn(32, device=dev)
def rp():
    _lib.call("isp_jbu_range_proj", _lib.dptr(g4), _lib.dptr(proj), B * H * W, _lib.dptr(w0), _lib.dptr(b0), _lib.dptr(w1), _lib.dptr(b1), _lib.stream_ptr())
for _ in range(2): rp()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): rp()
e1.record(); torch.cuda.synchronize()
print("isp_jbu_range_proj", e0.elapsed_time(e1) / 5, "ms")
