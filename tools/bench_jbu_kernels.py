"""Per-kernel CUDA-event timings of the JBU stage kernels at BASELINE sizes
(B=16, 448^2 guidance, 384 channels).  Diagnostic; bench.py is the contract."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isegprobe_b200 import _lib  # noqa: E402

dev = "cuda:0"
B = int(os.environ.get("B", 16))
C = 384


def timeit(fn, iters=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def call(name, *a):
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


res = {}
gd = torch.rand(B, 3, 448, 448, device=dev)
for GH in (64, 128, 256, 512):
    h = GH // 2
    src = torch.randn(B, h, h, C, device=dev)
    g = torch.empty(B, GH, GH, 4, device=dev)
    proj = torch.empty(B, GH, GH, 32, device=dev)
    filt = torch.rand(B, GH, GH, 49, device=dev)
    filt56 = torch.empty(B, GH, GH, 56, device=dev)
    hr = torch.empty(B, GH + 6, GH + 6, C, device=dev)
    out = torch.empty(B, GH, GH, C, device=dev)
    w0, b0 = torch.randn(32, 3, device=dev), torch.randn(32, device=dev)
    w1, b1 = torch.randn(32, 32, device=dev) * 0.2, torch.randn(32, device=dev)
    f0, fb0 = torch.randn(49, 52, device=dev) * 0.1, torch.randn(49, device=dev)
    f1, fb1 = torch.randn(49, 49, device=dev) * 0.1, torch.randn(49, device=dev)
    res[f"pool_{GH}"] = timeit(lambda: call("isp_jbu_pool_guidance", gd, g, B, 448, 448, GH, GH, *gd.stride()))
    res[f"range_proj_{GH}"] = timeit(lambda: call("isp_jbu_range_proj", g, proj, B * GH * GH, w0, b0, w1, b1))
    res[f"filters_{GH}"] = timeit(lambda: call("isp_jbu_filters", proj, g, filt56, B, GH, GH, 1.0, 1.0, f0, fb0, f1, fb1, 56))
    res[f"bicubic_pad_{GH}"] = timeit(lambda: call("isp_jbu_bicubic2x_reflectpad", src, hr, B, h, h, C))
    t = timeit(lambda: call("isp_adaptive_conv_fwd", hr, filt56, out, B, GH, GH, C, 56))
    res[f"adaptive_conv_{GH}"] = t
    op_bytes = 4 * B * (C * (GH + 6) ** 2 + 49 * GH * GH + C * GH * GH)
    res[f"adaptive_conv_{GH}_GBs"] = op_bytes / t / 1e6
    t1 = float("nan")  # the first-generation kernel was removed in round 2
    res[f"adaptive_conv_v1_{GH}"] = t1
    if GH == 512:
        wf, bf = torch.randn(C, C, device=dev) * 0.05, torch.randn(C, device=dev)
        o2 = torch.empty_like(out)
        res["final_gemm_simt_512"] = timeit(lambda: call("isp_gemm_f32_simt", out, wf, bf, out, 0.1, o2, B * GH * GH, C, C), 3)
        o3 = torch.empty(B, 448, 448, C, device=dev)
        res["bilinear_512_448"] = timeit(lambda: call("isp_bilinear_ac_nhwc", o2, o3, B, C, 512, 512, 448, 448, 0, C))
    del src, g, proj, filt, hr, out
print(json.dumps({k: round(v, 4) for k, v in res.items()}, indent=1))
