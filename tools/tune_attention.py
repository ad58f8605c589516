"""Times the LoftUp cross-attention kernel variants on chunk-shaped operands (4 images x 200704 queries x 1024 keys,
4 heads x 101 in 112): plain entry vs isp_attention_bf16_tc_opt (ones-column row sum, polynomial exp2 share).
    python tools/tune_attention.py [images] [iters]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from isegprobe_b200 import _lib  # noqa: E402

images = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
nh, hd, HP, KP, T, HW = 4, 101, 112, 128, 1024, 448 * 448
bf = torch.bfloat16
torch.manual_seed(0)
Q = torch.zeros(images * HW, nh * HP, device=dev, dtype=bf)
Q.view(-1, nh, HP)[:, :, :hd] = (torch.randn(images * HW, nh, hd, device=dev) * hd ** -0.25).to(bf)
Kp = torch.zeros(images, nh, T, KP, device=dev, dtype=bf)
Kp[..., :hd] = (torch.randn(images, nh, T, hd, device=dev) * hd ** -0.25).to(bf)
Vt = torch.zeros(images, nh, HP, T, device=dev, dtype=bf)
Vt[:, :, :hd] = torch.randn(images, nh, hd, T, device=dev).to(bf)
Vt1 = Vt.clone()
Vt1[:, :, hd] = 1.0
O = torch.empty(images * HW, nh * HP, dtype=bf, device=dev)
st = torch.cuda.current_stream().cuda_stream
flops = 2.0 * 2 * nh * HW * T * hd * images


def run(lsum, poly):
    if lsum:
        _lib.call("isp_attention_bf16_tc_opt", Q.data_ptr(), nh * HP, HP, Kp.data_ptr(), Vt1.data_ptr(), O.data_ptr(), nh * HP,
                  HP, images, HW, nh, T, 1, None, hd, poly, st)
    else:
        _lib.call("isp_attention_bf16_tc", Q.data_ptr(), nh * HP, HP, Kp.data_ptr(), Vt.data_ptr(), O.data_ptr(), nh * HP, HP,
                  images, HW, nh, T, 3 if lsum is None else 1, st)


ref = None
for lsum, poly in [(None, 0), (False, 0), (True, 0), (True, 2), (True, 3), (True, 4)]:  # None = one-tile kernel
    for _ in range(3):
        run(lsum, poly)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run(lsum, poly)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    o = O.view(-1, nh, HP)[:4096, :, :hd].float().clone()
    if ref is None:
        ref = o
    print(json.dumps({"kernel": "one-tile" if lsum is None else "two-tile", "lsum": lsum, "poly": poly, "ms": round(ms, 4), "tflops_unpadded": round(flops / ms / 1e9, 1),
                      "max_abs_diff_vs_plain": float((o - ref).abs().max()), "ref_absmax": float(ref.abs().max())}))
