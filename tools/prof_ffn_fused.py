"""LoftUp FeedForward block at BASELINE size (4 images: 802816 rows): fused kernel vs the two GEMM launches."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isegprobe_b200 import tc
dev, bf = "cuda", torch.bfloat16
M = int(os.environ.get("IMAGES", "4")) * 448 * 448
D, C, Dp = 404, 384, 416
torch.manual_seed(0)
x = torch.randn(M, Dp, device=dev).to(bf); x[:, D:] = 0
st = torch.zeros(M, tc.stats_slots(D, bf, True), 2, device=dev)
st[:, 0, 0], st[:, 0, 1] = x.float().sum(1), (x.float() ** 2).sum(1)
ones, zeros = torch.ones(D), torch.zeros(D)
W1g, g1, b1 = [t.to(dev) for t in tc.pack_ln_linear(torch.randn(C, D) * 0.05, torch.zeros(C), ones, zeros)]
W2 = tc.pack_linear_weight(torch.randn(D, C) * 0.05).to(dev)
b2 = torch.zeros(D, device=dev)
st_a = torch.empty(M, tc.stats_slots(D, bf, True), 2, device=dev)
def two():
    h1 = tc.gemm(x, W1g, bias=b1, act="gelu_tanh", out_dtype=bf, N=C, K=D, ln_stats=st, ln_g=g1, ln_eps=1e-5)
    return tc.gemm(h1, W2, bias=b2, resid=x, out_dtype=bf, N=D, K=C, ldd=Dp, stats_out=st_a)
def fused():
    return tc.ffn_fused(x, W1g, g1, b1, W2, b2, D, D, st)[0]
for name, fn in (("two_gemms", two), ("fused", fused)):
    if os.environ.get("WHICH") and os.environ["WHICH"] != name:
        continue
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / 5, 4), "ms")
