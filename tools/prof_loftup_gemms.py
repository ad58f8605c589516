"""The four skinny GEMMs of a LoftUp transformer layer at BASELINE size (4 images: M = 802816 rows), for ncu / timing:
q-proj (fused LayerNorm, N=448), out-proj (K=448, residual, row statistics), FF1 (fused LayerNorm, GELU), FF2 (residual)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isegprobe_b200 import tc
dev, bf = "cuda", torch.bfloat16
M = int(float(os.environ.get("IMAGES", "4")) * 448 * 448)
D, C, Dp, NQ = 404, 384, 416, 448
torch.manual_seed(0)
x = torch.randn(M, Dp, device=dev).to(bf); x[:, D:] = 0
O = torch.randn(M, NQ, device=dev).to(bf)
h1 = torch.randn(M, C, device=dev).to(bf)
st_in = torch.empty(M, tc.stats_slots(D, bf, False), 2, device=dev)
st_in[..., 0] = x.float().sum(1, keepdim=True) / st_in.shape[1]
st_in[..., 1] = (x.float() ** 2).sum(1, keepdim=True) / st_in.shape[1]
st_a = torch.empty(M, tc.stats_slots(D, bf, False), 2, device=dev)
st_b = torch.empty(M, tc.stats_slots(D, bf, True), 2, device=dev)
ones, zeros = torch.ones(D), torch.zeros(D)
Wq, gq, bq = [t.to(dev) for t in tc.pack_ln_linear(torch.randn(NQ, D) * 0.05, torch.zeros(NQ), ones, zeros)]
W1, g1, b1 = [t.to(dev) for t in tc.pack_ln_linear(torch.randn(C, D) * 0.05, torch.zeros(C), ones, zeros)]
Wo = tc.pack_linear_weight(torch.randn(D, NQ) * 0.05).to(dev)
W2 = tc.pack_linear_weight(torch.randn(D, C) * 0.05).to(dev)
bo = torch.zeros(D, device=dev)
runs = {
    "q_proj": lambda: tc.gemm(x, Wq, bias=bq, out_dtype=bf, N=NQ, K=D, ln_stats=st_in, ln_g=gq, ln_eps=1e-5),
    "out_proj": lambda: tc.gemm(O, Wo, bias=bo, resid=x, out_dtype=bf, N=D, K=NQ, ldd=Dp, stats_out=st_b),
    "ff1": lambda: tc.gemm(x, W1, bias=b1, act="gelu_tanh", out_dtype=bf, N=C, K=D, ln_stats=st_in, ln_g=g1, ln_eps=1e-5),
    "ff2": lambda: tc.gemm(h1, W2, bias=bo, resid=x, out_dtype=bf, N=D, K=C, ldd=Dp, stats_out=st_a),
}
which = os.environ.get("WHICH", ",".join(runs)).split(",")
for name in which:
    fn = runs[name]
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
