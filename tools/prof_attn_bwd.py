"""Flash attention backward at the LoftUp shape (one image: 200 704 queries x 1024 keys, 4 heads x 101): CUDA-event
times with and without dQ.  ONCE=1: a single dQ launch (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isegprobe_b200 import _lib

DEV = "cuda:0"
B, nh, rows, T, hd, HP = int(os.environ.get("BATCH", "1")), 4, 448 * 448, 1024, 101, 112
g = torch.Generator(device=DEV).manual_seed(0)


def call(name, *a):
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


def rnd(*shape, s=1.0):
    t = torch.zeros(*shape, device=DEV)
    t[..., :hd] = torch.randn(*shape[:-1], hd, generator=g, device=DEV) * s
    return t.to(torch.bfloat16)


Q = rnd(B, rows, nh, HP, s=0.35).reshape(B * rows, nh * HP)
dO = rnd(B, rows, nh, HP, s=0.5).reshape(B * rows, nh * HP)
K, V = rnd(B, nh, T, HP), rnd(B, nh, T, HP)
Kp = torch.zeros(B, nh, T, 128, dtype=torch.bfloat16, device=DEV)
Kp[..., :HP] = K
Vt = V.transpose(2, 3).contiguous()
O = torch.empty_like(Q)
lse = torch.zeros(B * nh * rows + 64, device=DEV)
dvec = torch.zeros(B * nh * rows + 64, device=DEV)
call("isp_attention_bf16_tc_lse", Q, nh * HP, HP, Kp, Vt, O, nh * HP, HP, B, rows, nh, T, 1, lse)
call("isp_attention_rowdot_heads", dO, nh * HP, O, nh * HP, dvec, B, rows, nh, HP)
dK, dV = torch.zeros(B, nh, T, HP, device=DEV), torch.zeros(B, nh, T, HP, device=DEV)
dQ = torch.zeros(B * rows, nh * HP, device=DEV)


def run(dq):
    call("isp_attention_bwd_bf16_tc", Q, nh * HP, dO, nh * HP, K, V, lse, dvec, dK, dV, dQ if dq else None, nh * HP, B, rows,
         nh, T, HP)


if os.environ.get("ONCE"):
    run(int(os.environ.get("DQ", "1")))
    torch.cuda.synchronize()
    sys.exit(0)
flops = 2.0 * B * nh * rows * T * hd
for dq in (0, 1):
    for _ in range(3):
        run(dq)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run(dq)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("attention_bwd B=%d dq=%d: %.3f ms  %.0f TFLOP/s (%d GEMMs, un-padded)" % (B, dq, ms, (4 + dq) * flops / ms / 1e9, 4 + dq))
