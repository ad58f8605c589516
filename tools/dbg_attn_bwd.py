"""Stage-by-stage check of the flash attention backward (prints after every kernel so a stall is attributable)."""
import sys
import torch
from isegprobe_b200 import _lib

DEV = "cuda:0"
B, nh, rows, T, hd = [int(x) for x in sys.argv[1:6]]
need_dq = int(sys.argv[6])
HP = 112 if hd > 64 else 64
variant, DKC = (1, 128) if hd > 64 else (0, 64)
g = torch.Generator().manual_seed(5)


def call(name, *a):
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


def rnd(*shape, s=1.0):
    t = torch.zeros(*shape[:-1], HP)
    t[..., :hd] = torch.randn(*shape[:-1], hd, generator=g) * s
    return t.to(torch.bfloat16)


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm()).clamp(min=1e-30))


Q = rnd(B, rows, nh, HP, s=0.35).reshape(B * rows, nh * HP).to(DEV)
dO = rnd(B, rows, nh, HP, s=0.5).reshape(B * rows, nh * HP).to(DEV)
K = rnd(B, nh, T, HP).to(DEV)
V = rnd(B, nh, T, HP).to(DEV)
Tp = (T + 127) // 128 * 128
Kp = torch.zeros(B, nh, Tp, DKC, dtype=torch.bfloat16, device=DEV)
Kp[:, :, :T, :HP] = K
Vt = torch.zeros(B, nh, HP, Tp, dtype=torch.bfloat16, device=DEV)
Vt[:, :, :, :T] = V.transpose(2, 3)
O = torch.empty(B * rows, nh * HP, dtype=torch.bfloat16, device=DEV)
lse = torch.zeros(B * nh * rows + 64, device=DEV)
call("isp_attention_bf16_tc_lse", Q, nh * HP, HP, Kp, Vt, O, nh * HP, HP, B, rows, nh, T, variant, lse)
torch.cuda.synchronize()
print("forward + lse ok", flush=True)
q = Q.float().view(B, rows, nh, HP).permute(0, 2, 1, 3).requires_grad_(True)
k = K.float().requires_grad_(True)
v = V.float().requires_grad_(True)
s = q @ k.transpose(-1, -2)
o = torch.softmax(s, -1) @ v
want_lse = torch.logsumexp(s.detach(), -1) * 1.4426950408889634
print("lse max err", float((lse[:B * nh * rows].view(B, nh, rows) - want_lse).abs().max()), flush=True)
do = dO.float().view(B, rows, nh, HP).permute(0, 2, 1, 3)
(o * do).sum().backward()
dvec = torch.zeros(B * nh * rows + 64, device=DEV)
call("isp_attention_rowdot_heads", dO, nh * HP, O, nh * HP, dvec, B, rows, nh, HP)
torch.cuda.synchronize()
print("rowdot ok, err", float((dvec[:B * nh * rows].view(B, nh, rows) - (do * o.detach()).sum(-1)).abs().max()), flush=True)
for dq_on in ([0, 1] if need_dq else [0]):
    dK = torch.zeros(B, nh, T, HP, device=DEV)
    dV = torch.zeros(B, nh, T, HP, device=DEV)
    dQ = torch.zeros(B * rows, nh * HP, device=DEV) if dq_on else None
    call("isp_attention_bwd_bf16_tc", Q, nh * HP, dO, nh * HP, K, V, lse, dvec, dK, dV, dQ, nh * HP, B, rows, nh, T, HP)
    torch.cuda.synchronize()
    print("bwd dq=%d ok: cos dV %.5f dK %.5f  norms %.3e/%.3e %.3e/%.3e" % (
        dq_on, cos(dV, v.grad), cos(dK, k.grad), float(dV.norm()), float(v.grad.norm()), float(dK.norm()),
        float(k.grad.norm())), flush=True)
    if dq_on:
        got = dQ.view(B, rows, nh, HP).permute(0, 2, 1, 3)
        print("   cos dQ %.5f norms %.3e/%.3e" % (cos(got, q.grad), float(got.norm()), float(q.grad.norm())), flush=True)
