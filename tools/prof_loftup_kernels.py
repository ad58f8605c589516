"""A few representative LoftUp-path launches at BASELINE size (one 448x448 image) for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isegprobe_b200 import _lib, tc
dev = "cuda"
bf = torch.bfloat16
M, T, nh = 448 * 448, 1024, 4
which = os.environ.get("WHICH", "attn,conv,gemm,ln").split(",")
def call(name, *a):
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())
if "attn" in which:
    Q = (torch.randn(M, 448, device=dev) * 0.3).to(bf)
    K = (torch.randn(1, nh, 1024, 128, device=dev) * 0.3).to(bf); K[..., 101:] = 0
    Vt = torch.randn(1, nh, 112, 1024, device=dev).to(bf); Vt[:, :, 101:] = 0
    O = torch.empty(M, 448, dtype=bf, device=dev)
    for _ in range(2):
        call("isp_attention_bf16_tc", Q, 448, 112, K, Vt, O, 448, 112, 1, M, nh, T, 1)
if "conv" in which:
    x = torch.randn(1, 448, 448, 416, device=dev).to(bf)
    w = tc.pack_conv3x3_weight(torch.randn(404, 404, 3, 3) * 0.02).to(dev)
    b = torch.zeros(404, device=dev)
    for _ in range(2):
        y = tc.conv3x3(x, w, b, 404, 404, act="relu", ldy=416)
if "gemm" in which:
    A = torch.randn(M, 416, device=dev).to(bf)
    W = tc.pack_linear_weight(torch.randn(404, 404) * 0.05).to(dev)
    bias = torch.zeros(404, device=dev)
    R = torch.randn(M, 416, device=dev).to(bf)
    for _ in range(2):
        y = tc.gemm(A, W, bias=bias, resid=R, out_dtype=bf, N=404, K=404, ldd=416)
if "ln" in which:
    A = torch.randn(M, 416, device=dev).to(bf)
    g = torch.ones(404, device=dev)
    out = torch.empty(M, 416, dtype=bf, device=dev)
    for _ in range(2):
        call("isp_layernorm_rows", A, 1, 416, out, 1, 416, g, g, M, 404, 1e-5)
torch.cuda.synchronize()
print("ok")
