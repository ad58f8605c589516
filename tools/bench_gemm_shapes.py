"""CUDA-event timings of the tensor-core kernels at the shapes the LoftUp / ViT / JBU / head paths
use (4-image chunk at 448^2 unless stated).  Prints TFLOP/s on un-padded flops and the
memory-bound floor next to each.  Diagnostic; bench.py is the contract."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isegprobe_b200 import _lib, tc  # noqa: E402

dev = "cuda:0"
bf, f32 = torch.bfloat16, torch.float32
NI = int(os.environ.get("NI", 4))
M = NI * 448 * 448
which = os.environ.get("WHICH", "gemm,conv,ln,attn").split(",")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5):
    for _ in range(2):
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


rows = []


def report(name, ms, flops, bytes_):
    rows.append({"name": name, "ms": round(ms, 4), "TFLOPs": round(flops / ms / 1e9, 1),
                 "GBs": round(bytes_ / ms / 1e6, 1), "mem_floor_ms": round(bytes_ / 6545e6, 4),
                 "mma_floor_ms": round(flops / 1355e9, 4)})
    print(rows[-1], flush=True)


def gemm_case(name, m, n, k, lda, ldd, act=None, resid=False, out_dtype=bf, bias=True):
    A = (torch.randn(m, lda, device=dev) * 0.5).to(bf)
    W = tc.pack_linear_weight(torch.randn(n, k, device=dev) * 0.05)
    b = torch.randn(n, device=dev) if bias else None
    R = (torch.randn(m, ldd, device=dev)).to(out_dtype) if resid else None
    out = torch.empty(m, ldd, dtype=out_dtype, device=dev)
    ms = timeit(lambda: tc.gemm(A, W, bias=b, resid=R, act=act, out_dtype=out_dtype, N=n, K=k, out=out))
    esz = 2 if out_dtype == bf else 4
    by = m * k * 2 + m * n * esz * (2 if resid else 1)
    report(name, ms, 2.0 * m * n * k, by)


if "gemm" in which:
    gemm_case("loftup q-proj 404->448 bf16", M, 448, 404, 416, 448)
    gemm_case("loftup out-proj 448->404 +resid bf16", M, 404, 448, 448, 416, resid=True)
    gemm_case("loftup ffn1 404->384 gelu_tanh bf16", M, 384, 404, 416, 384, act="gelu_tanh")
    gemm_case("loftup ffn1 404->384 gelu(erf) bf16", M, 384, 404, 416, 384, act="gelu")
    gemm_case("loftup ffn2 384->404 +resid bf16", M, 404, 384, 384, 416, resid=True)
    gemm_case("loftup final 404->384 f32", M, 384, 404, 416, 384, out_dtype=f32)
    gemm_case("plain 384->384 bf16 no-bias", M, 384, 384, 384, 384, bias=False)
    m_vit = 16 * 1025
    gemm_case("vit qkv 384->1152 bf16 (B=16)", m_vit, 1152, 384, 384, 1152)
    gemm_case("vit proj 384->384 +resid f32 (B=16)", m_vit, 384, 384, 384, 384, resid=True, out_dtype=f32)
    gemm_case("vit fc1 384->1536 gelu_tanh (B=16)", m_vit, 1536, 384, 384, 1536, act="gelu_tanh")
    gemm_case("vit fc2 1536->384 +resid f32 (B=16)", m_vit, 384, 1536, 1536, 384, resid=True, out_dtype=f32)
    mj = 4 * 448 * 448
    gemm_case("jbu final 1x1 384->384 +resid f32 (4 img @448)", mj, 384, 384, 384, 384, resid=True, out_dtype=f32)

if "conv" in which:
    for name, cin, ldx, cout, ldy, ni in (("loftup conv1 203->404", 203, 208, 404, 416, NI),
                                          ("loftup conv2 404->404", 404, 416, 404, 416, NI),
                                          ("head conv 384->384", 384, 384, 384, 384, NI)):
        x = torch.randn(ni, 448, 448, ldx, device=dev).to(bf)
        w = tc.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device=dev) * 0.02)
        b = torch.zeros(cout, device=dev)
        ms = timeit(lambda: tc.conv3x3(x, w, b, cin, cout, act="relu", ldy=ldy), 3)
        report(name, ms, 2.0 * ni * 448 * 448 * cout * cin * 9, ni * 448 * 448 * (ldx + ldy) * 2)
        del x

if "ln" in which:
    A = torch.randn(M, 416, device=dev).to(bf)
    g = torch.ones(404, device=dev)
    out = torch.empty(M, 416, dtype=bf, device=dev)
    ms = timeit(lambda: call("isp_layernorm_rows", A, 1, 416, out, 1, 416, g, g, M, 404, 1e-5))
    report("layernorm 404 bf16->bf16", ms, 0, M * 416 * 4)
    A32 = torch.randn(M, 384, device=dev)
    o32 = torch.empty(M, 384, device=dev)
    g = torch.ones(384, device=dev)
    ms = timeit(lambda: call("isp_layernorm_rows", A32, 0, 384, o32, 0, 384, g, g, M, 384, 1e-6))
    report("layernorm 384 f32->f32", ms, 0, M * 384 * 8)
    del A, A32, o32, out

if "attn" in which:
    nh, T = 4, 1024
    Q = (torch.randn(M, 448, device=dev) * 0.3).to(bf)
    K = (torch.randn(NI, nh, 1024, 128, device=dev) * 0.3).to(bf)
    K[..., 101:] = 0
    Vt = torch.randn(NI, nh, 112, 1024, device=dev).to(bf)
    Vt[:, :, 101:] = 0
    O = torch.empty(M, 448, dtype=bf, device=dev)
    ms = timeit(lambda: call("isp_attention_bf16_tc", Q, 448, 112, K, Vt, O, 448, 112, NI, 448 * 448, nh, T, 1), 3)
    report("loftup attention (4 heads x 101, 1024 keys)", ms, 4.0 * M * nh * T * 101, M * 448 * 4)

json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                                  "gemm_shapes.json"), "w"), indent=1)
