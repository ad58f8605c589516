"""Host-side helpers around the tcgen05 GEMM / implicit-GEMM conv entry points:
weight packing (done once, at module construction) and thin call wrappers."""
import torch

from . import _lib

ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2, "quick_gelu": 3, "gelu_tanh": 4}


def round_up(x, m):
    return (x + m - 1) // m * m


def pack_linear_weight(w: torch.Tensor) -> torch.Tensor:
    """[N, K] (any float dtype) -> bf16 [N, K8], K8 = K rounded up to 8 (TMA 16-byte row stride)."""
    N, K = w.shape
    out = torch.zeros(N, round_up(K, 8), dtype=torch.bfloat16, device=w.device)
    out[:, :K] = w.detach().to(torch.bfloat16)
    return out


def pack_conv3x3_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> bf16 [Cout, 9 * Cin_pad], k = (r*3+s)*Cin_pad + c, Cin_pad = ceil64(Cin)."""
    Cout, Cin = w.shape[:2]
    cp = round_up(Cin, 64)
    out = torch.zeros(Cout, 9, cp, dtype=torch.bfloat16, device=w.device)
    out[:, :, :Cin] = w.detach().permute(0, 2, 3, 1).reshape(Cout, 9, Cin).to(torch.bfloat16)
    return out.reshape(Cout, 9 * cp).contiguous()


def stats_slots(N, out_dtype=torch.bfloat16, has_resid=False):
    """Slots per row of the (sum, sum of squares) buffer a GEMM / conv with N output columns writes."""
    return _lib.lib().isp_gemm_stats_slots(int(N), int(out_dtype == torch.bfloat16), int(has_resid))


def pack_ln_linear(W, b, gamma, beta):
    """Fold LayerNorm(gamma, beta) into the Linear that follows it (see isp_gemm_bf16_tc_ex):
    returns (bf16 W*gamma [N,K8], g[n] = sum_k bf16(W*gamma)[n,k], bias' = W beta + b), fp32 on W's device."""
    W = W.detach().float()
    Wg = pack_linear_weight(W * gamma.detach().float().to(W.device)[None, :])
    g = Wg.float().sum(dim=1).contiguous()
    bias = W @ beta.detach().float().to(W.device)
    if b is not None:
        bias = bias + b.detach().float().to(W.device)
    return Wg, g, bias.contiguous()


def gemm(A, W, bias=None, resid=None, alpha=1.0, act=None, out_dtype=torch.bfloat16, N=None, K=None, ldd=None,
         out=None, ln_stats=None, ln_g=None, ln_eps=1e-5, stats_out=None):
    """D[M,N] = alpha*act(A[M,:K] @ W[:N,:K]^T + bias) + resid.  A, W bf16 2-D with 8-aligned row strides.
    ln_stats / ln_g: LayerNorm of the A rows fused in (W, bias, ln_g from pack_ln_linear; ln_stats the
    [M, slots, 2] statistics a previous call produced with stats_out=)."""
    assert A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16 and A.dim() == 2 and W.dim() == 2
    assert A.stride(1) == 1 and W.stride(1) == 1
    M = A.shape[0]
    N = W.shape[0] if N is None else N
    K = min(A.shape[1], W.shape[1]) if K is None else K
    if out is None:
        ldd = N if ldd is None else ldd
        out = torch.empty(M, ldd, dtype=out_dtype, device=A.device)
    assert out.stride(1) == 1
    rb, ldr = 0, 0
    if resid is not None:
        assert resid.dim() == 2 and resid.stride(1) == 1
        rb, ldr = int(resid.dtype == torch.bfloat16), resid.stride(0)
        assert resid.dtype in (torch.bfloat16, torch.float32)
    if bias is not None:
        assert bias.dtype == torch.float32
    if ln_stats is None and stats_out is None:
        _lib.call("isp_gemm_bf16_tc", _lib.dptr(A), A.stride(0), _lib.dptr(W), W.stride(0), _lib.dptr(bias),
                  _lib.dptr(resid), rb, ldr, float(alpha), ACT[act], _lib.dptr(out), out.stride(0),
                  int(out.dtype == torch.bfloat16), M, N, K, _lib.stream_ptr())
        return out
    for t in (ln_stats, stats_out):
        assert t is None or (t.dtype == torch.float32 and t.dim() == 3 and t.shape[0] == M and t.shape[2] == 2
                             and t.is_contiguous())
    _lib.call("isp_gemm_bf16_tc_ex", _lib.dptr(A), A.stride(0), _lib.dptr(W), W.stride(0), _lib.dptr(bias),
              _lib.dptr(resid), rb, ldr, float(alpha), ACT[act], _lib.dptr(out), out.stride(0),
              int(out.dtype == torch.bfloat16), M, N, K, _lib.dptr(ln_stats), 0 if ln_stats is None else ln_stats.shape[1],
              _lib.dptr(ln_g), float(ln_eps), _lib.dptr(stats_out), 0 if stats_out is None else stats_out.shape[1],
              _lib.stream_ptr())
    return out


def conv3x3(x, w_packed, bias, cin, cout, act="relu", out_dtype=torch.bfloat16, ldy=None, stats_out=None):
    """x: NHWC bf16 [B,H,W,ldx] (first `cin` channels real) -> NHWC [B,H,W,ldy].  stats_out: [B*H*W, slots, 2]
    row statistics of the output for a LayerNorm fused into the next GEMM."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    B, H, W, ldx = x.shape
    ldy = round_up(cout, 8) if ldy is None else ldy
    y = torch.empty(B, H, W, ldy, dtype=out_dtype, device=x.device)
    if stats_out is None:
        _lib.call("isp_conv3x3_bf16_tc", _lib.dptr(x), _lib.dptr(w_packed), _lib.dptr(bias), ACT[act], _lib.dptr(y),
                  int(out_dtype == torch.bfloat16), B, H, W, cin, ldx, cout, ldy, _lib.stream_ptr())
    else:
        assert stats_out.dtype == torch.float32 and stats_out.is_contiguous() and stats_out.shape[0] == B * H * W
        _lib.call("isp_conv3x3_bf16_tc_ex", _lib.dptr(x), _lib.dptr(w_packed), _lib.dptr(bias), ACT[act], _lib.dptr(y),
                  int(out_dtype == torch.bfloat16), B, H, W, cin, ldx, cout, ldy, _lib.dptr(stats_out),
                  stats_out.shape[1], _lib.stream_ptr())
    return y


def ffn_fused(x, W1g, g1, b1, W2, b2, K1, N2, ln_stats, ln_eps=1e-5, ldo=None, want_stats=True):
    """out = x + Linear2(GELU(Linear1(LayerNorm(x)))) in one kernel (isp_ffn_fused_bf16_tc).  x bf16 [M, ldx]; W1g / g1 / b1
    from pack_ln_linear; W2 bf16 [N2, K8]; ln_stats [M, slots, 2].  Returns (out bf16 [M, ldo], stats [M, 4, 2] or None)."""
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    M = x.shape[0]
    NH = W1g.shape[0]
    ldo = round_up(N2, 16) if ldo is None else ldo
    out = torch.empty(M, ldo, dtype=torch.bfloat16, device=x.device)
    stats = torch.empty(M, 4, 2, dtype=torch.float32, device=x.device) if want_stats else None
    _lib.call("isp_ffn_fused_bf16_tc", _lib.dptr(x), x.stride(0), int(K1), _lib.dptr(W1g), W1g.stride(0), _lib.dptr(g1),
              _lib.dptr(b1), int(NH), _lib.dptr(W2), W2.stride(0), _lib.dptr(b2), int(N2), _lib.dptr(out), ldo, M,
              _lib.dptr(ln_stats), ln_stats.shape[1], float(ln_eps), _lib.dptr(stats), _lib.stream_ptr())
    return out, stats
