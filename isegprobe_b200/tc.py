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


def gemm(A, W, bias=None, resid=None, alpha=1.0, act=None, out_dtype=torch.bfloat16, N=None, K=None, ldd=None,
         out=None):
    """D[M,N] = alpha*act(A[M,:K] @ W[:N,:K]^T + bias) + resid.  A, W bf16 2-D with 8-aligned row strides."""
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
    _lib.call("isp_gemm_bf16_tc", _lib.dptr(A), A.stride(0), _lib.dptr(W), W.stride(0), _lib.dptr(bias),
              _lib.dptr(resid), rb, ldr, float(alpha), ACT[act], _lib.dptr(out), out.stride(0),
              int(out.dtype == torch.bfloat16), M, N, K, _lib.stream_ptr())
    return out


def conv3x3(x, w_packed, bias, cin, cout, act="relu", out_dtype=torch.bfloat16, ldy=None):
    """x: NHWC bf16 [B,H,W,ldx] (first `cin` channels real) -> NHWC [B,H,W,ldy]."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    B, H, W, ldx = x.shape
    ldy = round_up(cout, 8) if ldy is None else ldy
    y = torch.empty(B, H, W, ldy, dtype=out_dtype, device=x.device)
    _lib.call("isp_conv3x3_bf16_tc", _lib.dptr(x), _lib.dptr(w_packed), _lib.dptr(bias), ACT[act], _lib.dptr(y),
              int(out_dtype == torch.bfloat16), B, H, W, cin, ldx, cout, ldy, _lib.stream_ptr())
    return y
