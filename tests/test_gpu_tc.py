"""GPU parity: tcgen05 GEMM and implicit-GEMM 3x3 conv vs a plain torch fp32 reference of
the same op on the same bf16-rounded operands (tolerance: bf16 output rounding + fp32
accumulation order, 1e-2 relative to max; cosine >= 0.9999)."""
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import DEV, cosine, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tc():
    from isegprobe_b200 import tc
    return tc


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 128, 128), (1000, 404, 404), (4096, 384, 1536),
                                   (333, 1152, 384), (70000, 384, 404), (128, 16, 8)])
def test_gemm_plain(tc, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * K ** -0.5).to(torch.bfloat16)
    Ap, Wp = tc.pack_linear_weight(A).to(DEV), tc.pack_linear_weight(W).to(DEV)
    out = tc.gemm(Ap, Wp, out_dtype=torch.float32, N=N, K=K)
    want = A.float() @ W.float().T
    assert out.shape == (M, N)
    assert relerr(out, want) < 1e-4, relerr(out, want)


def test_gemm_epilogue(tc):
    g = torch.Generator().manual_seed(5)
    M, N, K = 777, 404, 384
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g)
    R = torch.randn(M, N, generator=g)
    for act, fn in (("relu", torch.relu), ("gelu", F.gelu), (None, lambda x: x)):
        out = tc.gemm(A.to(DEV), W.to(DEV), bias=b.to(DEV), resid=R.to(DEV), alpha=0.5, act=act,
                      out_dtype=torch.float32)
        want = 0.5 * fn(A.float() @ W.float().T + b) + R
        assert relerr(out, want) < 1e-4, act
    # bf16 output with padded leading dimension, bf16 residual
    Rb = torch.zeros(M, 416, dtype=torch.bfloat16)  # residual rows must be 16-byte aligned (TMA)
    Rb[:, :N] = R.to(torch.bfloat16)
    out = tc.gemm(A.to(DEV), W.to(DEV), bias=b.to(DEV), resid=Rb.to(DEV), act=None, out_dtype=torch.bfloat16, ldd=416,
                  N=N)
    want = A.float() @ W.float().T + b + Rb[:, :N].float()
    assert out.shape == (M, 416) and float(out[:, 404:].abs().max()) == 0
    assert relerr(out[:, :404].float(), want) < 1e-2 and cosine(out[:, :404].float(), want) > 0.9999


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 16, 64, 64), (2, 16, 32, 128, 96), (1, 14, 18, 203, 404),
                                            (1, 32, 64, 384, 384), (2, 9, 11, 404, 404)])
def test_conv3x3(tc, B, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(B, Cin, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g)
    ldx = tc.round_up(Cin, 8)
    xn = torch.full((B, H, W, ldx), float("nan"), dtype=torch.bfloat16)  # pad channels are poison: must be ignored
    xn[..., :Cin] = x.permute(0, 2, 3, 1)
    y = tc.conv3x3(xn.to(DEV), tc.pack_conv3x3_weight(w).to(DEV), b.to(DEV), Cin, Cout, act="relu",
                   out_dtype=torch.float32)
    want = torch.relu(F.conv2d(x.float(), w.float(), b, padding=1)).permute(0, 2, 3, 1)
    assert relerr(y[..., :Cout], want) < 1e-4, relerr(y[..., :Cout], want)


@pytest.mark.parametrize("M,D,N2", [(1000, 404, 448), (70001, 404, 384), (300, 384, 1536)])
def test_gemm_fused_layernorm(tc, M, D, N2):
    """Producer GEMM writes per-row (sum, sum of squares) slots of its stored output; the consumer GEMM applies
    LayerNorm(gamma, beta) of those rows in its epilogue (isp_gemm_bf16_tc_ex).  Checked against LN-then-Linear in
    fp32 on the same bf16-stored activations (loftup/layers.py:161-174,186-202)."""
    g = torch.Generator().manual_seed(M + D)
    K0 = 384
    A = torch.randn(M, K0, generator=g).to(torch.bfloat16)
    W0 = (torch.randn(D, K0, generator=g) * K0 ** -0.5).to(torch.bfloat16)
    b0 = torch.randn(D, generator=g) * 0.5 + 0.7   # a mean well away from zero (cancellation in rstd*(acc - mean*g))
    Dp = tc.round_up(D, 16)
    R = torch.zeros(M, Dp, dtype=torch.bfloat16)
    R[:, :D] = torch.randn(M, D, generator=g).to(torch.bfloat16)
    for resid in (None, R):
        slots = tc.stats_slots(D, torch.bfloat16, resid is not None)
        st = torch.full((M, slots, 2), float("nan"), device=DEV)
        x = tc.gemm(A.to(DEV), W0.to(DEV), bias=b0.to(DEV), resid=None if resid is None else resid.to(DEV),
                    out_dtype=torch.bfloat16, N=D, ldd=Dp, stats_out=st)
        xs = x[:, :D].float().cpu()
        ssum = st.sum(1).cpu()
        assert torch.allclose(ssum[:, 0], xs.sum(1), rtol=1e-4, atol=1e-3)
        assert torch.allclose(ssum[:, 1], (xs * xs).sum(1), rtol=1e-4, atol=1e-3)
        gamma, beta = torch.randn(D, generator=g) * 0.3 + 1.0, torch.randn(D, generator=g) * 0.2
        W1 = torch.randn(N2, D, generator=g) * D ** -0.5
        b1 = torch.randn(N2, generator=g)
        Wg, gsum, bias = tc.pack_ln_linear(W1, b1, gamma, beta)
        out = tc.gemm(x, Wg.to(DEV), bias=bias.to(DEV), out_dtype=torch.float32, N=N2, K=D, ln_stats=st,
                      ln_g=gsum.to(DEV), ln_eps=1e-5)
        want = F.layer_norm(xs, (D,), gamma, beta, 1e-5) @ W1.T + b1
        assert relerr(out, want) < 1e-2 and cosine(out, want) > 0.9999, (relerr(out, want), cosine(out, want))


def test_conv3x3_row_stats(tc):
    g = torch.Generator().manual_seed(11)
    B, H, W, Cin, Cout = 2, 9, 21, 64, 404
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g)
    st = torch.full((B * H * W, tc.stats_slots(Cout), 2), float("nan"), device=DEV)
    y = tc.conv3x3(x.to(DEV), tc.pack_conv3x3_weight(w).to(DEV), b.to(DEV), Cin, Cout, act="relu", ldy=416,
                   stats_out=st)
    ys = y.view(-1, 416)[:, :Cout].float()
    assert torch.allclose(st.sum(1)[:, 0], ys.sum(1), rtol=1e-4, atol=1e-3)
    assert torch.allclose(st.sum(1)[:, 1], (ys * ys).sum(1), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 224, 352, 64, 96), (1, 160, 256, 128, 404), (2, 161, 250, 64, 80)])
def test_conv3x3_paired_tiles(tc, B, H, W, Cin, Cout):
    """Sizes with at least two waves of tile PAIRS: the kernel computes two vertically adjacent 128-pixel tiles per CTA
    from one weight tile (gemm_tc.cu Params::pair); also the row statistics and the ReLU-mask dgrad in that mode."""
    from isegprobe_b200 import _lib
    g = torch.Generator().manual_seed(Cin + Cout + H)
    x = torch.randn(B, Cin, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g)
    xn = x.permute(0, 2, 3, 1).contiguous()
    ldy = tc.round_up(Cout, 16)
    st = torch.full((B * H * W, tc.stats_slots(Cout), 2), float("nan"), device=DEV)
    y = tc.conv3x3(xn.to(DEV), tc.pack_conv3x3_weight(w).to(DEV), b.to(DEV), Cin, Cout, act="relu", ldy=ldy, stats_out=st)
    want = torch.relu(F.conv2d(x.float(), w.float(), b, padding=1)).permute(0, 2, 3, 1)
    ys = y[..., :Cout].float().cpu()
    assert relerr(ys, want) < 1e-2 and cosine(ys, want) > 0.9999
    assert float(y[..., Cout:].float().abs().max()) == 0 if ldy > Cout else True
    ysf = ys.reshape(-1, Cout)
    assert torch.allclose(st.sum(1)[:, 0].cpu(), ysf.sum(1), rtol=1e-4, atol=1e-3)
    y32 = tc.conv3x3(xn.to(DEV), tc.pack_conv3x3_weight(w).to(DEV), b.to(DEV), Cin, Cout, act=None, out_dtype=torch.float32)
    want32 = F.conv2d(x.float(), w.float(), b, padding=1).permute(0, 2, 3, 1)
    assert relerr(y32[..., :Cout], want32) < 1e-4
    if Cin == Cout or Cin % 8 == 0 and Cout % 8 == 0:
        # dgrad of that conv with the ReLU mask of its input: dX = conv_transpose(dY, W) * (act > 0)
        dy = torch.randn(B, H, W, Cout, generator=g).to(torch.bfloat16)
        act = torch.relu(torch.randn(B, H, W, Cin, generator=g)).to(torch.bfloat16)
        wT = tc.pack_conv3x3_weight(w.float().flip(2, 3).transpose(0, 1)).to(DEV)
        dx = torch.empty(B, H, W, Cin, dtype=torch.bfloat16, device=DEV)
        dy_d, act_d = dy.to(DEV), act.to(DEV)  # keep the device copies alive across the raw-pointer call
        _lib.call("isp_conv3x3_dgrad_bf16_tc", _lib.dptr(dy_d), _lib.dptr(wT), _lib.dptr(act_d), Cin,
                  _lib.dptr(dx), 1, B, H, W, Cout, Cout, Cin, Cin, _lib.stream_ptr())
        wantd = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w.float(), padding=1).permute(0, 2, 3, 1) * (act.float() > 0)
        assert relerr(dx.float(), wantd) < 1e-2 and cosine(dx.float(), wantd) > 0.9999


@pytest.mark.parametrize("M", [128, 1000, 4096 + 37])
def test_ffn_fused_matches_two_gemms_and_fp32(M):
    """isp_ffn_fused_bf16_tc (LoftUp FeedForward, loftup/layers.py:161-174) vs the two-launch path it replaces (same fused
    LayerNorm, same tanh-form GELU: identical rounding points except that the hidden tile is rounded to bf16 in both) and vs
    the fp32 torch expression; row statistics of the output."""
    import torch.nn.functional as F
    from isegprobe_b200 import tc
    D, C, Dp = 404, 384, 416
    g = torch.Generator().manual_seed(M)
    x = torch.zeros(M, Dp)
    x[:, :D] = torch.randn(M, D, generator=g)
    xb = x.to(torch.bfloat16).to(DEV)
    gamma, beta = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    W1, b1 = torch.randn(C, D, generator=g) * D ** -0.5, 0.1 * torch.randn(C, generator=g)
    W2, b2 = torch.randn(D, C, generator=g) * C ** -0.5, 0.1 * torch.randn(D, generator=g)
    W1g, g1, b1f = [t.to(DEV) for t in tc.pack_ln_linear(W1, b1, gamma, beta)]
    W2p, b2d = tc.pack_linear_weight(W2).to(DEV), b2.to(DEV)
    xf = xb.float()[:, :D]
    st = torch.zeros(M, 3, 2, device=DEV)  # three slots, as a producer GEMM would write them
    st[:, 0, 0], st[:, 0, 1] = xf.sum(1), (xf * xf).sum(1)
    out, stats = tc.ffn_fused(xb, W1g, g1, b1f, W2p, b2d, D, D, st)
    # the two-launch path
    h1 = tc.gemm(xb, W1g, bias=b1f, act="gelu_tanh", out_dtype=torch.bfloat16, N=C, K=D, ln_stats=st, ln_g=g1, ln_eps=1e-5)
    st2 = torch.empty(M, tc.stats_slots(D, torch.bfloat16, True), 2, device=DEV)
    want2 = tc.gemm(h1, W2p, bias=b2d, resid=xb, out_dtype=torch.bfloat16, N=D, K=C, ldd=Dp, stats_out=st2)
    assert tuple(out.shape) == (M, Dp) and float(out[:, D:].float().abs().max()) == 0.0
    assert relerr(out[:, :D].float(), want2[:, :D].float()) < 1e-2
    # fp32 reference of the block
    ref = xf.cpu() + F.linear(F.gelu(F.linear(F.layer_norm(xf.cpu(), (D,), gamma, beta, 1e-5), W1, b1)), W2, b2)
    assert relerr(out[:, :D].float(), ref) < 2e-2 and cosine(out[:, :D].float(), ref) > 0.9999
    # statistics of the stored rows
    of = out.float()[:, :D]
    assert relerr(stats.sum(1)[:, 0], of.sum(1)) < 1e-4 and relerr(stats.sum(1)[:, 1], (of * of).sum(1)) < 1e-4


@pytest.mark.parametrize("N,K,resid,out_dtype", [(404, 448, True, torch.bfloat16), (448, 404, False, torch.bfloat16),
                                                 (384, 404, False, torch.float32), (384, 384, True, torch.float32),
                                                 (404, 384, True, torch.bfloat16)])
def test_gemm_cta_pair_resident_weights(tc, N, K, resid, out_dtype):
    """Shapes of the LoftUp transformer GEMMs at a row count that selects the CTA-pair kernel with the weight tile resident
    in shared memory (even number of row tiles, at least two waves of clusters; the last row tile is partial), against
    fp32 torch on the same bf16 operands, with bias + GELU / residual and the row statistics of the stored output."""
    M = 152 * 128 - 37
    g = torch.Generator().manual_seed(N + K)
    lda, ldd = tc.round_up(K, 16), tc.round_up(N, 16)
    A = torch.zeros(M, lda, dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g)
    R = None
    if resid:
        R = torch.zeros(M, ldd, dtype=out_dtype)
        R[:, :N] = torch.randn(M, N, generator=g).to(out_dtype)
    act = None if resid else "gelu"
    st = torch.full((M, tc.stats_slots(N, out_dtype, resid), 2), float("nan"), device=DEV)
    out = tc.gemm(A.to(DEV), tc.pack_linear_weight(W).to(DEV), bias=b.to(DEV), resid=None if R is None else R.to(DEV),
                  act=act, out_dtype=out_dtype, N=N, K=K, ldd=ldd, stats_out=st)
    want = A[:, :K].float() @ W.float().T + b
    want = F.gelu(want) if act else want + R[:, :N].float()
    got = out[:, :N].float().cpu()
    tol = 1e-2 if out_dtype == torch.bfloat16 else 1e-4
    assert relerr(got, want) < tol and cosine(got, want) > 0.9999, (relerr(got, want), cosine(got, want))
    assert ldd == N or float(out[:, N:].float().abs().max()) == 0.0
    assert torch.allclose(st.sum(1)[:, 0].cpu(), got.sum(1), rtol=1e-4, atol=2e-3)
    assert torch.allclose(st.sum(1)[:, 1].cpu(), (got * got).sum(1), rtol=1e-4, atol=2e-3)
