"""GPU parity: tcgen05 flash attention and the row-wise helper kernels vs plain torch fp32
references of the same ops (bf16 operands; P is rounded to bf16 inside the kernel, so the
tolerance is 2e-2 of max / cosine >= 0.9995)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import loftup as oloft
from oracle import synth
from tests.gpu_util import DEV, cosine, relerr

pytestmark = pytest.mark.gpu


def _call(name, *a):
    from isegprobe_b200 import _lib
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


def _attention(Q, K, V, heads, hd, variant, q_head_stride, big_logits=False, lsum=False, poly=0):
    """Q [B,T,heads*hd(+pad)], K,V [B,S,heads,hd] float tensors (bf16-representable).
    lsum / poly: the tuned entry isp_attention_bf16_tc_opt (ones row in V^T at index hd, polynomial exp2)."""
    B, T = Q.shape[:2]
    S = K.shape[1]
    DV, DKC = (112, 128) if variant else (64, 64)  # variant 1: two-tile kernel, 3: one-tile kernel, same geometry
    Spad = (S + 127) // 128 * 128
    Kp = torch.zeros(B, heads, Spad, DKC, dtype=torch.bfloat16)
    Kp[:, :, :S, :hd] = K.permute(0, 2, 1, 3).to(torch.bfloat16)
    Vt = torch.zeros(B, heads, DV, Spad, dtype=torch.bfloat16)
    Vt[:, :, :hd, :S] = V.permute(0, 2, 3, 1).to(torch.bfloat16)
    ldq = Q.shape[2]
    Qd = Q.reshape(B * T, ldq).to(torch.bfloat16).to(DEV)
    out = torch.zeros(B * T, heads * DV, dtype=torch.bfloat16, device=DEV)
    if lsum:
        Vt[:, :, hd, :S] = 1.0
        _call("isp_attention_bf16_tc_opt", Qd, ldq, q_head_stride, Kp.to(DEV), Vt.to(DEV), out, heads * DV, DV, B, T, heads,
              S, variant, None, hd, poly)
    else:
        _call("isp_attention_bf16_tc", Qd, ldq, q_head_stride, Kp.to(DEV), Vt.to(DEV), out, heads * DV, DV, B, T, heads, S,
              variant)
    torch.cuda.synchronize()
    out = out.float().cpu().reshape(B, T, heads, DV)
    q = torch.stack([Q[:, :, h * q_head_stride:h * q_head_stride + hd] for h in range(heads)], 2)  # B,T,h,hd
    s = torch.einsum("bthd,bshd->bhts", q.float(), K.float())
    want = torch.einsum("bhts,bshd->bthd", torch.softmax(s, -1), V.float())
    if lsum:  # the ones-column holds the normalised row sum: 1.0
        dev1 = float((out[..., hd] - 1.0).abs().max())
        assert dev1 < 1e-2, dev1
        assert float(out[..., hd + 1:].abs().max()) == 0.0
    elif DV > hd:
        assert float(out[..., hd:].abs().max()) == 0.0  # padded head columns stay zero
    return out[..., :hd], want


@pytest.mark.parametrize("B,T,S,heads,hd,variant", [(1, 128, 128, 1, 64, 0), (2, 300, 257, 6, 64, 0),
                                                    (1, 256, 1024, 4, 101, 1), (2, 1025, 1025, 6, 64, 0),
                                                    (1, 1000, 196, 4, 101, 1), (2, 1000, 196, 4, 101, 3),
                                                    (3, 130, 1024, 4, 101, 1), (1, 5000, 1024, 4, 101, 1)])
def test_attention_vs_torch(B, T, S, heads, hd, variant):
    g = torch.Generator().manual_seed(T + S)
    bf = lambda x: x.to(torch.bfloat16).float()
    qs = (hd + 15) // 16 * 16  # per-head column stride of Q (16-byte aligned TMA box starts)
    Q = torch.zeros(B, T, heads * qs)
    for h in range(heads):
        Q[:, :, h * qs:h * qs + hd] = bf(torch.randn(B, T, hd, generator=g) * hd ** -0.25)
    K = bf(torch.randn(B, S, heads, hd, generator=g) * hd ** -0.25)
    V = bf(torch.randn(B, S, heads, hd, generator=g))
    out, want = _attention(Q, K, V, heads, hd, variant, qs)
    assert relerr(out, want) < 2e-2 and cosine(out, want) > 0.9995, (relerr(out, want), cosine(out, want))


@pytest.mark.parametrize("poly", [0, 2, 3, 4])
@pytest.mark.parametrize("B,T,S", [(1, 256, 1024), (2, 1000, 196), (1, 300, 1000)])
def test_attention_tuned_entry_vs_torch(B, T, S, poly):
    """isp_attention_bf16_tc_opt: row sum accumulated by the tensor pipe in the ones-column, `poly` of every 8 exponentials
    on the FMA pipe; same tolerance as the plain entry (S = 196 / 1000: masked keys in the last block; big logits exercise the
    lazy rescale with the ones-column)."""
    heads, hd = 4, 101
    g = torch.Generator().manual_seed(T + S + poly)
    bf = lambda x: x.to(torch.bfloat16).float()
    qs = 112
    Q = torch.zeros(B, T, heads * qs)
    for h in range(heads):
        Q[:, :, h * qs:h * qs + hd] = bf(torch.randn(B, T, hd, generator=g) * hd ** -0.25)
    K = bf(torch.randn(B, S, heads, hd, generator=g) * hd ** -0.25)
    if S == 1000:
        K[:, 600:] *= 6.0  # later key blocks carry much larger scores: the running max jumps by more than 2^8
    V = bf(torch.randn(B, S, heads, hd, generator=g))
    out, want = _attention(Q, K, V, heads, hd, 1, qs, lsum=True, poly=poly)
    assert relerr(out, want) < 2e-2 and cosine(out, want) > 0.9995, (relerr(out, want), cosine(out, want))


def test_attention_peaky_rows_rescale_path():
    """Logits whose running max jumps by far more than 2^8 between key blocks exercise the
    lazy O-rescale branch; rows must still match."""
    g = torch.Generator().manual_seed(0)
    B, T, S, heads, hd = 1, 256, 512, 2, 64
    bf = lambda x: x.to(torch.bfloat16).float()
    Q = bf(torch.randn(B, T, heads * hd, generator=g))
    K = bf(torch.randn(B, S, heads, hd, generator=g))
    K[:, 300:] *= 4.0  # later blocks carry much larger scores
    V = bf(torch.randn(B, S, heads, hd, generator=g))
    out, want = _attention(Q, K, V, heads, hd, 0, hd)
    assert relerr(out, want) < 2e-2 and cosine(out, want) > 0.9995


def test_layernorm_rows():
    g = torch.Generator().manual_seed(1)
    for C, ldi, ldo, ib, ob in [(404, 416, 416, 1, 1), (384, 384, 384, 0, 0), (203, 208, 208, 0, 1), (404, 404, 448, 0, 1)]:
        x = torch.randn(777, ldi, generator=g)
        w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
        xin = x.to(torch.bfloat16) if ib else x
        out = torch.full((777, ldo), 7.0, dtype=torch.bfloat16 if ob else torch.float32, device=DEV)
        _call("isp_layernorm_rows", xin.to(DEV), ib, ldi, out, ob, ldo, w.to(DEV), b.to(DEV), 777, C, 1e-5)
        want = F.layer_norm(xin.float()[:, :C], (C,), w, b, 1e-5)
        assert relerr(out[:, :C].float(), want) < (1e-2 if ob else 1e-5)
        if ldo > C:
            assert float(out[:, C:].float().abs().max()) == 0


def test_fourier_chnorm_and_lr_prepare():
    sd = synth.loftup_state_dict(384, seed=0)
    cn = synth.channelnorm_state_dict(384, seed=1)
    B, H, W = 2, 28, 42
    img = (synth.image_batch(B, H, W, seed=1) - 0.45) / 0.225
    x = oloft.fourier_features(oloft.minmax_scale(img), sd["fourier_feat.1.biases"], 20, True)
    want = oloft.channel_layernorm(x, sd["first_conv.0.norm.weight"], sd["first_conv.0.norm.bias"], 1e-5)
    d = img.to(DEV)
    mm = torch.empty(6, dtype=torch.int32, device=DEV)
    _call("isp_minmax_per_channel", d, mm, B, H, W, d.stride(0), d.stride(1))
    out = torch.empty(B, H, W, 208, dtype=torch.bfloat16, device=DEV)
    to = lambda t: t.contiguous().to(DEV)
    _call("isp_loftup_fourier_chnorm", d, *d.stride(), mm, to(torch.linspace(-1, 1, H)), to(torch.linspace(-1, 1, W)),
          to(torch.exp(torch.linspace(-2, 10, 20))), to(sd["fourier_feat.1.biases"][0].flatten()),
          to(sd["fourier_feat.1.biases"][1].flatten()), to(sd["first_conv.0.norm.weight"]),
          to(sd["first_conv.0.norm.bias"]), out, B, H, W, 208, 1e-5)
    got = out.float().cpu()[..., :203].permute(0, 3, 1, 2)
    # bf16 output rounding (2^-9) on values of magnitude <= ~4; sin/cos of 2e4-rad arguments agree to ~1e-3
    assert (got - want).abs().max() < 3e-2 and cosine(got, want) > 0.9999
    assert float(out[..., 203:].float().abs().max()) == 0
    # LR side
    lr = synth.lr_features(B, 384, 4, 6, seed=2)
    lrd = lr.to(DEV)
    kv = torch.empty(B * 24, 404, device=DEV)
    _call("isp_loftup_lr_prepare", lrd, *lrd.stride(), to(cn["norm.weight"]), to(cn["norm.bias"]),
          to(torch.linspace(-1, 1, 4)), to(torch.linspace(-1, 1, 6)), to(torch.exp(torch.linspace(-2, 10, 5))),
          to(sd["lr_pe.biases"][0].flatten()), to(sd["lr_pe.biases"][1].flatten()), kv, B, 384, 4, 6, 1e-5)
    want_kv = oloft.keys_values(oloft.channel_layernorm(lr, cn["norm.weight"], cn["norm.bias"], 1e-5), sd)
    assert relerr(kv.reshape(B, 24, 404), want_kv) < 1e-4


def test_repack_patchify_tokens():
    g = torch.Generator().manual_seed(2)
    B, T, heads, hd = 2, 50, 3, 20
    src = torch.randn(B * T, 200, generator=g)
    k = torch.empty(B, heads, 128, 64, dtype=torch.bfloat16, device=DEV)
    vt = torch.empty(B, heads, 64, 128, dtype=torch.bfloat16, device=DEV)
    _call("isp_repack_heads", src.to(DEV), 0, 200, 60, hd, k, B, T, 128, heads, 64, 0)
    _call("isp_repack_heads", src.to(DEV), 0, 200, 60, hd, vt, B, T, 128, heads, 64, 1)
    want = torch.zeros(B, heads, 128, 64)
    want[:, :, :T, :hd] = src.reshape(B, T, 200)[:, :, 60:60 + heads * hd].reshape(B, T, heads, hd).permute(0, 2, 1, 3)
    assert torch.equal(k.float().cpu(), want.to(torch.bfloat16).float())
    assert torch.equal(vt.float().cpu(), want.transpose(2, 3).to(torch.bfloat16).float())
    img = torch.randn(2, 3, 28, 42, generator=g)
    d = img.to(DEV)
    out = torch.empty(2 * 6, 592, dtype=torch.bfloat16, device=DEV)
    _call("isp_vit_patchify", d, *d.stride(), out, 2, 3, 28, 42, 14, 592)
    want = F.unfold(img, 14, stride=14).transpose(1, 2).reshape(12, 588)
    assert torch.equal(out[:, :588].float().cpu(), want.to(torch.bfloat16).float())
    assert float(out[:, 588:].float().abs().max()) == 0
