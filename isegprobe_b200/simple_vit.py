"""SimpleViTFeaturizer -- the TRAINABLE click embedding of models/sbd/dinov2/simple-vit_noup.py (`embed_coords` type
"simple_vit", core/utils/model_builder.py:39-49): a small pre-norm ViT over the 3-channel click / previous-mask map
(core/model/featurizers/simple_ViT.py:96-155, adapted there from lucidrains/vit-pytorch).  Same constructor keywords and
state-dict keys as the reference class, forward `[B, C, H, W] -> [B, N, dim]`.

Forward and backward run on libisp_b200: tcgen05 GEMMs (activation gradients with transposed packed weights, weight
gradients dW = dY^T X on the reduction-major batched GEMM, split over the tokens), the flash attention kernel forward /
the materialised-probability attention backward shared with the frozen ViTs, LayerNorm forward / input-gradient /
affine-gradient row kernels.  Every parameter gets its gradient (this module is trained, not frozen)."""
from typing import Tuple, Union

import torch
import torch.nn as nn

from . import _lib, tc
from .featurizers import DINOv2Featurizer, _call, _ln


def _pair(t):
    return tuple(t) if isinstance(t, (tuple, list)) and len(t) == 2 else (t, t)


def posemb_sincos_2d(h, w, dim, temperature: int = 10000):
    """simple_ViT.py:18-28 (fixed table, not a parameter)."""
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    assert dim % 4 == 0, "feature dimension must be multiple of 4 for sincos emb"
    omega = torch.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    y = y.flatten()[:, None] * omega[None, :]
    x = x.flatten()[:, None] * omega[None, :]
    return torch.cat((x.sin(), x.cos(), y.sin(), y.cos()), dim=1).float()


class _Attention(nn.Module):
    def __init__(self, dim, heads, dim_head):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.to_qkv = nn.Linear(dim, heads * dim_head * 3, bias=False)
        self.to_out = nn.Linear(heads * dim_head, dim, bias=False)


class _FeedForward(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim))


class _Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.layers = nn.ModuleList([nn.ModuleList([_Attention(dim, heads, dim_head), _FeedForward(dim, mlp_dim)])
                                     for _ in range(depth)])


class SimpleViTFeaturizer(nn.Module):
    def __init__(self, *, image_size: Union[int, Tuple[int]], patch_size: Union[int, Tuple[int]], dim: int, depth: int,
                 heads: int, mlp_dim: int, channels: int = 3, dim_head: int = 64) -> None:
        super().__init__()
        ih, iw = _pair(image_size)
        ph, pw = _pair(patch_size)
        assert ih % ph == 0 and iw % pw == 0, "Image dimensions must be divisible by the patch size."
        if ph != pw or dim_head != 64:
            raise NotImplementedError("SimpleViTFeaturizer: square patches and dim_head = 64 (the shipped configuration)")
        self.patch, self.channels, self.dim, self.heads, self.dim_head = ph, channels, dim, heads, dim_head
        pd = channels * ph * pw
        # index 0 is the parameter-free Rearrange of the reference: keeps the keys to_patch_embedding.{1,2,3}.*
        self.to_patch_embedding = nn.Sequential(nn.Identity(), nn.LayerNorm(pd), nn.Linear(pd, dim), nn.LayerNorm(dim))
        self.transformer = _Transformer(dim, depth, heads, dim_head, mlp_dim)
        self.image_hw, self.patch_hw = (ih, iw), (ph, pw)
        self._packed, self._pos = None, {}

    # ------------------------------------------------------------------ parameters in a fixed order
    def _param_list(self):
        pe, tr = self.to_patch_embedding, self.transformer
        ps = [pe[1].weight, pe[1].bias, pe[2].weight, pe[2].bias, pe[3].weight, pe[3].bias]
        for at, ff in tr.layers:
            ps += [at.norm.weight, at.norm.bias, at.to_qkv.weight, at.to_out.weight, ff.net[0].weight, ff.net[0].bias,
                   ff.net[1].weight, ff.net[1].bias, ff.net[3].weight, ff.net[3].bias]
        ps += [tr.norm.weight, tr.norm.bias]
        return ps

    def _perm(self, dev):
        """ours (c, p1, p2) patch-vector index -> the reference's (p1, p2, c) index ('b c (h p1) (w p2) -> b (h w) (p1 p2 c)')."""
        p, Cin = self.patch, self.channels
        o = torch.arange(Cin * p * p, device=dev)
        c, ij = o // (p * p), o % (p * p)
        return ij * Cin + c

    def _pack(self, dev):
        key = (str(dev), sum(q._version for q in self.parameters()))
        if self._packed is not None and self._packed["key"] == key:
            return self._packed
        f32 = lambda t: t.detach().float().contiguous().to(dev)
        pw = lambda w: tc.pack_linear_weight(w.detach().float()).to(dev)
        pwT = lambda w: tc.pack_linear_weight(w.detach().float().t().contiguous()).to(dev)
        pe, tr = self.to_patch_embedding, self.transformer
        perm = self._perm(pe[1].weight.device)
        We = pe[2].weight.detach().float()[:, perm]
        P = {"key": key, "g1": f32(pe[1].weight.detach()[perm]), "b1": f32(pe[1].bias.detach()[perm]),
             "We": pw(We), "WeT": pwT(We), "be": f32(pe[2].bias), "g3": f32(pe[3].weight), "b3": f32(pe[3].bias),
             "gn": f32(tr.norm.weight), "bn": f32(tr.norm.bias), "layers": []}
        inner, sc = self.heads * self.dim_head, self.dim_head ** -0.5
        for at, ff in tr.layers:
            Wqkv = at.to_qkv.weight.detach().float().clone()
            Wqkv[:inner] *= sc  # the attention kernels expect pre-scaled queries
            P["layers"].append({
                "ga": f32(at.norm.weight), "ba": f32(at.norm.bias), "Wqkv": pw(Wqkv), "WqkvT": pwT(Wqkv),
                "Wo": pw(at.to_out.weight), "WoT": pwT(at.to_out.weight),
                "gf": f32(ff.net[0].weight), "bf": f32(ff.net[0].bias),
                "W1": pw(ff.net[1].weight), "W1T": pwT(ff.net[1].weight), "b1": f32(ff.net[1].bias),
                "W2": pw(ff.net[3].weight), "W2T": pwT(ff.net[3].weight), "b2": f32(ff.net[3].bias)})
        self._packed = P
        return P

    def _pos_table(self, h, w, dev):
        key = (h, w, str(dev))
        if key not in self._pos:
            self._pos[key] = posemb_sincos_2d(h, w, self.dim).to(dev)
        return self._pos[key]

    # ------------------------------------------------------------------ forward
    def forward(self, img: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _SimpleViTFn.apply(self, img, *self._param_list())
        return self._forward_impl(img, None)

    def _forward_impl(self, img, saved):
        x = img.detach().float()
        dev, bf = x.device, torch.bfloat16
        B, Cin, H, W = x.shape
        p, dim, nh, hd = self.patch, self.dim, self.heads, self.dim_head
        inner = nh * hd
        h, w = H // p, W // p
        N, M, K = h * w, B * h * w, Cin * p * p
        P = self._pack(dev)
        Kp = P["We"].shape[1]
        cols = torch.empty(M, Kp, dtype=bf, device=dev)
        _call("isp_vit_patchify", x, *x.stride(), cols, B, Cin, H, W, p, Kp)
        x0n = _ln(cols, P["g1"], P["b1"], K, 1e-5, bf, Kp)
        e = tc.gemm(x0n, P["We"], bias=P["be"], out_dtype=torch.float32, K=K)
        xs = _ln(e, P["g3"], P["b3"], dim, 1e-5, torch.float32)
        xs.view(B, N, dim).add_(self._pos_table(h, w, dev))
        Tp = tc.round_up(N, 128)
        if saved is not None:
            saved.update({"B": B, "N": N, "cols": cols, "x0n": x0n, "e": e, "layers": []})
        for L in P["layers"]:
            x_in = xs
            hn = _ln(xs, L["ga"], L["ba"], dim, 1e-5, bf)
            qkv = tc.gemm(hn, L["Wqkv"], out_dtype=bf)
            Kh = torch.empty(B, nh, Tp, 64, dtype=bf, device=dev)
            Vt = torch.empty(B, nh, 64, Tp, dtype=bf, device=dev)
            _call("isp_repack_heads", qkv, 1, 3 * inner, inner, hd, Kh, B, N, Tp, nh, 64, 0)
            _call("isp_repack_heads", qkv, 1, 3 * inner, 2 * inner, hd, Vt, B, N, Tp, nh, 64, 1)
            O = torch.empty(M, inner, dtype=bf, device=dev)
            _call("isp_attention_bf16_tc", qkv, 3 * inner, hd, Kh, Vt, O, inner, hd, B, N, nh, N, 0)
            xs = tc.gemm(O, L["Wo"], resid=xs, out_dtype=torch.float32)
            hn2 = _ln(xs, L["gf"], L["bf"], dim, 1e-5, bf)
            h1 = tc.gemm(hn2, L["W1"], bias=L["b1"], act="gelu", out_dtype=bf)
            if saved is not None:
                saved["layers"].append((x_in, qkv, O, xs, h1))
            xs = tc.gemm(h1, L["W2"], bias=L["b2"], resid=xs, out_dtype=torch.float32)
        if saved is not None:
            saved["x_final"] = xs
        return _ln(xs, P["gn"], P["bn"], dim, 1e-5, torch.float32).view(B, N, dim)

    def reshape_feats_to_patches(self, feats: torch.Tensor) -> torch.Tensor:
        """simple_ViT.py:148-155."""
        B = feats.shape[0]
        return feats.reshape(B, self.image_hw[0] // self.patch_hw[0], self.image_hw[1] // self.patch_hw[1], -1).permute(0, 3, 1, 2)

    # ------------------------------------------------------------------ backward (all parameter gradients)
    @staticmethod
    def _wgrad(dy, x, n_out, k_in):
        """dW [n_out, k_in] fp32 = dy[:, :n_out]^T x[:, :k_in] (bf16 operands as stored, reduction over the rows split over
        the batched GEMM's batch dimension)."""
        Mr = dy.shape[0]
        nsplit = next((n for n in (16, 8, 4, 2) if Mr % n == 0 and Mr // n >= 512), 1)
        ck = Mr // nsplit
        part = torch.empty(nsplit, n_out, k_in, dtype=torch.float32, device=dy.device)
        lda, ldx = dy.stride(0), x.stride(0)
        _lib.call("isp_gemm_bf16_tc_batched_tn", _lib.dptr(dy), lda, Mr * lda, ck * lda, _lib.dptr(x), ldx, Mr * ldx, ck * ldx,
                  _lib.dptr(part), k_in, n_out * k_in, n_out * k_in, 0, n_out, k_in, ck, 1, nsplit, 1.0, _lib.stream_ptr())
        return part.sum(0) if nsplit > 1 else part[0]

    @staticmethod
    def _colsum(x_bf, C):
        out = torch.zeros(C, dtype=torch.float32, device=x_bf.device)
        for c0 in range(0, C, 1024):  # the kernel takes up to 1024 columns per call
            _call("isp_colsum_bf16", x_bf[:, c0:], x_bf.stride(0), out[c0:], x_bf.shape[0], min(1024, C - c0))
        return out

    @staticmethod
    def _ln_affine(dy, x, C, eps=1e-5):
        dg = torch.zeros(C, dtype=torch.float32, device=dy.device)
        db = torch.zeros(C, dtype=torch.float32, device=dy.device)
        _call("isp_layernorm_affine_bwd", dy, dy.stride(0), x, int(x.dtype == torch.bfloat16), x.stride(0), dg, db,
              x.shape[0], C, float(eps))
        return dg, db

    def _backward_impl(self, saved, grad_out):
        dev, bf = grad_out.device, torch.bfloat16
        P = self._pack(dev)
        B, N = saved["B"], saved["N"]
        dim, nh, hd = self.dim, self.heads, self.dim_head
        inner, M = nh * hd, B * N
        K = self.channels * self.patch * self.patch
        mlp = P["layers"][0]["W1"].shape[0]
        ln_bwd, attn_bwd = DINOv2Featurizer._ln_bwd, DINOv2Featurizer._attention_bwd
        g = grad_out.detach().float().reshape(M, dim).contiguous()
        dgn, dbn = self._ln_affine(g, saved["x_final"], dim)
        dx, dxb = ln_bwd(g, saved["x_final"], P["gn"], None, dim, 1e-5)
        layer_grads = []
        for L, (x_in, qkv, O, x_mid, h1) in zip(reversed(P["layers"]), reversed(saved["layers"])):
            # FeedForward: x_out = x_mid + W2 gelu(W1 LN_f(x_mid) + b1) + b2
            db2 = self._colsum(dxb, dim)
            dW2 = self._wgrad(dxb, h1, dim, mlp)
            hn2 = _ln(x_mid, L["gf"], L["bf"], dim, 1e-5, bf)
            pre = tc.gemm(hn2, L["W1"], bias=L["b1"], out_dtype=bf)
            dh = tc.gemm(dxb, L["W2T"], out_dtype=bf)
            _call("isp_gelu_bwd_bf16", dh, pre, dh, dh.numel(), 0)
            db1 = self._colsum(dh, mlp)
            dW1 = self._wgrad(dh, hn2, mlp, dim)
            dn = tc.gemm(dh, L["W1T"], out_dtype=torch.float32)
            dgf, dbf = self._ln_affine(dn, x_mid, dim)
            dx, dxb = ln_bwd(dn, x_mid, L["gf"], dx, dim, 1e-5)
            # Attention: x_mid = x_in + to_out(attn(to_qkv(LN_a(x_in))))
            dWo = self._wgrad(dxb, O, dim, inner)
            dO = tc.gemm(dxb, L["WoT"], out_dtype=bf)
            dqkv = attn_bwd(None, qkv, dO, B, N, inner, nh)
            hn = _ln(x_in, L["ga"], L["ba"], dim, 1e-5, bf)
            dWqkv = self._wgrad(dqkv, hn, 3 * inner, dim)
            dWqkv[:inner] *= hd ** -0.5  # the packed query rows carry the 1/sqrt(d) scale
            dn = tc.gemm(dqkv, L["WqkvT"], out_dtype=torch.float32)
            dga, dba = self._ln_affine(dn, x_in, dim)
            dx, dxb = ln_bwd(dn, x_in, L["ga"], dx, dim, 1e-5)
            layer_grads.append([dga, dba, dWqkv, dWo, dgf, dbf, dW1, db1, dW2, db2])
        # patch embedding: x = LN_3(We LN_1(patches) + be) + pos
        dg3, db3 = self._ln_affine(dx, saved["e"], dim)
        de, deb = ln_bwd(dx, saved["e"], P["g3"], None, dim, 1e-5)
        dbe = self._colsum(deb, dim)
        perm = self._perm(dev)
        dWe_p = self._wgrad(deb, saved["x0n"], dim, K)          # columns in our (c, p1, p2) order
        dWe = torch.empty_like(dWe_p)
        dWe[:, perm] = dWe_p
        dn0 = tc.gemm(deb, P["WeT"], out_dtype=torch.float32, N=K, K=dim)
        dg1p, db1p = self._ln_affine(dn0, saved["cols"], K)
        dg1, db1 = torch.empty_like(dg1p), torch.empty_like(db1p)
        dg1[perm], db1[perm] = dg1p, db1p
        grads = [dg1, db1, dWe, dbe, dg3, db3]
        for lg in reversed(layer_grads):
            grads += lg
        grads += [dgn, dbn]
        return grads


class _SimpleViTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, img, *params):
        saved = {}
        out = mod._forward_impl(img, saved)
        ctx.mod, ctx.saved = mod, saved
        return out

    @staticmethod
    def backward(ctx, grad_out):
        grads = ctx.mod._backward_impl(ctx.saved, grad_out)
        ctx.saved = None
        return (None, None, *grads)
