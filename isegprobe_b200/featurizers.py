"""Frozen ViT backbone and the trainable click-map patch embedding on libisp_b200.

`DINOv2Featurizer` mirrors the reference adapter (core/model/featurizers/DINOv2.py:468-546):
same constructor, `forward(x, additional_features)`, `[B, C, h, w]` output, click-embedding
injection before the blocks.  The reference hub-loads `dinov2_vits14` (network); here the
ViT-S/14 is built locally with the hub/vendored parameter names (`model.blocks.{i}...`,
SURVEY Appendix A.5) so that checkpoint loads with `load_state_dict`, random-init otherwise.

Per block (dinov2/layers/block.py:92-117): LN -> QKV GEMM -> tcgen05 flash attention ->
proj GEMM (+LayerScale, +residual) -> LN -> fc1 GEMM + GELU -> fc2 GEMM (+LayerScale, +residual).
LayerScale and the 1/sqrt(d) query scale are folded into the packed weights.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, tc


def _call(name, *a):
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


def _ln(x, w, b, C, eps, out_dtype, ldo=None):
    ldo = C if ldo is None else ldo
    out = torch.empty(x.shape[0], ldo, dtype=out_dtype, device=x.device)
    _call("isp_layernorm_rows", x, int(x.dtype == torch.bfloat16), x.stride(0), out, int(out_dtype == torch.bfloat16),
          ldo, w, b, x.shape[0], C, float(eps))
    return out


class PatchEmbed(nn.Module):
    """Click-map embedding Conv2d(in_chans -> D, k = s = patch) -> [B, N, D]
    (core/model/featurizers/utils/patch_embed.py:12-42); same ctor and parameter names."""

    def __init__(self, img_size=(224, 224), patch_size=(16, 16), in_chans=3, embed_dim=768, norm_layer=None,
                 flatten=True):
        super().__init__()
        self.in_chans, self.img_size, self.patch_size = in_chans, img_size, patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()
        self._packed = None

    def _pack(self, dev):
        key = (str(dev), self.proj.weight._version, self.proj.bias._version)
        if self._packed is None or self._packed[0] != key:
            w = self.proj.weight.detach().float().reshape(self.proj.weight.shape[0], -1)
            self._packed = (key, tc.pack_linear_weight(w).to(dev), self.proj.bias.detach().float().to(dev))
        return self._packed[1], self._packed[2]

    def _embed(self, x: torch.Tensor):
        x = x.detach().float()
        B, Cin, H, W = x.shape
        p = self.patch_size[0]
        K = Cin * p * p
        Wp, b = self._pack(x.device)
        N = (H // p) * (W // p)
        cols = torch.empty(B * N, Wp.shape[1], dtype=torch.bfloat16, device=x.device)
        _call("isp_vit_patchify", x, *x.stride(), cols, B, Cin, H, W, p, Wp.shape[1])
        y = tc.gemm(cols, Wp, bias=b, out_dtype=torch.float32, K=K)
        return y.view(B, N, -1), cols, K

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, Cin, H, W = x.shape
        p = self.patch_size[0]
        if torch.is_grad_enabled() and (self.proj.weight.requires_grad or self.proj.bias.requires_grad):
            y = _PatchEmbedFn.apply(self, x, self.proj.weight, self.proj.bias)  # trainable click embedding
        else:
            y = self._embed(x)[0]
        if not self.flatten:
            y = y.transpose(1, 2).reshape(B, -1, H // p, W // p)
        return self.norm(y)


class _PatchEmbedFn(torch.autograd.Function):
    """Conv2d(k = s = patch) as a GEMM over patch rows; backward = weight / bias gradient (the input is the click
    map: no gradient).  dW = d_y^T cols on the batched tcgen05 GEMM (reduction over the tokens), fp32 result."""

    @staticmethod
    def forward(ctx, mod, x, weight, bias):
        y, cols, K = mod._embed(x)
        ctx.cols, ctx.K, ctx.wshape = cols, K, tuple(weight.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        cols, K = ctx.cols, ctx.K
        dev, bf = gy.device, torch.bfloat16
        M, Kp = cols.shape
        D = gy.shape[-1]
        g = gy.detach().reshape(M, D).to(bf).contiguous()
        Mp = tc.round_up(M, 8)
        gT = torch.empty(D, Mp, dtype=bf, device=dev)
        cT = torch.empty(K, Mp, dtype=bf, device=dev)
        _call("isp_transpose_bf16_batched", g, D, 0, gT, Mp, 0, 1, M, D)
        _call("isp_transpose_bf16_batched", cols, Kp, 0, cT, Mp, 0, 1, M, K)
        ldw = tc.round_up(K, 4)
        dW = torch.empty(D, ldw, dtype=torch.float32, device=dev)
        _lib.call("isp_gemm_bf16_tc_batched", _lib.dptr(gT), Mp, D * Mp, D * Mp, _lib.dptr(cT), Mp, K * Mp, K * Mp,
                  _lib.dptr(dW), ldw, D * ldw, D * ldw, 0, D, K, M, 1, 1, 1.0, _lib.stream_ptr())
        db = torch.zeros(D, dtype=torch.float32, device=dev)
        _call("isp_colsum_bf16", g, D, db, M, D)
        return None, None, dW[:, :K].reshape(ctx.wshape), db


class _Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio=4):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = nn.Module()
        self.attn.qkv = nn.Linear(dim, 3 * dim)
        self.attn.proj = nn.Linear(dim, dim)
        self.ls1 = nn.Module()
        self.ls1.gamma = nn.Parameter(torch.ones(dim))
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = nn.Module()
        self.mlp.fc1 = nn.Linear(dim, mlp_ratio * dim)
        self.mlp.fc2 = nn.Linear(mlp_ratio * dim, dim)
        self.ls2 = nn.Module()
        self.ls2.gamma = nn.Parameter(torch.ones(dim))


class _ViT(nn.Module):
    """Parameter container with DinoVisionTransformer's names (DINOv2.py:53-160)."""

    def __init__(self, dim=384, depth=12, heads=6, patch=14, img_size=518):
        super().__init__()
        self.embed_dim, self.num_heads, self.patch_size = dim, heads, patch
        n = (img_size // patch) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, dim))
        self.mask_token = nn.Parameter(torch.zeros(1, dim))
        self.patch_embed = nn.Module()
        self.patch_embed.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)
        self.blocks = nn.ModuleList([_Block(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)


class DINOv2Featurizer(nn.Module):
    def __init__(self, arch: str = "dinov2_vits14", feats_injection_mode: str = "no_injection") -> None:
        super().__init__()
        self.arch, self.feats_injection_mode = arch, feats_injection_mode
        if arch != "dinov2_vits14":
            raise NotImplementedError(f"Only 'dinov2_vits14' is supported, got {arch}")
        if feats_injection_mode not in ("no_injection", "before_backbone", "after_backbone"):
            raise NameError(f"Unknown feats_injection_mode: {feats_injection_mode}")
        self.model = _ViT()
        self.patch_size = self.model.patch_size
        self._packed = None
        self._pos_cache = {}

    feat_type = "token"  # DINOFeaturizer: "key" returns the keys of the last block instead of the normed tokens

    def _version(self):
        return sum(p._version for p in self.parameters())

    @staticmethod
    def _layer_scale(blk):
        """LayerScale gammas of a block (DINOv2), ones for the plain timm / DINO ViT."""
        w = blk.norm1.weight
        if hasattr(blk, "ls1"):
            return blk.ls1.gamma.detach().float(), blk.ls2.gamma.detach().float()
        return torch.ones_like(w, dtype=torch.float32), torch.ones_like(w, dtype=torch.float32)

    def _pack(self, dev):
        key = (str(dev), self._version())
        if self._packed is not None and self._packed["key"] == key:
            return self._packed
        m = self.model
        C, nh = m.embed_dim, m.num_heads
        f32 = lambda t: t.detach().float().contiguous().to(dev)
        P = {"key": key}
        P["Wpe"] = tc.pack_linear_weight(m.patch_embed.proj.weight.detach().float().reshape(C, -1)).to(dev)
        P["bpe"] = f32(m.patch_embed.proj.bias)
        P["cls"] = f32(m.cls_token.reshape(-1))
        blocks = []
        sc = (C // nh) ** -0.5
        for blk in m.blocks:
            Wqkv, bqkv = blk.attn.qkv.weight.detach().float().clone(), blk.attn.qkv.bias.detach().float().clone()
            Wqkv[:C] *= sc  # attention.py:66: q * scale
            bqkv[:C] *= sc
            g1, g2 = self._layer_scale(blk)
            blocks.append({
                "n1w": f32(blk.norm1.weight), "n1b": f32(blk.norm1.bias),
                "Wqkv": tc.pack_linear_weight(Wqkv).to(dev), "bqkv": f32(bqkv),
                "Wproj": tc.pack_linear_weight(blk.attn.proj.weight.detach().float() * g1[:, None]).to(dev),
                "bproj": f32(blk.attn.proj.bias.detach().float() * g1),
                "n2w": f32(blk.norm2.weight), "n2b": f32(blk.norm2.bias),
                "W1": tc.pack_linear_weight(blk.mlp.fc1.weight).to(dev), "b1": f32(blk.mlp.fc1.bias),
                "W2": tc.pack_linear_weight(blk.mlp.fc2.weight.detach().float() * g2[:, None]).to(dev),
                "b2": f32(blk.mlp.fc2.bias.detach().float() * g2),
            })
        P["blocks"] = blocks
        P["nw"], P["nb"] = f32(m.norm.weight), f32(m.norm.bias)
        self._packed = P
        self._pos_cache = {}
        return P

    def _pos(self, H, W, dev):
        """interpolate_pos_encoding (DINOv2.py:199-230): bicubic resize of the patch position
        table with the +0.1 scale-factor quirk.  Parameter preprocessing, cached per size."""
        key = (H, W, str(dev), self.model.pos_embed._version)
        if key not in self._pos_cache:
            pe = self.model.pos_embed.detach().float().cpu()
            N = pe.shape[1] - 1
            p = self.patch_size
            if (H // p) * (W // p) == N and H == W:
                out = pe
            else:
                w0, h0 = H // p + 0.1, W // p + 0.1  # reference passes (input_h, input_w) as (w, h)
                s = int(math.sqrt(N))
                pp = F.interpolate(pe[:, 1:].reshape(1, s, s, -1).permute(0, 3, 1, 2),
                                   scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
                out = torch.cat([pe[:, :1], pp.permute(0, 2, 3, 1).reshape(1, -1, pe.shape[-1])], 1)
            self._pos_cache[key] = out.reshape(-1, pe.shape[-1]).contiguous().to(dev)
        return self._pos_cache[key]

    def forward(self, x, additional_features=None):
        """Frozen backbone.  When the click embedding requires grad (training: it is injected before the blocks,
        DINOv2.py:518-523) the call goes through an autograd Function whose backward runs the activation backward
        of all twelve blocks on libisp_b200 and returns d(loss)/d(additional_features)."""
        if (torch.is_grad_enabled() and additional_features is not None and additional_features.requires_grad
                and self.feats_injection_mode == "before_backbone"):
            return _DinoBackboneFn.apply(self, x, additional_features)
        return self._forward_impl(x, additional_features, None)

    def _forward_impl(self, x, additional_features, saved):
        x = x.detach().float()
        dev = x.device
        B, _, H, W = x.shape
        p = self.patch_size
        h, w = H // p, W // p
        m = self.model
        C, nh = m.embed_dim, m.num_heads
        P = self._pack(dev)
        inject = additional_features is not None and self.feats_injection_mode != "no_injection"
        extra = None
        if inject and self.feats_injection_mode == "before_backbone":
            extra = additional_features.detach().float().contiguous()
            assert tuple(extra.shape) == (B, h * w, C), f"x.shape: {(B, h * w, C)}, additional_features.shape: {tuple(extra.shape)}"
        N, T = h * w, h * w + 1
        cols = torch.empty(B * N, P["Wpe"].shape[1], dtype=torch.bfloat16, device=dev)
        _call("isp_vit_patchify", x, *x.stride(), cols, B, 3, H, W, p, P["Wpe"].shape[1])
        patch = tc.gemm(cols, P["Wpe"], bias=P["bpe"], out_dtype=torch.float32, K=3 * p * p)
        xs = torch.empty(B * T, C, dtype=torch.float32, device=dev)
        _call("isp_vit_assemble_tokens", patch, extra, P["cls"], self._pos(H, W, dev), xs, B, N, C)
        Tp = tc.round_up(T, 128)
        hd = C // nh
        bf = torch.bfloat16
        if saved is not None:
            saved.update({"B": B, "T": T, "h": h, "w": w, "blocks": []})
        key_feats = None
        for li, L in enumerate(P["blocks"]):
            x_in = xs
            hn = _ln(xs, L["n1w"], L["n1b"], C, 1e-6, bf)
            if self.feat_type == "key" and li == len(P["blocks"]) - 1:
                # DINO.py:591-592: keys of the last block, class token dropped, channels head_dim-major / head-minor
                # (`permute(0, 2, 3, 1).flatten(-2)`); only the K rows of the qkv projection are needed, in fp32
                kk = tc.gemm(hn, L["Wqkv"][C:2 * C], bias=L["bqkv"][C:2 * C], out_dtype=torch.float32)
                key_feats = kk.view(B, T, nh, hd)[:, 1:].permute(0, 1, 3, 2).reshape(B, N, C)
                if saved is not None:
                    saved["key_x"] = x_in
                break
            qkv = tc.gemm(hn, L["Wqkv"], bias=L["bqkv"], out_dtype=bf)
            Kp = torch.empty(B, nh, Tp, 64, dtype=bf, device=dev)
            Vt = torch.empty(B, nh, 64, Tp, dtype=bf, device=dev)
            _call("isp_repack_heads", qkv, 1, 3 * C, C, hd, Kp, B, T, Tp, nh, 64, 0)
            _call("isp_repack_heads", qkv, 1, 3 * C, 2 * C, hd, Vt, B, T, Tp, nh, 64, 1)
            O = torch.empty(B * T, C, dtype=bf, device=dev)
            _call("isp_attention_bf16_tc", qkv, 3 * C, hd, Kp, Vt, O, C, hd, B, T, nh, T, 0)
            xs = tc.gemm(O, L["Wproj"], bias=L["bproj"], resid=xs, out_dtype=torch.float32)
            hn = _ln(xs, L["n2w"], L["n2b"], C, 1e-6, bf)
            h1 = tc.gemm(hn, L["W1"], bias=L["b1"], act="gelu_tanh", out_dtype=bf)
            if saved is not None:  # block input, qkv and the MLP's input: what the backward re-reads
                saved["blocks"].append((x_in, qkv, xs))
            xs = tc.gemm(h1, L["W2"], bias=L["b2"], resid=xs, out_dtype=torch.float32)
        if saved is not None:
            saved["x_final"] = xs
        if key_feats is not None:
            feats = key_feats
        else:
            xn = _ln(xs, P["nw"], P["nb"], C, 1e-6, torch.float32)
            feats = xn.view(B, T, C)[:, 1:]
        if inject and self.feats_injection_mode == "after_backbone":
            feats = feats + additional_features.to(feats.dtype)
        return feats.reshape(B, h, w, C).permute(0, 3, 1, 2)

    # ---------------------------------------------------------------- activation backward
    def _pack_bwd(self, dev):
        """Transposed packed weights of every block (dgrad of y = a W^T is dy W, i.e. a GEMM with W^T as weight)."""
        P = self._pack(dev)
        if "bwd" not in P:
            m = self.model
            C, nh = m.embed_dim, m.num_heads
            sc = (C // nh) ** -0.5
            out = []
            for blk in m.blocks:
                Wqkv = blk.attn.qkv.weight.detach().float().clone()
                Wqkv[:C] *= sc
                g1, g2 = self._layer_scale(blk)
                out.append({
                    "WqkvT": tc.pack_linear_weight(Wqkv.t().contiguous()).to(dev),
                    "WprojT": tc.pack_linear_weight((blk.attn.proj.weight.detach().float() * g1[:, None]).t().contiguous()).to(dev),
                    "W1T": tc.pack_linear_weight(blk.mlp.fc1.weight.detach().float().t().contiguous()).to(dev),
                    "W2T": tc.pack_linear_weight((blk.mlp.fc2.weight.detach().float() * g2[:, None]).t().contiguous()).to(dev),
                })
            P["bwd"] = out
        return P, P["bwd"]

    @staticmethod
    def _ln_bwd(dy, x, gamma, resid, C, eps=1e-6, ldb=None):
        """LayerNorm backward over rows (x fp32 or bf16): returns (dx fp32 [M,C], bf16 copy [M,ldb])."""
        M = x.shape[0]
        dx = torch.empty(M, C, dtype=torch.float32, device=x.device)
        ldb = C if ldb is None else ldb
        dxb = torch.empty(M, ldb, dtype=torch.bfloat16, device=x.device)
        if ldb > C:
            dxb[:, C:].zero_()
        _call("isp_layernorm_rows_bwd", dy, dy.stride(0), x, int(x.dtype == torch.bfloat16), x.stride(0), gamma, resid,
              0 if resid is None else resid.stride(0), dx, C, dxb, dxb.stride(0), M, C, float(eps))
        return dx, dxb

    flash_backward = True  # False: the materialised-probability path below (kept for head dims != 64 and as a cross-check)

    def _attention_bwd(self, qkv, dO, B, T, C, nh):
        """d(qkv) of O = softmax(Q K^T) V per (image, head) (dinov2/layers/attention.py:54-71; Q already carries the
        1/sqrt(d) scale) on the flash-style backward kernel: the forward attention is re-run for O and the rows'
        log-sum-exp (cheap at ViT sizes), then isp_attention_bwd_bf16_tc recomputes P / dS tile by tile on chip."""
        hd = C // nh
        if hd != 64 or not DINOv2Featurizer.flash_backward:
            return DINOv2Featurizer._attention_bwd_materialised(self, qkv, dO, B, T, C, nh)
        dev, bf, M = qkv.device, torch.bfloat16, B * T
        Tp = tc.round_up(T, 128)
        Kf = torch.empty(B, nh, Tp, 64, dtype=bf, device=dev)
        Vt = torch.empty(B, nh, 64, Tp, dtype=bf, device=dev)
        _call("isp_repack_heads", qkv, 1, 3 * C, C, hd, Kf, B, T, Tp, nh, 64, 0)
        _call("isp_repack_heads", qkv, 1, 3 * C, 2 * C, hd, Vt, B, T, Tp, nh, 64, 1)
        O = torch.empty(M, C, dtype=bf, device=dev)
        lse = torch.empty(B * nh * T + 64, dtype=torch.float32, device=dev)
        dvec = torch.empty(B * nh * T + 64, dtype=torch.float32, device=dev)
        _call("isp_attention_bf16_tc_lse", qkv, 3 * C, hd, Kf, Vt, O, C, hd, B, T, nh, T, 0, lse)
        _call("isp_attention_rowdot_heads", dO, C, O, C, dvec, B, T, nh, 64)
        Kb = torch.empty(B, nh, T, 64, dtype=bf, device=dev)   # rows = keys, no padding: the backward's K / V operands
        Vb = torch.empty(B, nh, T, 64, dtype=bf, device=dev)
        _call("isp_repack_heads", qkv, 1, 3 * C, C, hd, Kb, B, T, T, nh, 64, 0)
        _call("isp_repack_heads", qkv, 1, 3 * C, 2 * C, hd, Vb, B, T, T, nh, 64, 0)
        dK = torch.zeros(B, nh, T, 64, dtype=torch.float32, device=dev)
        dV = torch.zeros(B, nh, T, 64, dtype=torch.float32, device=dev)
        dQ = torch.zeros(M, C, dtype=torch.float32, device=dev)
        _call("isp_attention_bwd_bf16_tc", qkv, 3 * C, dO, C, Kb, Vb, lse, dvec, dK, dV, dQ, C, B, T, nh, T, 64)
        dqkv = torch.empty(M, 3 * C, dtype=bf, device=dev)
        dqkv[:, :C] = dQ
        dqkv[:, C:2 * C] = dK.permute(0, 2, 1, 3).reshape(M, C)
        dqkv[:, 2 * C:] = dV.permute(0, 2, 1, 3).reshape(M, C)
        return dqkv

    def _attention_bwd_materialised(self, qkv, dO, B, T, C, nh):
        """d(qkv) of O = softmax(Q K^T) V per (image, head), probabilities recomputed from Q and K
        (dinov2/layers/attention.py:54-71; Q already carries the 1/sqrt(d) scale).  All products are batched
        tcgen05 GEMMs over (image, head); P, dP, dS are materialised (T x T per head: 2 MB)."""
        dev, bf, hd = qkv.device, torch.bfloat16, C // nh
        Tp = tc.round_up(T, 8)
        st = _lib.stream_ptr()
        q0 = _lib.dptr(qkv)
        S = torch.empty(B, nh, T, Tp, dtype=torch.float32, device=dev)
        dP = torch.empty(B, nh, T, Tp, dtype=torch.float32, device=dev)
        Pm = torch.empty(B, nh, T, Tp, dtype=bf, device=dev)
        dS = torch.empty(B, nh, T, Tp, dtype=bf, device=dev)

        def bgemm(A, a_str, W, w_str, D, d_str, out_bf16, M, N, K):
            _lib.call("isp_gemm_bf16_tc_batched", A, *a_str, W, *w_str, D, *d_str, int(out_bf16), M, N, K, nh, B, 1.0, st)

        tok = (3 * C, hd, T * 3 * C)            # (row, head, image) strides of a head slice of qkv
        sq = (Tp, T * Tp, nh * T * Tp)          # ... of the [B, nh, T, Tp] score-shaped tensors
        bgemm(q0, tok, q0 + 2 * C, tok, _lib.dptr(S), sq, False, T, T, hd)                     # S = Q K^T
        _call("isp_softmax_rows", S, Tp, Pm, Tp, B * nh * T, T, Tp)
        bgemm(_lib.dptr(dO), (C, hd, T * C), q0 + 4 * C, tok, _lib.dptr(dP), sq, False, T, T, hd)  # dP = dO V^T
        _call("isp_attn_ds_rows", Pm, Tp, dP, 0, Tp, dS, Tp, B * nh * T, T, Tp)
        del S, dP
        dqkv = torch.empty(B * T, 3 * C, dtype=bf, device=dev)
        d0 = _lib.dptr(dqkv)

        def bg(name, A, a_str, W, w_str, D, d_str, M, N, K):
            _lib.call(name, A, *a_str, W, *w_str, D, *d_str, 1, M, N, K, nh, B, 1.0, st)

        do_tok = (C, hd, T * C)
        # dQ = dS K (K as stored: keys x head_dim); dK = dS^T Q and dV = P^T dO reduce over the queries with both
        # operands as stored (reduction-major): no transposed copies
        bg("isp_gemm_bf16_tc_batched_nn", _lib.dptr(dS), sq, q0 + 2 * C, tok, d0, tok, T, hd, T)
        bg("isp_gemm_bf16_tc_batched_tn", _lib.dptr(dS), sq, q0, tok, d0 + 2 * C, tok, T, hd, T)
        bg("isp_gemm_bf16_tc_batched_tn", _lib.dptr(Pm), sq, _lib.dptr(dO), do_tok, d0 + 4 * C, tok, T, hd, T)
        return dqkv

    def _backward_impl(self, saved, grad_feats):
        dev = grad_feats.device
        m = self.model
        C, nh = m.embed_dim, m.num_heads
        P, PB = self._pack_bwd(dev)
        B, T, h, w = saved["B"], saved["T"], saved["h"], saved["w"]
        M = B * T
        bf = torch.bfloat16
        g = grad_feats.detach().float().permute(0, 2, 3, 1).reshape(B, h * w, C)
        blocks, blocks_b = P["blocks"], PB
        if self.feat_type == "key":  # features = K rows of the last block's qkv projection (channel order d * heads + head)
            hd = C // nh
            dk = torch.zeros(B, T, C, dtype=bf, device=dev)
            dk[:, 1:] = g.view(B, h * w, hd, nh).permute(0, 1, 3, 2).reshape(B, h * w, C).to(bf)
            L, LB = blocks[-1], blocks_b[-1]
            dn = tc.gemm(dk.view(M, C), LB["WqkvT"][:, C:2 * C], out_dtype=torch.float32, N=C, K=C)
            dx, dxb = self._ln_bwd(dn, saved["key_x"], L["n1w"], None, C)
            blocks, blocks_b = blocks[:-1], blocks_b[:-1]
        else:
            dxn = torch.zeros(B, T, C, dtype=torch.float32, device=dev)  # the CLS token is not part of the features
            dxn[:, 1:] = g
            dx, dxb = self._ln_bwd(dxn.view(M, C), saved["x_final"], P["nw"], None, C)
        for L, LB, (x_in, qkv, x_mid) in zip(reversed(blocks), reversed(blocks_b), reversed(saved["blocks"])):
            # MLP: x_out = x_mid + fc2(gelu(fc1(LN2 x_mid)))   (LayerScale folded into fc2)
            hn = _ln(x_mid, L["n2w"], L["n2b"], C, 1e-6, bf)
            pre = tc.gemm(hn, L["W1"], bias=L["b1"], out_dtype=bf)
            dh = tc.gemm(dxb, LB["W2T"], out_dtype=bf)
            _call("isp_gelu_bwd_bf16", dh, pre, dh, dh.numel(), 0)
            dn = tc.gemm(dh, LB["W1T"], out_dtype=torch.float32)
            dx, dxb = self._ln_bwd(dn, x_mid, L["n2w"], dx, C)
            # attention: x_mid = x_in + proj(attn(qkv(LN1 x_in)))
            dO = tc.gemm(dxb, LB["WprojT"], out_dtype=bf)
            dqkv = self._attention_bwd(qkv, dO, B, T, C, nh)
            dn = tc.gemm(dqkv, LB["WqkvT"], out_dtype=torch.float32)
            dx, dxb = self._ln_bwd(dn, x_in, L["n1w"], dx, C)
        return dx.view(B, T, C)[:, 1:].contiguous()


class _TimmBlock(nn.Module):
    def __init__(self, dim, mlp_ratio=4):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = nn.Module()
        self.attn.qkv = nn.Linear(dim, 3 * dim)
        self.attn.proj = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = nn.Module()
        self.mlp.fc1 = nn.Linear(dim, mlp_ratio * dim)
        self.mlp.fc2 = nn.Linear(mlp_ratio * dim, dim)


class _TimmViT(nn.Module):
    """Parameter container with the names of DINO.py's VisionTransformer / timm's vit_small_patch16_224 (minus the
    classifier head, DINO.py:498-504): no LayerScale, no mask token."""

    def __init__(self, dim=384, depth=12, heads=6, patch=16, img_size=224):
        super().__init__()
        self.embed_dim, self.num_heads, self.patch_size = dim, heads, patch
        n = (img_size // patch) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, dim))
        self.patch_embed = nn.Module()
        self.patch_embed.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)
        self.blocks = nn.ModuleList([_TimmBlock(dim) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)


class DINOFeaturizer(DINOv2Featurizer):
    """`type="vit"` / `"dino"` backbone (core/model/featurizers/DINO.py:470-611): ViT-S with 16-pixel patches, features =
    the keys of the last block (`feat_type="key"`, channel order head_dim-major) or the normed tokens ("token").  Same
    constructor as the reference; the reference copies timm / hub weights into its own ViT (network) -- here the ViT is
    built locally with those parameter names (a timm `vit_small_patch16_224` state dict without `head.*` loads with
    `self.model.load_state_dict`) and random-initialised otherwise.  Runs on the DINOv2 code path (same kernels,
    same activation backward)."""

    def __init__(self, arch: str = "vit_small_patch16_224", patch_size: int = 16, feat_type: str = "key",
                 feats_injection_mode: str = "no_injection") -> None:
        nn.Module.__init__(self)
        if feat_type not in ("key", "token"):
            raise ValueError("Unknown feat type:{}".format(feat_type))  # DINO.py:606
        assert feats_injection_mode in ("before_backbone", "after_backbone"), \
            f"Unknown feats_injection_mode: {feats_injection_mode}"       # DINO.py:517-520
        self.arch, self.feat_type, self.feats_injection_mode = arch, feat_type, feats_injection_mode
        self.model = _TimmViT(patch=patch_size)
        self.patch_size = patch_size
        self.n_feats = 384 if arch == "vit_small" else 768  # DINO.py:508-511 (kept as is)
        self._packed = None
        self._pos_cache = {}

    def _forward_impl(self, x, additional_features, saved):
        if (x.shape[2] // self.patch_size) % 2:
            raise NotImplementedError("odd patch-grid heights are cropped by the reference's PatchEmbed (DINO.py:207-208)")
        return super()._forward_impl(x, additional_features, saved)


class _DinoBackboneFn(torch.autograd.Function):
    """Frozen DINOv2 backbone with an input gradient for the injected click embedding."""

    @staticmethod
    def forward(ctx, feat, x, additional_features):
        saved = {}
        out = feat._forward_impl(x, additional_features, saved)
        ctx.feat, ctx.saved = feat, saved
        return out

    @staticmethod
    def backward(ctx, grad_out):
        d_extra = ctx.feat._backward_impl(ctx.saved, grad_out)
        ctx.saved = None
        return None, None, d_extra


class _ClipResBlock(nn.Module):
    """Parameter container with ResidualAttentionBlock's names (maskclip/model.py:225-243)."""

    def __init__(self, width, heads):
        super().__init__()
        self.attn = nn.MultiheadAttention(width, heads)  # container only: in_proj_weight/bias, out_proj.*
        self.ln_1 = nn.LayerNorm(width)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(width, width * 4))
        self.mlp.add_module("gelu", nn.Identity())  # QuickGELU has no parameters
        self.mlp.add_module("c_proj", nn.Linear(width * 4, width))
        self.ln_2 = nn.LayerNorm(width)


class _ClipViT(nn.Module):
    """Parameter container with maskclip VisionTransformer's names (maskclip/model.py:286-319)."""

    def __init__(self, input_resolution=224, patch_size=16, width=768, layers=12, heads=12, output_dim=512):
        super().__init__()
        self.input_resolution, self.output_dim, self.patch_size = input_resolution, output_dim, patch_size
        self.width, self.heads = width, heads
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = nn.Module()
        self.transformer.resblocks = nn.Sequential(*[_ClipResBlock(width, heads) for _ in range(layers)])
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))


_CLIP_ARCHS = {"ViT-B/16": dict(input_resolution=224, patch_size=16, width=768, layers=12, heads=12, output_dim=512)}


class MaskCLIPFeaturizer(nn.Module):
    """Frozen MaskCLIP (CLIP ViT-B/16) dense features on libisp_b200: same constructor and
    `forward(x, additional_features)` as the reference adapter (core/model/featurizers/MaskCLIP.py:14-92).
    11 full blocks, then the last block's value path only (`forward_v`, maskclip/model.py:251-263),
    `ln_post` and `@ proj` on the patch tokens.  The reference downloads the OpenAI weights
    (network); here the parameters are `self.model.visual.*` with CLIP's names, random-init
    unless a state dict is loaded.  bf16 tensor-core GEMMs, fp32 residual stream (the reference
    runs fp16 on CUDA)."""

    def __init__(self, model_name: str = "ViT-B/16", feats_injection_mode: str = "no_injection") -> None:
        super().__init__()
        if model_name not in _CLIP_ARCHS:
            raise NotImplementedError(f"Only {sorted(_CLIP_ARCHS)} are supported, got {model_name}")
        self.feats_injection_mode = feats_injection_mode
        self.model = nn.Module()
        self.model.visual = _ClipViT(**_CLIP_ARCHS[model_name])
        self.patch_size = self.model.visual.patch_size
        self._packed = None
        self._pos_cache = {}

    def _pack(self, dev):
        key = (str(dev), sum(p._version for p in self.parameters()))
        if self._packed is not None and self._packed["key"] == key:
            return self._packed
        v = self.model.visual
        C, nh = v.width, v.heads
        f32 = lambda t: t.detach().float().contiguous().to(dev)
        P = {"key": key}
        P["Wpe"] = tc.pack_linear_weight(v.conv1.weight.detach().float().reshape(C, -1)).to(dev)
        P["cls"] = f32(v.class_embedding)
        P["lpw"], P["lpb"] = f32(v.ln_pre.weight), f32(v.ln_pre.bias)
        sc = (C // nh) ** -0.5
        blocks = []
        for blk in v.transformer.resblocks:
            Wqkv, bqkv = blk.attn.in_proj_weight.detach().float().clone(), blk.attn.in_proj_bias.detach().float().clone()
            Wv, bv = Wqkv[-C:].clone(), bqkv[-C:].clone()
            Wqkv[:C] *= sc
            bqkv[:C] *= sc
            blocks.append({
                "n1w": f32(blk.ln_1.weight), "n1b": f32(blk.ln_1.bias),
                "Wqkv": tc.pack_linear_weight(Wqkv).to(dev), "bqkv": f32(bqkv),
                "Wv": tc.pack_linear_weight(Wv).to(dev), "bv": f32(bv),
                "Wo": tc.pack_linear_weight(blk.attn.out_proj.weight).to(dev), "bo": f32(blk.attn.out_proj.bias),
                "n2w": f32(blk.ln_2.weight), "n2b": f32(blk.ln_2.bias),
                "W1": tc.pack_linear_weight(blk.mlp.c_fc.weight).to(dev), "b1": f32(blk.mlp.c_fc.bias),
                "W2": tc.pack_linear_weight(blk.mlp.c_proj.weight).to(dev), "b2": f32(blk.mlp.c_proj.bias),
            })
        P["blocks"] = blocks
        P["npw"], P["npb"] = f32(v.ln_post.weight), f32(v.ln_post.bias)
        P["Wout"] = tc.pack_linear_weight(v.proj.detach().float().t().contiguous()).to(dev)
        self._packed = P
        self._pos_cache = {}
        return P

    def _pos(self, w_arg, h_arg, n, dev):
        """interpolate_positional_embedding (maskclip/interpolate.py:5-60), called with the (w, h)
        the reference passes; parameter preprocessing, cached per size."""
        v = self.model.visual
        key = (w_arg, h_arg, str(dev), v.positional_embedding._version)
        if key not in self._pos_cache:
            pe = v.positional_embedding.detach().float().cpu()
            n_og = pe.shape[0] - 1
            if not (n == n_og and w_arg == h_arg):
                w0, h0 = w_arg // self.patch_size + 0.1, h_arg // self.patch_size + 0.1
                s = int(math.sqrt(n_og))
                pp = F.interpolate(pe[1:].reshape(1, s, s, -1).permute(0, 3, 1, 2), scale_factor=(w0 / s, h0 / s),
                                   mode="bicubic", align_corners=False, recompute_scale_factor=False)
                pe = torch.cat([pe[:1], pp.permute(0, 2, 3, 1).reshape(-1, pe.shape[-1])], 0)
            self._pos_cache[key] = pe.contiguous().to(dev)
        return self._pos_cache[key]

    def forward(self, x: torch.Tensor, additional_features: torch.Tensor = None) -> torch.Tensor:
        """Frozen backbone; like DINOv2Featurizer the call is differentiable w.r.t. a click embedding injected before
        the blocks (models/sbd/maskclip/patch-embed_noup.py trains it through the frozen CLIP ViT)."""
        if (torch.is_grad_enabled() and additional_features is not None and additional_features.requires_grad
                and self.feats_injection_mode == "before_backbone"):
            return _MaskClipBackboneFn.apply(self, x, additional_features)
        return self._forward_impl(x, additional_features, None)

    def _forward_impl(self, x, additional_features, saved):
        x = x.detach().float()
        dev = x.device
        B, _, H, W = x.shape
        p = self.patch_size
        h, w = H // p, W // p
        v = self.model.visual
        C, nh, Co = v.width, v.heads, v.output_dim
        P = self._pack(dev)
        before = additional_features is not None and self.feats_injection_mode == "before_backbone"
        extra = None
        if before:
            extra = additional_features.detach().float().contiguous()
            assert tuple(extra.shape) == (B, h * w, C), f"x.shape: {(B, h * w, C)}, additional_features.shape: {tuple(extra.shape)}"
        N, T = h * w, h * w + 1
        cols = torch.empty(B * N, P["Wpe"].shape[1], dtype=torch.bfloat16, device=dev)
        _call("isp_vit_patchify", x, *x.stride(), cols, B, 3, H, W, p, P["Wpe"].shape[1])
        patch = tc.gemm(cols, P["Wpe"], out_dtype=torch.float32, K=3 * p * p)
        # VisionTransformer.forward reads `_, _, w, h = x.shape` (w := H, h := W, model.py:321) while
        # forward_without_patch_embed unpacks `h, w = orig_image_hw` (:388): the two entry points hand
        # the interpolation opposite orders; mirrored here (they agree for square images).
        pos = self._pos(W, H, N, dev) if before else self._pos(H, W, N, dev)
        tok = torch.empty(B * T, C, dtype=torch.float32, device=dev)
        _call("isp_vit_assemble_tokens", patch, extra, P["cls"], pos, tok, B, N, C)
        xs = _ln(tok, P["lpw"], P["lpb"], C, 1e-5, torch.float32)
        Tp = tc.round_up(T, 128)
        hd = C // nh
        bf = torch.bfloat16
        if saved is not None:
            saved.update({"B": B, "T": T, "h": h, "w": w, "tok": tok, "blocks": []})
        for L in P["blocks"][:-1]:
            x_in = xs
            hn = _ln(xs, L["n1w"], L["n1b"], C, 1e-5, bf)
            qkv = tc.gemm(hn, L["Wqkv"], bias=L["bqkv"], out_dtype=bf)
            Kp = torch.empty(B, nh, Tp, 64, dtype=bf, device=dev)
            Vt = torch.empty(B, nh, 64, Tp, dtype=bf, device=dev)
            _call("isp_repack_heads", qkv, 1, 3 * C, C, hd, Kp, B, T, Tp, nh, 64, 0)
            _call("isp_repack_heads", qkv, 1, 3 * C, 2 * C, hd, Vt, B, T, Tp, nh, 64, 1)
            O = torch.empty(B * T, C, dtype=bf, device=dev)
            _call("isp_attention_bf16_tc", qkv, 3 * C, hd, Kp, Vt, O, C, hd, B, T, nh, T, 0)
            xs = tc.gemm(O, L["Wo"], bias=L["bo"], resid=xs, out_dtype=torch.float32)
            hn = _ln(xs, L["n2w"], L["n2b"], C, 1e-5, bf)
            h1 = tc.gemm(hn, L["W1"], bias=L["b1"], act="quick_gelu", out_dtype=bf)
            if saved is not None:
                saved["blocks"].append((x_in, qkv, xs))
            xs = tc.gemm(h1, L["W2"], bias=L["b2"], resid=xs, out_dtype=torch.float32)
        L = P["blocks"][-1]  # forward_v: value projection -> output projection, no residual (model.py:251-263)
        hn = _ln(xs, L["n1w"], L["n1b"], C, 1e-5, bf)
        vv = tc.gemm(hn, L["Wv"], bias=L["bv"], out_dtype=bf)
        vo = tc.gemm(vv, L["Wo"], bias=L["bo"], out_dtype=torch.float32)
        if saved is not None:
            saved["x_last"], saved["vo"] = xs, vo
        xn = _ln(vo, P["npw"], P["npb"], C, 1e-5, bf)
        feats = tc.gemm(xn, P["Wout"], out_dtype=torch.float32).view(B, T, Co)[:, 1:]
        if additional_features is not None and self.feats_injection_mode == "after_backbone":
            assert tuple(feats.shape) == tuple(additional_features.shape), \
                f"features.shape: {tuple(feats.shape)}, additional_features.shape: {tuple(additional_features.shape)}"
            feats = feats + additional_features.to(feats.dtype)
        return feats.reshape(B, h, w, Co).permute(0, 3, 1, 2)

    # ---------------------------------------------------------------- activation backward
    def _pack_bwd(self, dev):
        P = self._pack(dev)
        if "bwd" not in P:
            v = self.model.visual
            C, nh = v.width, v.heads
            sc = (C // nh) ** -0.5
            pw = lambda w: tc.pack_linear_weight(w.detach().float().t().contiguous()).to(dev)
            out = []
            for blk in v.transformer.resblocks:
                Wqkv = blk.attn.in_proj_weight.detach().float().clone()
                Wv = Wqkv[-C:].clone()
                Wqkv[:C] *= sc
                out.append({"WqkvT": pw(Wqkv), "WvT": pw(Wv), "WoT": pw(blk.attn.out_proj.weight),
                            "W1T": pw(blk.mlp.c_fc.weight), "W2T": pw(blk.mlp.c_proj.weight)})
            P["bwd"] = out
            P["WoutT"] = tc.pack_linear_weight(v.proj.detach().float().contiguous()).to(dev)  # (proj^T)^T
        return P, P["bwd"]

    def _backward_impl(self, saved, grad_feats):
        dev = grad_feats.device
        v = self.model.visual
        C, nh, Co = v.width, v.heads, v.output_dim
        P, PB = self._pack_bwd(dev)
        B, T, h, w = saved["B"], saved["T"], saved["h"], saved["w"]
        M, bf = B * T, torch.bfloat16
        ln_bwd, attn_bwd = DINOv2Featurizer._ln_bwd, DINOv2Featurizer._attention_bwd
        df = torch.zeros(B, T, Co, dtype=bf, device=dev)  # the class token is not part of the features
        df[:, 1:] = grad_feats.detach().permute(0, 2, 3, 1).reshape(B, h * w, Co).to(bf)
        dn = tc.gemm(df.view(M, Co), P["WoutT"], out_dtype=torch.float32)          # through `@ proj`
        _, dvo = ln_bwd(dn, saved["vo"], P["npw"], None, C, 1e-5)                  # ln_post
        L, LB = P["blocks"][-1], PB[-1]
        dvv = tc.gemm(dvo, LB["WoT"], out_dtype=bf)                                # out_proj of forward_v
        dn = tc.gemm(dvv, LB["WvT"], out_dtype=torch.float32)                      # value projection
        dx, dxb = ln_bwd(dn, saved["x_last"], L["n1w"], None, C, 1e-5)
        for L, LB, (x_in, qkv, x_mid) in zip(reversed(P["blocks"][:-1]), reversed(PB[:-1]), reversed(saved["blocks"])):
            hn = _ln(x_mid, L["n2w"], L["n2b"], C, 1e-5, bf)
            pre = tc.gemm(hn, L["W1"], bias=L["b1"], out_dtype=bf)
            dh = tc.gemm(dxb, LB["W2T"], out_dtype=bf)
            _call("isp_gelu_bwd_bf16", dh, pre, dh, dh.numel(), 1)                 # QuickGELU
            dn = tc.gemm(dh, LB["W1T"], out_dtype=torch.float32)
            dx, dxb = ln_bwd(dn, x_mid, L["n2w"], dx, C, 1e-5)
            dO = tc.gemm(dxb, LB["WoT"], out_dtype=bf)
            dqkv = attn_bwd(None, qkv, dO, B, T, C, nh)
            dn = tc.gemm(dqkv, LB["WqkvT"], out_dtype=torch.float32)
            dx, dxb = ln_bwd(dn, x_in, L["n1w"], dx, C, 1e-5)
        dtok, _ = ln_bwd(dx, saved["tok"], P["lpw"], None, C, 1e-5)               # ln_pre
        return dtok.view(B, T, C)[:, 1:].contiguous()


class _MaskClipBackboneFn(torch.autograd.Function):
    """Frozen MaskCLIP ViT with an input gradient for the injected click embedding."""

    @staticmethod
    def forward(ctx, feat, x, additional_features):
        saved = {}
        out = feat._forward_impl(x, additional_features, saved)
        ctx.feat, ctx.saved = feat, saved
        return out

    @staticmethod
    def backward(ctx, grad_out):
        d = ctx.feat._backward_impl(ctx.saved, grad_out)
        ctx.saved = None
        return None, None, d
