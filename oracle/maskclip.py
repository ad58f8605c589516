"""ORACLE (test infrastructure) -- frozen MaskCLIP (CLIP ViT-B/16) dense features, torch-CPU fp32.
Follows MaskCLIPFeaturizer.forward (/root/reference/core/model/featurizers/MaskCLIP.py:41-92) and
the patch_output path of VisionTransformer (maskclip/model.py:320-360, 370-430): 11 full residual
attention blocks (:225-268, QuickGELU :219-221), the last block's value projection -> out
projection only (`forward_v` :251-263), ln_post on the patch tokens, `@ proj` 768 -> 512;
positional table resized by maskclip/interpolate.py:5-60 (the +0.1 scale-factor variant).
Keys = VisionTransformer(224, 16, 768, 12, 12, 512).state_dict().  The reference runs this in
fp16 on CUDA and fp32 on CPU; the oracle is the fp32 form."""
import math

import torch
import torch.nn.functional as F


def interpolate_positional_embedding(pos, num_patches, patch, w, h):
    """maskclip/interpolate.py:5-60.  pos: [1 + n_og, dim]; returns [1 + num_patches, dim]."""
    n_og = pos.shape[0] - 1
    if num_patches == n_og and w == h:
        return pos
    dim = pos.shape[-1]
    w0, h0 = w // patch, h // patch
    assert w0 * h0 == num_patches
    w0, h0 = w0 + 0.1, h0 + 0.1
    s = int(math.sqrt(n_og))
    pe = F.interpolate(pos[1:].reshape(1, s, s, dim).permute(0, 3, 1, 2), scale_factor=(w0 / s, h0 / s),
                       mode="bicubic", align_corners=False, recompute_scale_factor=False)
    pe = pe.permute(0, 2, 3, 1).reshape(-1, dim)
    return torch.cat([pos[:1], pe], 0)


def _mha(x, sd, p, heads):
    """nn.MultiheadAttention(x, x, x) on [B, T, C] (batch-first restatement of the LND call, model.py:244-249)."""
    B, T, C = x.shape
    qkv = F.linear(x, sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"])
    q, k, v = (t.reshape(B, T, heads, C // heads).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    a = torch.softmax((q * (C // heads) ** -0.5) @ k.transpose(-2, -1), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, T, C)
    return F.linear(o, sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"])


def _block(x, sd, p, heads):
    C = x.shape[-1]
    x = x + _mha(F.layer_norm(x, (C,), sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"]), sd, p, heads)
    h = F.linear(F.layer_norm(x, (C,), sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"]), sd[p + ".mlp.c_fc.weight"],
                 sd[p + ".mlp.c_fc.bias"])
    h = h * torch.sigmoid(1.702 * h)  # QuickGELU
    return x + F.linear(h, sd[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"])


def _forward_v(x, sd, p):
    """Last block, value path only (model.py:251-263)."""
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"])
    v = F.linear(h, sd[p + ".attn.in_proj_weight"][-C:], sd[p + ".attn.in_proj_bias"][-C:])
    return F.linear(v, sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"])


def maskclip_forward(sd, img, coord_emb=None, patch=16, heads=12, layers=12):
    """[B,3,H,W] (+ optional [B, N, width] click embedding injected after conv1, MaskCLIP.py:52-66)
    -> [B, out_dim, H/patch, W/patch]."""
    B, _, H, W = img.shape
    x = F.conv2d(img, sd["conv1.weight"], None, stride=patch)
    x = x.reshape(B, x.shape[1], -1).permute(0, 2, 1)
    if coord_emb is not None:
        x = x + coord_emb
    C = x.shape[-1]
    x = torch.cat([sd["class_embedding"].reshape(1, 1, C).expand(B, -1, -1), x], 1)
    # the reference passes (w, h) = x.shape[2:] of an NCHW tensor, i.e. w := H and h := W (model.py:321,337-339;
    # forward_without_patch_embed unpacks h, w = orig_image_hw and passes w=w, h=h: the OTHER order, :388,401).
    # Both orders agree for square inputs; the injected path (the one the IS model uses) is restated here.
    if coord_emb is not None:
        pos = interpolate_positional_embedding(sd["positional_embedding"], x.shape[1] - 1, patch, w=W, h=H)
    else:
        pos = interpolate_positional_embedding(sd["positional_embedding"], x.shape[1] - 1, patch, w=H, h=W)
    x = x + pos
    x = F.layer_norm(x, (C,), sd["ln_pre.weight"], sd["ln_pre.bias"])
    for i in range(layers - 1):
        x = _block(x, sd, f"transformer.resblocks.{i}", heads)
    x = _forward_v(x, sd, f"transformer.resblocks.{layers - 1}")
    x = F.layer_norm(x[:, 1:], (C,), sd["ln_post.weight"], sd["ln_post.bias"])
    x = x @ sd["proj"]
    return x.reshape(B, H // patch, W // patch, -1).permute(0, 3, 1, 2)
