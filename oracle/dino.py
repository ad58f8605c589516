"""ORACLE (test infrastructure) -- frozen DINO / timm ViT-S/16 adapter (`type="vit"` backbone of
models/sbd/vit/patch-embed_noup.py), torch-CPU fp32.  Follows DINOFeaturizer.forward
(/root/reference/core/model/featurizers/DINO.py:529-611) on the vendored VisionTransformer (:107-335): click embedding
added to the patch tokens before the blocks (:545-552), positional encoding interpolated with the +0.1 scale-factor
quirk and called with (w, h) = (image rows, image columns) (:289-313, :558), features = the KEYS of the last block
without the class token, channels ordered head_dim-major / head-minor (`permute(0, 2, 3, 1).flatten(-2)`, :591-592) for
feat_type 'key', or norm(x) tokens for 'token' (:579-589).  Keys = vit_small(patch_size=16).state_dict() (qkv bias,
LayerNorm eps 1e-6, no LayerScale).  Pinned by tests/golden/dino_vit_64x96.npz (oracle/make_golden.py)."""
import torch
import torch.nn.functional as F

from .vit import interpolate_pos_encoding


def _block(x, sd, p, heads, eps=1e-6):
    B, T, C = x.shape
    h = F.layer_norm(x, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    qkv = F.linear(h, sd[p + ".attn.qkv.weight"], sd[p + ".attn.qkv.bias"])
    qkv = qkv.reshape(B, T, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = torch.softmax((q @ k.transpose(-2, -1)) * (C // heads) ** -0.5, dim=-1)
    h = (a @ v).transpose(1, 2).reshape(B, T, C)
    x = x + F.linear(h, sd[p + ".attn.proj.weight"], sd[p + ".attn.proj.bias"])
    h = F.layer_norm(x, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    h = F.linear(F.gelu(F.linear(h, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])),
                 sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
    return x + h, k


def dino_vit_forward(sd, img, coord_emb=None, patch=16, heads=6, depth=12, feat_type="key"):
    """[B,3,H,W] (+ [B, N, C] click embedding) -> [B, C, H/patch, W/patch]."""
    B, _, H, W = img.shape
    x = F.conv2d(img, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch)
    assert x.shape[-2] % 2 == 0, "odd patch grids are cropped by the reference's PatchEmbed (DINO.py:207-208)"
    x = x.flatten(2).transpose(1, 2)
    if coord_emb is not None:
        x = x + coord_emb
    C = x.shape[-1]
    x = torch.cat([sd["cls_token"].expand(B, -1, -1), x], 1)
    x = x + interpolate_pos_encoding(sd["pos_embed"], x.shape[1] - 1, H, W, patch)
    k = None
    for i in range(depth):
        x, k = _block(x, sd, f"blocks.{i}", heads)
    if feat_type == "token":
        f = F.layer_norm(x, (C,), sd["norm.weight"], sd["norm.bias"], 1e-6)[:, 1:]
    else:
        f = k[:, :, 1:, :].permute(0, 2, 3, 1).flatten(-2)  # [B, N, head_dim * heads]
    return f.reshape(B, H // patch, W // patch, C).permute(0, 3, 1, 2)
