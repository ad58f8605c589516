"""ORACLE (test infrastructure) -- SimpleViTFeaturizer, the TRAINABLE click embedding of
models/sbd/dinov2/simple-vit_noup.py (`embed_coords` type "simple_vit"), torch-CPU fp32.  Follows
/root/reference/core/model/featurizers/simple_ViT.py: patch embedding Rearrange('b c (h p1) (w p2) -> b (h w) (p1 p2 c)')
-> LayerNorm -> Linear -> LayerNorm (:117-126), fixed 2-D sin-cos position table (:18-28, :128-132), pre-norm
transformer (Attention :42-70 without biases, FeedForward :31-39, final LayerNorm :73-93).  Keys =
SimpleViTFeaturizer(...).state_dict().  Differentiable (the tests take parameter gradients through it).
Pinned by tests/golden/simple_vit_56x84.npz (oracle/make_golden.py)."""
import torch
import torch.nn.functional as F


def posemb_sincos_2d(h, w, dim, temperature=10000):
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    omega = torch.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    y = y.flatten()[:, None] * omega[None, :]
    x = x.flatten()[:, None] * omega[None, :]
    return torch.cat((x.sin(), x.cos(), y.sin(), y.cos()), dim=1).float()


def simple_vit_forward(sd, img, patch=14, heads=8, dim_head=64):
    """[B, C, H, W] -> [B, (H/patch)*(W/patch), dim]."""
    B, Cin, H, W = img.shape
    h, w = H // patch, W // patch
    x = img.reshape(B, Cin, h, patch, w, patch).permute(0, 2, 4, 3, 5, 1).reshape(B, h * w, patch * patch * Cin)
    x = F.layer_norm(x, (x.shape[-1],), sd["to_patch_embedding.1.weight"], sd["to_patch_embedding.1.bias"])
    x = F.linear(x, sd["to_patch_embedding.2.weight"], sd["to_patch_embedding.2.bias"])
    dim = x.shape[-1]
    x = F.layer_norm(x, (dim,), sd["to_patch_embedding.3.weight"], sd["to_patch_embedding.3.bias"])
    x = x + posemb_sincos_2d(h, w, dim)
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    inner = heads * dim_head
    for i in range(depth):
        p = f"transformer.layers.{i}"
        n = F.layer_norm(x, (dim,), sd[p + ".0.norm.weight"], sd[p + ".0.norm.bias"])
        qkv = F.linear(n, sd[p + ".0.to_qkv.weight"]).reshape(B, h * w, 3, heads, dim_head).permute(2, 0, 3, 1, 4)
        a = torch.softmax((qkv[0] @ qkv[1].transpose(-1, -2)) * dim_head ** -0.5, dim=-1)
        o = (a @ qkv[2]).transpose(1, 2).reshape(B, h * w, inner)
        x = F.linear(o, sd[p + ".0.to_out.weight"]) + x
        n = F.layer_norm(x, (dim,), sd[p + ".1.net.0.weight"], sd[p + ".1.net.0.bias"])
        n = F.gelu(F.linear(n, sd[p + ".1.net.1.weight"], sd[p + ".1.net.1.bias"]))
        x = F.linear(n, sd[p + ".1.net.3.weight"], sd[p + ".1.net.3.bias"]) + x
    return F.layer_norm(x, (dim,), sd["transformer.norm.weight"], sd["transformer.norm.bias"])
