"""ORACLE (test infrastructure) -- frozen DINOv2 ViT-S/14 with click-embedding
injection before the blocks, torch-CPU fp32.
Follows DINOv2Featurizer.forward 'before_backbone' branch
(/root/reference/core/model/featurizers/DINOv2.py:518-546), the vendored ViT
(`interpolate_pos_encoding` :199-230, blocks dinov2/layers/block.py:92-117,
attention.py:54-71, mlp.py:34-40, layer_scale.py:25-26).  Keys = vit_small().state_dict()."""
import math

import torch
import torch.nn.functional as F


def interpolate_pos_encoding(pos_embed, npatch, w, h, patch):
    N = pos_embed.shape[1] - 1
    if npatch == N and w == h:
        return pos_embed
    dim = pos_embed.shape[-1]
    cls_pe, patch_pe = pos_embed[:, 0], pos_embed[:, 1:]
    w0, h0 = w // patch + 0.1, h // patch + 0.1  # DINOv2.py:213
    s = int(math.sqrt(N))
    pe = F.interpolate(patch_pe.reshape(1, s, s, dim).permute(0, 3, 1, 2),
                       scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
    pe = pe.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat([cls_pe.unsqueeze(0), pe], 1)


def vit_block(x, sd, p, heads, eps=1e-6):
    B, T, C = x.shape
    h = F.layer_norm(x, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    qkv = F.linear(h, sd[p + ".attn.qkv.weight"], sd[p + ".attn.qkv.bias"])
    qkv = qkv.reshape(B, T, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
    a = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
    h = (a @ v).transpose(1, 2).reshape(B, T, C)
    h = F.linear(h, sd[p + ".attn.proj.weight"], sd[p + ".attn.proj.bias"])
    x = x + h * sd[p + ".ls1.gamma"]
    h = F.layer_norm(x, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    h = F.gelu(F.linear(h, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"]))
    h = F.linear(h, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
    return x + h * sd[p + ".ls2.gamma"]


def dinov2_forward(sd, img, coord_emb=None, patch=14, heads=6, depth=12):
    """Returns [B, C, h, w] patch features (DINOv2.py:518-546)."""
    B, _, H, W = img.shape
    x = F.conv2d(img, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch)
    x = x.flatten(2).transpose(1, 2)
    if coord_emb is not None:
        x = x + coord_emb  # DINOv2.py:523
    C = x.shape[-1]
    x = torch.cat([sd["cls_token"].expand(B, -1, -1), x], 1)
    x = x + interpolate_pos_encoding(sd["pos_embed"], x.shape[1] - 1, H, W, patch)
    for i in range(depth):
        x = vit_block(x, sd, f"blocks.{i}", heads)
    x = F.layer_norm(x, (C,), sd["norm.weight"], sd["norm.bias"], 1e-6)
    return x[:, 1:].reshape(B, H // patch, W // patch, C).permute(0, 3, 1, 2)
