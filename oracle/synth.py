"""ORACLE (test infrastructure) -- deterministic synthetic weights and inputs.

The reference ships no checkpoints we can fetch (no network), so parity runs on
random-init weights (BASELINE.json north_star).  These builders produce state
dicts with exactly the reference's key names and shapes
from a seeded CPU generator, so the authoring container (reference modules,
`make_golden.py`) and the GPU box (our CUDA modules + the oracle) see identical
weights without shipping them.
"""
import torch


class _Gen:
    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)

    def randn(self, *shape, std=1.0):
        return torch.randn(*shape, generator=self.g) * std

    def rand(self, *shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=self.g) * (hi - lo) + lo


def _linear(sd, G, name, cout, cin, bias=True, wkey="weight"):
    sd[f"{name}.{wkey}"] = G.randn(cout, cin, std=cin ** -0.5)
    if bias:
        sd[f"{name}.bias"] = G.randn(cout, std=0.1)


def _conv(sd, G, name, cout, cin, k, bias=True):
    sd[f"{name}.weight"] = G.randn(cout, cin, k, k, std=(cin * k * k) ** -0.5)
    if bias:
        sd[f"{name}.bias"] = G.randn(cout, std=0.1)


def _ln(sd, G, name, dim):
    sd[f"{name}.weight"] = 1.0 + G.randn(dim, std=0.1)
    sd[f"{name}.bias"] = G.randn(dim, std=0.1)


def _bn(sd, G, name, dim):
    _ln(sd, G, name, dim)
    sd[f"{name}.running_mean"] = G.randn(dim, std=0.1)
    sd[f"{name}.running_var"] = G.rand(dim, lo=0.5, hi=1.5)
    sd[f"{name}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def loftup_state_dict(dim=384, seed=0):
    """Keys/shapes of LoftUp(dim, lr_pe_type='sine').state_dict()
    (/root/reference/core/model/upsamplers/loftup/loftup.py:21-79)."""
    G, sd = _Gen(seed), {}
    D = dim + 20
    sd["lr_pe.biases"] = G.randn(2, 2, 5)
    sd["fourier_feat.1.biases"] = G.randn(2, 5, 20)
    _ln(sd, G, "first_conv.0.norm", 203)
    _conv(sd, G, "first_conv.1", D, 203, 3)
    _bn(sd, G, "first_conv.2", D)
    _conv(sd, G, "first_conv.4", D, D, 3)
    _bn(sd, G, "first_conv.5", D)
    for l in range(2):
        p = f"ca_transformer.layers.{l}"
        _ln(sd, G, f"{p}.0.norm_q", D)
        _ln(sd, G, f"{p}.0.norm_kv", D)
        sd[f"{p}.0.attention.in_proj_weight"] = G.randn(3 * D, D, std=D ** -0.5)
        sd[f"{p}.0.attention.in_proj_bias"] = G.randn(3 * D, std=0.1)
        _linear(sd, G, f"{p}.0.attention.out_proj", D, D)
        _ln(sd, G, f"{p}.1.net.0", D)
        _linear(sd, G, f"{p}.1.net.1", dim, D)
        _linear(sd, G, f"{p}.1.net.4", D, dim)
    _ln(sd, G, "ca_transformer.norm", D)
    _conv(sd, G, "final_conv.0", dim, D, 1)
    _ln(sd, G, "final_conv.1", dim)
    return sd


def channelnorm_state_dict(dim=384, seed=1):
    G, sd = _Gen(seed), {}
    _ln(sd, G, "norm", dim)
    return sd


def lift_state_dict(dim=384, seed=0):
    """Keys/shapes of LiFT(dim, 14).state_dict() (LiFT.py:47-90)."""
    G, sd = _Gen(seed), {}
    sd["up1.up.weight"] = G.randn(dim + 32, (dim + 32) // 2, 2, 2, std=(dim + 32) ** -0.5)
    sd["up1.up.bias"] = G.randn((dim + 32) // 2, std=0.1)
    cin = (dim + 32) // 2 + 32
    _conv(sd, G, "up1.conv_1.double_conv.0", dim // 2, cin, 3, bias=False)
    _bn(sd, G, "up1.conv_1.double_conv.1", dim // 2)
    _conv(sd, G, "up1.conv_1.double_conv.3", dim // 2, dim // 2, 3, bias=False)
    _bn(sd, G, "up1.conv_1.double_conv.4", dim // 2)
    _conv(sd, G, "outc", dim, dim // 2, 1)
    _conv(sd, G, "image_convs_1.0", 32, 3, 3)
    _bn(sd, G, "image_convs_1.1", 32)
    _conv(sd, G, "image_convs_1.3", 32, 32, 3)
    _bn(sd, G, "image_convs_1.4", 32)
    _conv(sd, G, "image_convs_2.0", 32, 32, 3)
    _bn(sd, G, "image_convs_2.1", 32)
    return sd


def convhead_state_dict(dim=384, num_layers=2, num_classes=1, seed=0):
    """Keys/shapes of ConvSegHead(dim, num_layers, num_classes).state_dict()."""
    G, sd = _Gen(seed), {}
    for i in range(num_layers):
        _conv(sd, G, f"convs.{i}.conv", dim, dim, 3)
    _conv(sd, G, "classifier", num_classes, dim, 1)
    return sd


def patch_embed_state_dict(dim=384, patch=14, in_chans=3, seed=0):
    G, sd = _Gen(seed), {}
    _conv(sd, G, "proj", dim, in_chans, patch)
    return sd


def vit_state_dict(dim=384, depth=12, patch=14, n_pos=37 * 37 + 1, mlp_ratio=4, seed=0):
    """Keys/shapes of vit_small(patch_size=14, img_size=518, init_values=1.0,
    block_chunks=0).state_dict() (DINOv2.py:53-160)."""
    G, sd = _Gen(seed), {}
    sd["cls_token"] = G.randn(1, 1, dim, std=0.02)
    sd["pos_embed"] = G.randn(1, n_pos, dim, std=0.02)
    sd["mask_token"] = torch.zeros(1, dim)
    _conv(sd, G, "patch_embed.proj", dim, 3, patch)
    for i in range(depth):
        p = f"blocks.{i}"
        _ln(sd, G, f"{p}.norm1", dim)
        _linear(sd, G, f"{p}.attn.qkv", 3 * dim, dim)
        _linear(sd, G, f"{p}.attn.proj", dim, dim)
        sd[f"{p}.ls1.gamma"] = 0.5 + G.rand(dim)
        _ln(sd, G, f"{p}.norm2", dim)
        _linear(sd, G, f"{p}.mlp.fc1", mlp_ratio * dim, dim)
        _linear(sd, G, f"{p}.mlp.fc2", dim, mlp_ratio * dim)
        sd[f"{p}.ls2.gamma"] = 0.5 + G.rand(dim)
    _ln(sd, G, "norm", dim)
    return sd


def dino_vit_state_dict(dim=384, depth=12, patch=16, n_pos=14 * 14 + 1, mlp_ratio=4, seed=0):
    """Keys/shapes of DINO.py vit_small(patch_size=16, num_classes=0).state_dict() (DINO.py:213-287, 394-405)."""
    G, sd = _Gen(seed), {}
    sd["cls_token"] = G.randn(1, 1, dim, std=0.02)
    sd["pos_embed"] = G.randn(1, n_pos, dim, std=0.02)
    _conv(sd, G, "patch_embed.proj", dim, 3, patch)
    for i in range(depth):
        p = f"blocks.{i}"
        _ln(sd, G, f"{p}.norm1", dim)
        _linear(sd, G, f"{p}.attn.qkv", 3 * dim, dim)
        _linear(sd, G, f"{p}.attn.proj", dim, dim)
        _ln(sd, G, f"{p}.norm2", dim)
        _linear(sd, G, f"{p}.mlp.fc1", mlp_ratio * dim, dim)
        _linear(sd, G, f"{p}.mlp.fc2", dim, mlp_ratio * dim)
    _ln(sd, G, "norm", dim)
    return sd


def simple_vit_state_dict(dim=384, depth=6, heads=8, dim_head=64, mlp_dim=2048, patch=14, channels=3, seed=0):
    """Keys/shapes of SimpleViTFeaturizer(...).state_dict() (simple_ViT.py:96-138)."""
    G, sd = _Gen(seed), {}
    pd, inner = channels * patch * patch, heads * dim_head
    _ln(sd, G, "to_patch_embedding.1", pd)
    _linear(sd, G, "to_patch_embedding.2", dim, pd)
    _ln(sd, G, "to_patch_embedding.3", dim)
    _ln(sd, G, "transformer.norm", dim)
    for i in range(depth):
        p = f"transformer.layers.{i}"
        _ln(sd, G, f"{p}.0.norm", dim)
        sd[f"{p}.0.to_qkv.weight"] = G.randn(3 * inner, dim, std=dim ** -0.5)
        sd[f"{p}.0.to_out.weight"] = G.randn(dim, inner, std=inner ** -0.5)
        _ln(sd, G, f"{p}.1.net.0", dim)
        _linear(sd, G, f"{p}.1.net.1", mlp_dim, dim)
        _linear(sd, G, f"{p}.1.net.3", dim, mlp_dim)
    return sd


def maskclip_state_dict(width=768, layers=12, patch=16, out_dim=512, resolution=224, seed=0):
    """Keys/shapes of maskclip.model.VisionTransformer(224, 16, 768, 12, 12, 512).state_dict()
    (/root/reference/core/model/featurizers/maskclip/model.py:286-319)."""
    G, sd = _Gen(seed), {}
    sd["class_embedding"] = G.randn(width, std=width ** -0.5)
    sd["positional_embedding"] = G.randn((resolution // patch) ** 2 + 1, width, std=width ** -0.5)
    sd["proj"] = G.randn(width, out_dim, std=width ** -0.5)
    _conv(sd, G, "conv1", width, 3, patch, bias=False)
    _ln(sd, G, "ln_pre", width)
    for i in range(layers):
        p = f"transformer.resblocks.{i}"
        sd[f"{p}.attn.in_proj_weight"] = G.randn(3 * width, width, std=width ** -0.5)
        sd[f"{p}.attn.in_proj_bias"] = G.randn(3 * width, std=0.1)
        _linear(sd, G, f"{p}.attn.out_proj", width, width)
        _ln(sd, G, f"{p}.ln_1", width)
        _linear(sd, G, f"{p}.mlp.c_fc", 4 * width, width)
        _linear(sd, G, f"{p}.mlp.c_proj", width, 4 * width)
        _ln(sd, G, f"{p}.ln_2", width)
    _ln(sd, G, "ln_post", width)
    return sd


def image_batch(b, h, w, seed=1):
    """Synthetic RGB images in [0,1] (north_star: identical synthetic inputs)."""
    return _Gen(seed).rand(b, 3, h, w)


def lr_features(b, c, h, w, seed=2):
    return _Gen(seed).randn(b, c, h, w)


def click_points(b, p, h, w, seed=3, frac=False, n_valid=None):
    """[b, 2p, 3] float32 (row, col, order); first p positive, last p negative;
    padding rows are (-1,-1,-1) (iseg_base_model.py / base_predictor.py:194-225)."""
    G = _Gen(seed)
    pts = torch.full((b, 2 * p, 3), -1.0)
    for bi in range(b):
        for s in range(2):
            n = int(torch.randint(0 if s else 1, p + 1, (1,), generator=G.g)) if n_valid is None else n_valid
            for k in range(n):
                r = float(G.rand(1) * (h - 1))
                c = float(G.rand(1) * (w - 1))
                if not frac:
                    r, c = float(round(r)), float(round(c))
                pts[bi, s * p + k] = torch.tensor([r, c, float(k)])
    return pts
