"""ORACLE (test infrastructure) -- IS head + click-map patch embedding, torch-CPU fp32.
ConvSegHead: /root/reference/core/model/heads/conv_heads.py:48-73 (+ mmcv ConvModule
defaults: Conv2d(bias) -> ReLU); PatchEmbed:
/root/reference/core/model/featurizers/utils/patch_embed.py:36-42."""
import torch
import torch.nn.functional as F


def convhead_forward(sd, x, num_layers=2):
    for i in range(num_layers):
        x = torch.relu(F.conv2d(x, sd[f"convs.{i}.conv.weight"], sd[f"convs.{i}.conv.bias"], padding=1))
    return F.conv2d(x, sd["classifier.weight"], sd["classifier.bias"])


def patch_embed_forward(sd, x):
    p = sd["proj.weight"].shape[-1]
    y = F.conv2d(x, sd["proj.weight"], sd["proj.bias"], stride=p)
    return y.flatten(2).transpose(1, 2)


def bilinear_align_corners(x, size):
    """iseg_probe_model.py:120-129 / iseg_base_model.py:75-80."""
    return F.interpolate(x, size=size, mode="bilinear", align_corners=True)


def normalize_image(img, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """BatchImageNormalize (core/model/ops.py:96-105)."""
    m = torch.tensor(mean)[None, :, None, None]
    s = torch.tensor(std)[None, :, None, None]
    return (img - m) / s
