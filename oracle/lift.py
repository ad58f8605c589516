"""ORACLE (test infrastructure) -- LiFT x2 conv upsampler, torch-CPU fp32.
Follows /root/reference/core/model/upsamplers/LiFT.py:106-122 (forward),
:12-44 (DoubleConv / Up).  BatchNorm in eval mode (running statistics) or, with train=True, as nn.BatchNorm2d computes it in
train() -- batch statistics, which is how the reference's trainer runs the frozen upsampler (core/training/trainer.py:213-214);
the running statistics of `sd` are then updated in place like the module's buffers.  Keys = LiFT(...).state_dict()."""
import torch
import torch.nn.functional as F


def _bn(x, sd, p, eps=1e-5, train=False):
    if train:  # F.batch_norm(training=True): biased variance normalises, running statistics move with momentum 0.1
        out = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], True, 0.1, eps)
        if p + ".num_batches_tracked" in sd:
            sd[p + ".num_batches_tracked"] += 1
        return out
    s = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + eps)
    return x * s[None, :, None, None] + (sd[p + ".bias"] - sd[p + ".running_mean"] * s)[None, :, None, None]


def _cbr(x, sd, conv, bn, stride, bias=True, train=False):
    x = F.conv2d(x, sd[conv + ".weight"], sd.get(conv + ".bias") if bias else None,
                 stride=stride, padding=1)
    return torch.relu(_bn(x, sd, bn, train=train))


def lift_forward(sd, source, guidance, train=False):
    """LiFTUpsampler.forward(source, guidance) == lift(guidance, source) (LiFT.py:145-146)."""
    h, w = source.shape[2], source.shape[3]
    i1 = _cbr(guidance, sd, "image_convs_1.0", "image_convs_1.1", 2, train=train)       # LiFT.py:69-72
    i1 = _cbr(i1, sd, "image_convs_1.3", "image_convs_1.4", 2, train=train)             # LiFT.py:73-75
    i1 = F.adaptive_max_pool2d(i1, (2 * h, 2 * w))                         # LiFT.py:110
    i2 = _cbr(i1, sd, "image_convs_2.0", "image_convs_2.1", 2, train=train)             # LiFT.py:111
    x = torch.cat([source, i2], 1)                                         # LiFT.py:117
    x = F.conv_transpose2d(x, sd["up1.up.weight"], sd["up1.up.bias"], stride=2)  # LiFT.py:41
    x = torch.cat([x, i1], 1)                                              # LiFT.py:42
    x = _cbr(x, sd, "up1.conv_1.double_conv.0", "up1.conv_1.double_conv.1", 1, bias=False, train=train)
    x = _cbr(x, sd, "up1.conv_1.double_conv.3", "up1.conv_1.double_conv.4", 1, bias=False, train=train)
    return F.conv2d(x, sd["outc.weight"], sd["outc.bias"])                 # LiFT.py:119
