"""ORACLE (test infrastructure) -- LoftUp upsampler, torch-CPU fp32 restatement.

Functional (weights come in as a flat dict with the reference's state-dict
keys).  Each step cites the reference lines it follows:
  /root/reference/core/model/upsamplers/loftup/loftup.py  (LoftUp, wrapper)
  /root/reference/core/model/upsamplers/loftup/layers.py  (building blocks)
Pinned against the reference modules by tests/test_oracle_golden.py.
"""
import math

import torch
import torch.nn.functional as F


def channel_layernorm(x, w, b, eps):
    """LayerNorm over the channel dim of an NCHW tensor (layers.py:26-35 uses
    nn.LayerNorm on the permuted tensor; layers.py:53-58 spells it out)."""
    u = x.mean(1, keepdim=True)
    v = (x - u).pow(2).mean(1, keepdim=True)
    return (x - u) / torch.sqrt(v + eps) * w[None, :, None, None] + b[None, :, None, None]


def minmax_scale(img):
    """layers.py:61-71 -- per-channel min/max over batch+space."""
    mn = img.amin(dim=(0, 2, 3), keepdim=True)
    mx = img.amax(dim=(0, 2, 3), keepdim=True)
    return (img - mn) / (mx - mn).clamp_min(1e-4) - 0.5


def fourier_features(x, biases, n_freqs, with_values):
    """layers.py:107-158.  x: [B, c, H, W] (c = 3 colour channels, or anything
    when with_values=False, in which case only its shape is used).  Channel
    order is frequency-major, `f*dm + d`, d = (row, col[, R, G, B]); the bias
    tensor [2, dm, n_freqs] is *reshaped* (not permuted) to [n_freqs, dm]
    (SURVEY Q2)."""
    B, _, H, W = x.shape
    gh = torch.linspace(-1, 1, H)
    gw = torch.linspace(-1, 1, W)
    rows = gh[:, None].expand(H, W)
    cols = gw[None, :].expand(H, W)
    base = torch.stack([rows, cols], 0)[None].expand(B, 2, H, W)
    vals = torch.cat([base, x], 1) if with_values else base
    dm = vals.shape[1]
    freqs = torch.exp(torch.linspace(-2, 10, n_freqs))
    arg = vals[:, None] * freqs[None, :, None, None, None]  # [B, F, dm, H, W]
    b_sin = biases[0].reshape(1, n_freqs, dm, 1, 1)
    b_cos = biases[1].reshape(1, n_freqs, dm, 1, 1)
    s = torch.sin((arg + b_sin).reshape(B, n_freqs * dm, H, W))
    c = torch.cos((arg + b_cos).reshape(B, n_freqs * dm, H, W))
    outs = [s, c] + ([x] if with_values else [])
    return torch.cat(outs, 1)


def batchnorm_eval(x, sd, prefix, eps=1e-5):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    scale = w / torch.sqrt(rv + eps)
    return x * scale[None, :, None, None] + (b - rm * scale)[None, :, None, None]


def batchnorm_train(x, sd, prefix, eps=1e-5, momentum=0.1, new_stats=None):
    """nn.BatchNorm2d.forward in train() -- what the reference's trainer runs on the FROZEN upsampler, because
    `self.net.train()` (core/training/trainer.py:213-214) reaches it: batch statistics over (B, H, W), biased variance for
    the normalisation; the running statistics move by `momentum` towards the batch mean / UNBIASED variance and
    num_batches_tracked increments (returned in `new_stats`, keyed like the state dict)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    n = x.numel() // x.shape[1]
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    if new_stats is not None:
        new_stats[prefix + ".running_mean"] = (1 - momentum) * sd[prefix + ".running_mean"] + momentum * mean
        new_stats[prefix + ".running_var"] = (1 - momentum) * sd[prefix + ".running_var"] + momentum * var * n / max(n - 1, 1)
        new_stats[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    scale = w / torch.sqrt(var + eps)
    return x * scale[None, :, None, None] + (b - mean * scale)[None, :, None, None]


def first_conv(x, sd, p="first_conv", train_stats=None):
    """loftup.py:55-65.  Eval mode: BatchNorm uses running statistics.  train_stats = {} selects train() mode (batch
    statistics) and receives the updated running statistics."""
    bn = batchnorm_eval if train_stats is None else (lambda t, d, k: batchnorm_train(t, d, k, new_stats=train_stats))
    x = channel_layernorm(x, sd[f"{p}.0.norm.weight"], sd[f"{p}.0.norm.bias"], 1e-5)
    x = F.conv2d(x, sd[f"{p}.1.weight"], sd[f"{p}.1.bias"], padding=1)
    x = torch.relu(bn(x, sd, f"{p}.2"))
    x = F.conv2d(x, sd[f"{p}.4.weight"], sd[f"{p}.4.bias"], padding=1)
    return torch.relu(bn(x, sd, f"{p}.5"))


def cross_attention(q_in, kv_in, sd, p, heads=4):
    """layers.py:186-202 + torch multi_head_attention_forward: packed in-proj,
    4 heads of dim//4 (=101), q scaled by 1/sqrt(head_dim), softmax over keys."""
    D = q_in.shape[-1]
    hd = D // heads
    q = F.layer_norm(q_in, (D,), sd[f"{p}.norm_q.weight"], sd[f"{p}.norm_q.bias"], 1e-5)
    kv = F.layer_norm(kv_in, (D,), sd[f"{p}.norm_kv.weight"], sd[f"{p}.norm_kv.bias"], 1e-5)
    Wi, bi = sd[f"{p}.attention.in_proj_weight"], sd[f"{p}.attention.in_proj_bias"]
    Q = F.linear(q, Wi[:D], bi[:D])
    K = F.linear(kv, Wi[D:2 * D], bi[D:2 * D])
    V = F.linear(kv, Wi[2 * D:], bi[2 * D:])
    B, Nq, _ = Q.shape
    Nk = K.shape[1]
    Q = Q.view(B, Nq, heads, hd).transpose(1, 2) * (1.0 / math.sqrt(hd))
    K = K.view(B, Nk, heads, hd).transpose(1, 2)
    V = V.view(B, Nk, heads, hd).transpose(1, 2)
    out = torch.empty(B, heads, Nq, hd)
    step = 16384  # chunk the queries: the full [B*4, HW, hw] matrix is 3.3 GB/img at 448^2
    for s in range(0, Nq, step):
        P = torch.softmax(Q[:, :, s:s + step] @ K.transpose(-1, -2), dim=-1)
        out[:, :, s:s + step] = P @ V
    out = out.transpose(1, 2).reshape(B, Nq, D)
    return F.linear(out, sd[f"{p}.attention.out_proj.weight"], sd[f"{p}.attention.out_proj.bias"])


def feed_forward(x, sd, p):
    """layers.py:161-174: LN -> Linear -> exact GELU -> Linear."""
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), sd[f"{p}.net.0.weight"], sd[f"{p}.net.0.bias"], 1e-5)
    h = F.gelu(F.linear(h, sd[f"{p}.net.1.weight"], sd[f"{p}.net.1.bias"]))
    return F.linear(h, sd[f"{p}.net.4.weight"], sd[f"{p}.net.4.bias"])


def ca_transformer(q, kv, sd, p="ca_transformer", depth=2):
    """layers.py:222-228."""
    for l in range(depth):
        q = cross_attention(q, kv, sd, f"{p}.layers.{l}.0") + q
        q = feed_forward(q, sd, f"{p}.layers.{l}.1") + q
    D = q.shape[-1]
    return F.layer_norm(q, (D,), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-5)


def queries(img, sd, train_stats=None):
    """loftup.py:102-110: Fourier features of the min-max-scaled image -> first_conv
    -> [B, HW, D]."""
    x = fourier_features(minmax_scale(img), sd["fourier_feat.1.biases"], 20, True)
    x = first_conv(x, sd, train_stats=train_stats)
    return x.flatten(2).permute(0, 2, 1)


def keys_values(lr, sd):
    """loftup.py:113-118 (lr_pe_type == 'sine')."""
    pe = fourier_features(lr, sd["lr_pe.biases"], 5, False)
    return torch.cat([lr, pe], 1).flatten(2).permute(0, 2, 1)


def loftup_forward(sd, lr_feats, img, cn_weight=None, cn_bias=None, train_stats=None):
    """LoftUp.forward (loftup.py:100-138), optionally preceded by the wrapper's
    ChannelNorm on the LR features (loftup.py:141-149).  sd keys are those of
    `LoftUp(dim).state_dict()`.  train_stats = {}: the module in train() mode (see batchnorm_train)."""
    if cn_weight is not None:
        lr_feats = channel_layernorm(lr_feats, cn_weight, cn_bias, 1e-5)
    B, _, H, W = img.shape
    q = queries(img, sd, train_stats)
    kv = keys_values(lr_feats, sd)
    x = ca_transformer(q, kv, sd)
    x = x.permute(0, 2, 1).reshape(B, -1, H, W)
    x = F.conv2d(x, sd["final_conv.0.weight"], sd["final_conv.0.bias"])
    return channel_layernorm(x, sd["final_conv.1.weight"], sd["final_conv.1.bias"], 1e-6)
