"""ORACLE (test infrastructure) -- FeatUp JBU stack, torch-CPU fp32 restatement.

PARITY UNPINNED.  The arithmetic is NOT in /root/reference: the reference
hub-loads it (`torch.hub.load("mhamilton723/FeatUp", backbone, use_norm).upsampler`,
/root/reference/core/model/upsamplers/JBUFeatUp.py:30-32; requirements.txt:28
installs FeatUp from git HEAD, no pinned version).  This file restates the
published algorithm of upstream `featup/upsamplers.py` (JBUStack,
JBULearnedRange) and `featup/adaptive_conv_cuda/adaptive_conv.{py,cpp,cu}`
(AdaptiveConv) as summarised in SURVEY.md section 3.5.  There is no
reference-side vector to pin it against; the only anchor the reference holds
is the shape contract of JBUFeatUp.py:36-45 ([1,384,14,14] + [1,3,224,224] ->
[1,384,224,224]), which tests/test_oracle_golden.py checks.

State-dict keys mirror upstream JBUStack: up{1..4}.{range_temp, sigma_spatial,
range_proj.{0,3}.{weight,bias}, fixup_proj.{0,3}.{weight,bias}},
fixup_proj.1.{weight,bias}.
"""
import torch
import torch.nn.functional as F

RADIUS = 3
DIAM = 7
KEY_DIM = 32


def init_state_dict(feat_dim, seed=0):
    """Random init with upstream's parameter shapes/defaults (range_temp=0,
    sigma_spatial=1, conv default init), for synthetic benchmarks and tests."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin):
        bound = 1.0 / (cin ** 0.5)  # nn.Conv2d default (kaiming_uniform a=sqrt(5)) for 1x1
        sd[name + ".weight"] = (torch.rand(cout, cin, 1, 1, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    for k in range(1, 5):
        p = f"up{k}"
        sd[p + ".range_temp"] = torch.tensor(0.0)
        sd[p + ".sigma_spatial"] = torch.tensor(1.0)
        conv(p + ".range_proj.0", KEY_DIM, 3)
        conv(p + ".range_proj.3", KEY_DIM, KEY_DIM)
        conv(p + ".fixup_proj.0", DIAM * DIAM, 3 + DIAM * DIAM)
        conv(p + ".fixup_proj.3", DIAM * DIAM, DIAM * DIAM)
    conv("fixup_proj.1", feat_dim, feat_dim)
    return sd


def adaptive_conv(inp_padded, filters):
    """AdaptiveConv.forward: out[b,c,h,w] = sum_{i,j} in[b,c,h+i,w+j] * f[b,h,w,i,j].
    inp_padded [B,C,H+6,W+6], filters [B,H,W,7,7]."""
    B, C, Hp, Wp = inp_padded.shape
    H, W = filters.shape[1], filters.shape[2]
    out = torch.zeros(B, C, H, W, dtype=inp_padded.dtype)
    for i in range(DIAM):
        for j in range(DIAM):
            out += inp_padded[:, :, i:i + H, j:j + W] * filters[:, None, :, :, i, j]
    return out


def adaptive_conv_grad_input(grad_out, filters):
    """AdaptiveConv.backward wrt the padded input:
    gi[b,c,y,x] = sum_{i,j} go[b,c,y-i,x-j] * f[b,y-i,x-j,i,j]."""
    B, C, H, W = grad_out.shape
    gi = torch.zeros(B, C, H + 2 * RADIUS, W + 2 * RADIUS, dtype=grad_out.dtype)
    for i in range(DIAM):
        for j in range(DIAM):
            gi[:, :, i:i + H, j:j + W] += grad_out * filters[:, None, :, :, i, j]
    return gi


def dropout2d_masks(batch, feat_dim, seed=0):
    """Multiplicative Dropout2d masks (0 or 1/(1-p) per sample and channel) for the stack in train() mode, keyed like the
    call sites: up{k}.range (p=0.1, 32 channels), up{k}.fixup (p=0.1, 49 channels), final (p=0.2, feat_dim channels)."""
    g = torch.Generator().manual_seed(seed)

    def m(c, p):
        return (torch.rand(batch, c, generator=g) >= p).float() / (1 - p)

    out = {"final": m(feat_dim, 0.2)}
    for k in range(1, 5):
        out[f"up{k}.range"] = m(KEY_DIM, 0.1)
        out[f"up{k}.fixup"] = m(DIAM * DIAM, 0.1)
    return out


def range_kernel(sd, p, g, masks=None):
    """JBULearnedRange.get_range_kernel: softmax_49(temp * <proj(nbr), proj(centre)>)."""
    proj = F.conv2d(g, sd[p + ".range_proj.0.weight"], sd[p + ".range_proj.0.bias"])
    proj = F.gelu(proj)  # Dropout2d(.1): identity in eval, a per-(sample, channel) mask in train()
    if masks is not None:
        proj = proj * masks[p + ".range"][:, :, None, None]
    proj = F.conv2d(proj, sd[p + ".range_proj.3.weight"], sd[p + ".range_proj.3.bias"])
    B, K, H, W = proj.shape
    pp = F.pad(proj, [RADIUS] * 4, mode="reflect")
    logits = torch.empty(B, DIAM * DIAM, H, W)
    for i in range(DIAM):
        for j in range(DIAM):
            logits[:, i * DIAM + j] = (pp[:, :, i:i + H, j:j + W] * proj).sum(1)
    temp = sd[p + ".range_temp"].exp().clamp_min(1e-4).clamp_max(1e4)
    return F.softmax(temp * logits, dim=1)


def spatial_kernel(sd, p):
    """JBULearnedRange.get_spatial_kernel: exp(-(dx^2+dy^2)/(2 sigma^2)) on linspace(-1,1,7)^2."""
    r = torch.linspace(-1, 1, DIAM)
    d2 = r[:, None] ** 2 + r[None, :] ** 2
    return torch.exp(-d2 / (2 * sd[p + ".sigma_spatial"] ** 2)).reshape(1, DIAM * DIAM, 1, 1)


def combined_kernel(sd, p, g, masks=None):
    k = range_kernel(sd, p, g, masks) * spatial_kernel(sd, p)
    k = k / k.sum(1, keepdim=True).clamp(1e-7)
    h = F.conv2d(torch.cat([k, g], 1), sd[p + ".fixup_proj.0.weight"], sd[p + ".fixup_proj.0.bias"])
    h = F.gelu(h)
    if masks is not None:
        h = h * masks[p + ".fixup"][:, :, None, None]
    h = F.conv2d(h, sd[p + ".fixup_proj.3.weight"], sd[p + ".fixup_proj.3.bias"])
    k = k + 0.1 * h
    B, _, H, W = k.shape
    return k.permute(0, 2, 3, 1).reshape(B, H, W, DIAM, DIAM)


def jbu_stage(sd, p, source, guidance, masks=None):
    """JBUStack.upsample + JBULearnedRange.forward for one x2 stage."""
    _, _, h, w = source.shape
    g = F.adaptive_avg_pool2d(guidance, (2 * h, 2 * w))
    filt = combined_kernel(sd, p, g, masks)
    hr = F.interpolate(source, size=(2 * h, 2 * w), mode="bicubic", align_corners=False)
    hr = F.pad(hr, [RADIUS] * 4, mode="reflect")
    return adaptive_conv(hr, filt)


def jbu_stack_forward(sd, source, guidance, masks=None):
    """JBUStack.forward: 4 stages, then fixup_proj(x)*0.1 + x.  masks=None: eval(); masks = dropout2d_masks(...): train()
    with those Dropout2d draws (what the reference's trainer runs on the frozen stack, trainer.py:213-214)."""
    x = source
    for k in range(1, 5):
        x = jbu_stage(sd, f"up{k}", x, guidance, masks)
    xd = x if masks is None else x * masks["final"][:, :, None, None]
    y = F.conv2d(xd, sd["fixup_proj.1.weight"], sd["fixup_proj.1.bias"])
    return y * 0.1 + x
