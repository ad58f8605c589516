"""Host-side algebra behind the JBU path's restructuring (isegprobe_b200/upsamplers.py `_mix_channels`), checked with the
ORACLE alone on the CPU: JBUStack's final `fixup_proj(x) * 0.1 + x` (a per-pixel channel map) commutes with the four
JBU stages and with the align_corners resize that follows the upsampler (every one of them applies one spatial linear
map to all channels alike), in eval() and with the Dropout2d masks of train().  Also the split-bf16 arithmetic the
channel map runs in on the tensor cores."""
import torch
import torch.nn.functional as F

from oracle import jbu as ojbu
from oracle import synth


def _stack_without_final(sd, src, gd, masks):
    x = src
    for k in range(1, 5):
        x = ojbu.jbu_stage(sd, f"up{k}", x, gd, masks)
    return x


def _mixed_source(sd, src, masks):
    W = sd["fixup_proj.1.weight"].reshape(src.shape[1], src.shape[1]).double()
    xs = src.double() if masks is None else src.double() * masks["final"][:, :, None, None].double()
    return src.double() + 0.1 * torch.einsum("oc,bchw->bohw", W, xs)


def test_final_channel_map_commutes_with_the_stack_and_the_resize():
    C = 16
    sd = {k: v.double() for k, v in ojbu.init_state_dict(C, seed=0).items()}
    src = synth.lr_features(2, C, 3, 4, seed=2).double()
    gd = ((synth.image_batch(2, 48, 64, seed=1) - 0.45) / 0.225).double()
    for masks in (None, {k: v.double() for k, v in ojbu.dropout2d_masks(2, C, seed=7).items()}):
        want = F.interpolate(ojbu.jbu_stack_forward(sd, src, gd, masks), size=(41, 57), mode="bilinear", align_corners=True)
        got = F.interpolate(_stack_without_final(sd, _mixed_source(sd, src, masks), gd, masks), size=(41, 57), mode="bilinear",
                            align_corners=True) + 0.1 * sd["fixup_proj.1.bias"].view(1, C, 1, 1)
        err = float((got - want).abs().max() / want.abs().max())
        assert err < 1e-12, err  # float64: the identity is exact up to rounding


def test_split_bf16_products_are_fp32_accurate():
    from isegprobe_b200.upsamplers import JBUFeatUpUpsampler as J
    g = torch.Generator().manual_seed(0)
    x, W = torch.randn(64, 48, generator=g), torch.randn(48, 48, generator=g) * 0.1
    a, w = J._split_rows(x).float(), J._split_weight(W).float()  # [hi | lo | hi], [hi | hi | lo]
    assert a.shape == (64, 144) and w.shape == (48, 144)
    got = x.double() + 0.1 * (a.double() @ w.double().T)
    want = x.double() + 0.1 * (x.double() @ W.double().T)
    assert float((got - want).abs().max() / want.abs().max()) < 2e-6  # only the lo * lo term (2^-18 relative) is dropped
