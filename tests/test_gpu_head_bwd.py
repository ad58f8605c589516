"""GPU parity: IS head backward (tcgen05 wgrad / dgrad, fused classifier backward) vs torch
autograd through the oracle head (oracle/head.py) on identical seeded weights and inputs.
bf16 tensor-core mode: cosine >= 0.999 per gradient tensor."""
import pytest
import torch
import torch.nn.functional as F

from oracle import head as ohead
from oracle import synth
from tests.gpu_util import DEV, cosine, relerr

pytestmark = pytest.mark.gpu


def _call(name, *a):
    from isegprobe_b200 import _lib
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 64, 64, 64), (2, 12, 40, 128, 192), (1, 16, 100, 384, 384)])
def test_conv3x3_wgrad(B, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(Cin + W)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16)
    dy = torch.randn(B, H, W, Cout, generator=g).to(torch.bfloat16)
    dW = torch.zeros(Cout, 9, Cin, device=DEV)
    _call("isp_conv3x3_wgrad_bf16_tc", x.to(DEV), Cin, dy.to(DEV), Cout, dW, B, H, W, Cin, Cout)
    # reference: gradient of conv2d wrt its weight
    w = torch.zeros(Cout, Cin, 3, 3, requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), w, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    want = w.grad.permute(0, 2, 3, 1).reshape(Cout, 9, Cin)
    assert relerr(dW, want) < 2e-3, relerr(dW, want)
    # accumulation semantics: a second call doubles the result
    _call("isp_conv3x3_wgrad_bf16_tc", x.to(DEV), Cin, dy.to(DEV), Cout, dW, B, H, W, Cin, Cout)
    assert relerr(dW, 2 * want) < 2e-3


def test_conv3x3_dgrad_with_relu_mask():
    from isegprobe_b200 import tc
    g = torch.Generator().manual_seed(3)
    B, H, W, C = 2, 16, 32, 128
    w = (torch.randn(C, C, 3, 3, generator=g) * (9 * C) ** -0.5).to(torch.bfloat16).float()
    dy = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16)
    act = torch.relu(torch.randn(B, H, W, C, generator=g)).to(torch.bfloat16)
    wT = tc.pack_conv3x3_weight(w.flip(2, 3).transpose(0, 1)).to(DEV)
    dx = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=DEV)
    _call("isp_conv3x3_dgrad_bf16_tc", dy.to(DEV), wT, act.to(DEV), C, dx, 1, B, H, W, C, C, C, C)
    want = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1) * (act.float() > 0)
    assert relerr(dx.float(), want) < 1e-2 and cosine(dx.float(), want) > 0.9999
    dx32 = torch.empty(B, H, W, C, dtype=torch.float32, device=DEV)
    _call("isp_conv3x3_dgrad_bf16_tc", dy.to(DEV), wT, None, 0, dx32, 0, B, H, W, C, C, C, C)
    want = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
    assert relerr(dx32, want) < 1e-4


@pytest.mark.parametrize("B,H,W", [(2, 24, 40), (1, 64, 64)])
def test_convhead_backward_matches_autograd(B, H, W):
    import isegprobe_b200 as isp
    C = 384
    sd = synth.convhead_state_dict(C, 2, 1, seed=0)
    head = isp.ConvSegHead(C, 2, 1)
    head.load_state_dict(sd)
    head = head.to(DEV).train()
    x = synth.lr_features(B, C, H, W, seed=4)
    gout = synth.lr_features(B, 1, H, W, seed=5)
    xg = x.to(DEV).requires_grad_(True)
    out = head(xg)
    out.backward(gout.to(DEV))
    # oracle: fp32 autograd on the CPU
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    want = ohead.convhead_forward(ref, xr)
    want.backward(gout)
    assert cosine(out.detach(), want.detach()) > 0.999
    # (1) Against the fp32 oracle.  Where the bf16 forward rounds a near-zero pre-activation to the other
    # side of 0 than fp32 does, the ReLU mask flips and that element's (not small) gradient appears /
    # vanishes -- a few 1e-3 of the elements per masked layer -- hence 0.995 for everything below a mask.
    for name, p in head.named_parameters():
        c = cosine(p.grad, ref[name].grad)
        assert c > (0.999 if name.startswith("classifier") else 0.995), (name, c)
    assert cosine(xg.grad, xr.grad) > 0.995, cosine(xg.grad, xr.grad)
    # (2) Against fp32 autograd through a model that rounds where ours rounds (bf16 input, weights and
    # stored activations, fp32 accumulation; straight-through rounding): same masks, so the kernels'
    # own error is what is left -- cosine >= 0.999 on every tensor.
    bf = lambda t: t.to(torch.bfloat16).float()
    emu = {k: v.clone().to(DEV).requires_grad_(True) for k, v in sd.items()}
    xe = x.clone().to(DEV).requires_grad_(True)
    f = xe + (bf(xe) - xe).detach()
    for i in range(2):
        wq = emu[f"convs.{i}.conv.weight"]
        a = torch.relu(F.conv2d(f, wq + (bf(wq) - wq).detach(), emu[f"convs.{i}.conv.bias"], padding=1))
        f = a + (bf(a) - a).detach()
    F.conv2d(f, emu["classifier.weight"], emu["classifier.bias"]).backward(gout.to(DEV))
    for name, p in head.named_parameters():
        c = cosine(p.grad, emu[name].grad)
        assert c > 0.999, ("emulated", name, c)
    assert cosine(xg.grad, xe.grad) > 0.999, cosine(xg.grad, xe.grad)
    # a second backward accumulates into .grad like any torch module
    g1 = head.classifier.weight.grad.clone()
    head(xg).backward(gout.to(DEV))
    assert relerr(head.classifier.weight.grad, 2 * g1) < 1e-3


def test_convhead_inference_path_unchanged_under_no_grad():
    import isegprobe_b200 as isp
    head = isp.ConvSegHead(384, 2, 1).to(DEV).eval()
    x = synth.lr_features(1, 384, 16, 16, seed=1).to(DEV)
    with torch.no_grad():
        a = head(x)
    for p in head.parameters():
        p.requires_grad = False
    b = head(x)
    assert torch.equal(a, b)


def test_head_trainer_step_reduces_loss():
    """HeadTrainer (frozen features -> head fwd/bwd -> flat-arena all-reduce -> Adam): on a fixed batch
    the loss must go down, and only the head's parameters may change."""
    import isegprobe_b200 as isp
    from isegprobe_b200.training import HeadTrainer
    torch.manual_seed(0)
    pipe = isp.ISegPipeline("bilinear", {}).to(DEV)
    pipe.embed_coords = isp.PatchEmbed((112, 112), (14, 14), 3, 384).to(DEV)
    tr = HeadTrainer(pipe, lr=5e-5)  # models/defaults.py:105
    img = torch.cat([synth.image_batch(2, 112, 112, seed=1), torch.zeros(2, 1, 112, 112)], 1).to(DEV)
    pts = synth.click_points(2, 3, 112, 112, seed=3).to(DEV)
    yy, xx = torch.meshgrid(torch.arange(112), torch.arange(112), indexing="ij")
    gt = (((yy - 56) ** 2 + (xx - 50) ** 2) < 30 ** 2).float()[None, None].repeat(2, 1, 1, 1).to(DEV)
    frozen = {k: v.clone() for k, v in pipe.backbone.state_dict().items()}
    losses = [float(tr.step(img, pts, gt)) for _ in range(8)]
    assert losses[-1] < losses[0], losses
    assert all(torch.equal(v, pipe.backbone.state_dict()[k]) for k, v in frozen.items())
    assert tr.arena.flat.data_ptr() == tr.params[0].grad.data_ptr()  # gradients live in the single all-reduce buffer
