"""GPU parity: LoftUp upsampler (bf16 tensor-core pipeline) vs the oracle and the reference's
golden vector.  north_star tolerance for bf16 mode: cosine >= 0.999."""
import pytest
import torch

from oracle import loftup as oloft
from oracle import synth
from tests.gpu_util import DEV, cosine, relerr

pytestmark = pytest.mark.gpu


def _module(dim=384):
    from isegprobe_b200.loftup import LoftUpUpsampler
    m = LoftUpUpsampler(None, n_dim=dim)
    sd, cn = synth.loftup_state_dict(dim, seed=0), synth.channelnorm_state_dict(dim, seed=1)
    m.upsampler.upsampler.load_state_dict(sd, strict=True)
    m.upsampler.channelnorm.load_state_dict(cn, strict=True)
    return m.to(DEV).eval(), sd, cn


def test_loftup_golden_reference(golden):
    g = golden("loftup_28x42")
    m, sd, cn = _module()
    img = (synth.image_batch(2, 28, 42, seed=1) - 0.45) / 0.225
    lr = synth.lr_features(2, 384, 2, 3, seed=2)
    with torch.no_grad():
        out = m(source=lr.to(DEV), guidance=img.to(DEV))
    want = torch.from_numpy(g["out"])
    assert tuple(out.shape) == (2, 384, 28, 42)
    c = cosine(out, want)
    assert c >= 0.999, c
    assert relerr(out, want) < 0.15


@pytest.mark.parametrize("B,H,W,h,w", [(1, 56, 56, 4, 4), (3, 64, 96, 8, 12), (5, 32, 32, 3, 3)])
def test_loftup_vs_oracle(B, H, W, h, w):
    m, sd, cn = _module()
    m.chunk_images = 2
    img = (synth.image_batch(B, H, W, seed=1) - 0.45) / 0.225
    lr = synth.lr_features(B, 384, h, w, seed=2)
    with torch.no_grad():
        out = m(source=lr.to(DEV), guidance=img.to(DEV))
        want = oloft.loftup_forward(sd, lr, img, cn["norm.weight"], cn["norm.bias"])
    c = cosine(out, want)
    assert c >= 0.999, c
    per_pixel = torch.nn.functional.cosine_similarity(out.cpu().float(), want, dim=1)
    assert float(per_pixel.min()) > 0.99, float(per_pixel.min())


def test_loftup_dim512_vs_oracle():
    """LoftUp(512) as used on MaskCLIP features (BASELINE config 4): 4 heads x 133 -> the
    144-column attention variant."""
    m, sd, cn = _module(512)
    img = (synth.image_batch(2, 48, 64, seed=1) - 0.45) / 0.225
    lr = synth.lr_features(2, 512, 3, 4, seed=2)
    with torch.no_grad():
        out = m(source=lr.to(DEV), guidance=img.to(DEV))
        want = oloft.loftup_forward(sd, lr, img, cn["norm.weight"], cn["norm.bias"])
    assert tuple(out.shape) == (2, 512, 48, 64)
    c = cosine(out, want)
    assert c >= 0.999, c


def test_loftup_state_dict_matches_reference_layout():
    """Reference checkpoints must load: same keys/shapes as LoftUp(dim).state_dict()."""
    from isegprobe_b200.loftup import LoftUpUpsampler
    m = LoftUpUpsampler(None, 384)
    ours = {k: tuple(v.shape) for k, v in m.upsampler.upsampler.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in synth.loftup_state_dict(384).items()}
    assert ours == ref
    ckpt = {"upsampler." + k: v for k, v in synth.loftup_state_dict(384, seed=3).items()}
    ckpt.update({"model.1." + k: v for k, v in synth.channelnorm_state_dict(384, seed=4).items()})
    m.load_reference_checkpoint(ckpt)


def test_loftup_fused_layernorm_matches_unfused():
    """LayerNorm applied in the consuming GEMM's epilogue (default) vs the stand-alone LayerNorm kernel: same
    function, different bf16 rounding points; both must sit within the bf16 tolerance of the oracle."""
    m, sd, cn = _module()
    img = (synth.image_batch(2, 64, 96, seed=1) - 0.45) / 0.225
    lr = synth.lr_features(2, 384, 8, 12, seed=2)
    with torch.no_grad():
        want = oloft.loftup_forward(sd, lr, img, cn["norm.weight"], cn["norm.bias"])
        outs = {}
        for fuse in (True, False):
            m.fuse_layernorm = fuse
            outs[fuse] = m(source=lr.to(DEV), guidance=img.to(DEV)).cpu().float()
    assert cosine(outs[True], outs[False]) > 0.9995
    ct, cf = cosine(outs[True], want), cosine(outs[False], want)
    assert ct >= 0.999 and cf >= 0.999, (ct, cf)
    assert ct > cf - 2e-4, (ct, cf)  # fusing must not cost accuracy
