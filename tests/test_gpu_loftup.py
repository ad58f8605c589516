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


def test_loftup_train_mode_golden_reference(golden):
    """module.train() (what the reference's trainer does to the frozen upsampler, core/training/trainer.py:213-214):
    BatchNorm2d of first_conv with batch statistics over the whole batch -- across the internal image chunks -- and the
    running statistics updated like nn.BatchNorm2d; pinned to the reference module run in train() (golden vectors)."""
    g = golden("loftup_train_28x42")
    m, sd, cn = _module()
    m.chunk_images = 2  # 3 images -> two chunks: the statistics must still be those of the whole batch
    m.train()
    img = (synth.image_batch(3, 28, 42, seed=5) - 0.45) / 0.225
    lr = synth.lr_features(3, 384, 2, 3, seed=6)
    with torch.no_grad():
        out = m(source=lr.to(DEV), guidance=img.to(DEV))
    want = torch.from_numpy(g["out"])
    c = cosine(out, want)
    assert c >= 0.999, c
    new = m.upsampler.upsampler.state_dict()
    for k in ("first_conv.2.running_mean", "first_conv.2.running_var", "first_conv.5.running_mean", "first_conv.5.running_var"):
        assert relerr(new[k], torch.from_numpy(g[k.replace(".", "_")])) < 2e-2, (k, relerr(new[k], torch.from_numpy(g[k.replace(".", "_")])))
        assert relerr(new[k], sd[k]) > 1e-3  # they did move
    assert int(new["first_conv.2.num_batches_tracked"]) == 1 and int(new["first_conv.5.num_batches_tracked"]) == 1
    # eval() afterwards folds the UPDATED running statistics (oracle eval forward on the new state dict)
    m.eval()
    sd2 = {k: v.detach().cpu().clone() for k, v in new.items()}
    img2 = (synth.image_batch(2, 28, 42, seed=1) - 0.45) / 0.225
    lr2 = synth.lr_features(2, 384, 2, 3, seed=2)
    with torch.no_grad():
        out2 = m(source=lr2.to(DEV), guidance=img2.to(DEV))
        want2 = oloft.loftup_forward(sd2, lr2, img2, cn["norm.weight"], cn["norm.bias"])
    assert cosine(out2, want2) >= 0.999


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


def test_loftup_fused_ffn_matches_two_gemm_path():
    """fuse_ffn (isp_ffn_fused_bf16_tc: the FeedForward block in one kernel) computes the same function as the two GEMM
    launches; both within the bf16 tolerance of the oracle."""
    m, sd, cn = _module()
    img = (synth.image_batch(2, 64, 96, seed=1) - 0.45) / 0.225
    lr = synth.lr_features(2, 384, 8, 12, seed=2)
    with torch.no_grad():
        a = m(source=lr.to(DEV), guidance=img.to(DEV)).float()
        m.fuse_ffn = True
        b = m(source=lr.to(DEV), guidance=img.to(DEV)).float()
        want = oloft.loftup_forward(sd, lr, img, cn["norm.weight"], cn["norm.bias"])
    assert cosine(a, b) > 0.9999 and relerr(a, b) < 3e-2
    assert cosine(b, want) >= 0.999


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
