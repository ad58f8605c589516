"""Pin the oracle (CPU restatement) against golden vectors produced by the
UNMODIFIED reference modules (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import distmaps as odm
from oracle import head as ohead
from oracle import jbu as ojbu
from oracle import lift as olift
from oracle import loftup as oloft
from oracle import maskclip as omc, synth, vit as ovit


def _relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


def test_distmaps_bit_exact(golden):
    g = golden("distmaps")
    for tag in ("int_p4", "f32_p4", "frac_p5", "frac_p24"):
        pts = g[f"{tag}_points"]
        for mode, disks in (("disk", True), ("tanh", False)):
            ref = g[f"{tag}_{mode}"]
            out = odm.distmaps(pts, 40, 56, norm_radius=5, use_disks=disks)
            if disks:
                assert np.array_equal(out, ref), (tag, mode)
            else:  # tanh/sqrt: libm vs ATen vectorised math, allow 2 ulp
                np.testing.assert_allclose(out, ref, rtol=0, atol=3e-7)
                d2 = odm.squared_distance_maps(pts, 40, 56, 5, 1.0, False)
                ref_d2 = torch.atanh(torch.from_numpy(ref).double().clamp(max=1 - 1e-12)) / 2
                near = ref < 0.99
                np.testing.assert_allclose(np.sqrt(d2)[near], ref_d2.numpy()[near], rtol=1e-4, atol=1e-5)


def test_distmaps_bfs_matches_cython(golden):
    g = golden("distmaps")
    if "bfs_points" not in g.files:
        import pytest
        pytest.skip("cython golden not generated")
    pts = g["bfs_points"]
    for b in range(pts.shape[0]):
        d2 = odm.bfs_squared_distance_maps(pts[b], 24, 32, 1.0)
        assert np.array_equal((d2 <= 25.0).astype(np.float32), g["bfs_disk"][b])
        # integer clicks: BFS == direct minimum
        direct = odm.squared_distance_maps(pts[b:b + 1], 24, 32, 5, 1.0, True)[0]
        assert np.array_equal(d2, direct)


def test_loftup_matches_reference(golden):
    g = golden("loftup_28x42")
    sd = synth.loftup_state_dict(384, seed=0)
    cn = synth.channelnorm_state_dict(384, seed=1)
    img = (synth.image_batch(2, 28, 42, seed=1) - 0.45) / 0.225
    lr = synth.lr_features(2, 384, 2, 3, seed=2)
    with torch.no_grad():
        ff = oloft.fourier_features(oloft.minmax_scale(img), sd["fourier_feat.1.biases"], 20, True)
        assert _relerr(ff[:, :, ::3, ::3], g["fourier"]) < 1e-5
        fc = oloft.first_conv(ff, sd)
        assert _relerr(fc[:, :, ::3, ::3], g["first_conv"]) < 1e-4
        out = oloft.loftup_forward(sd, lr, img, cn["norm.weight"], cn["norm.bias"])
    assert out.shape == (2, 384, 28, 42)
    assert _relerr(out, g["out"]) < 1e-4


def test_loftup_train_mode_matches_reference(golden):
    """train() semantics of the frozen upsampler (trainer.py:213-214): batch-statistics BatchNorm + running-stat updates."""
    g = golden("loftup_train_28x42")
    sd = synth.loftup_state_dict(384, seed=0)
    cn = synth.channelnorm_state_dict(384, seed=1)
    img = (synth.image_batch(3, 28, 42, seed=5) - 0.45) / 0.225
    lr = synth.lr_features(3, 384, 2, 3, seed=6)
    new = {}
    with torch.no_grad():
        out = oloft.loftup_forward(sd, lr, img, cn["norm.weight"], cn["norm.bias"], train_stats=new)
    assert _relerr(out, g["out"]) < 1e-4
    assert len(new) == 6
    for k, v in new.items():
        assert _relerr(v.float(), g[k.replace(".", "_")]) < 1e-5, k


def test_lift_matches_reference(golden):
    g = golden("lift_56x84")
    sd = synth.lift_state_dict(384, seed=0)
    img = (synth.image_batch(2, 56, 84, seed=1) - 0.45) / 0.225
    lr = synth.lr_features(2, 384, 4, 6, seed=2)
    with torch.no_grad():
        out = olift.lift_forward(sd, lr, img)
    assert _relerr(out, g["out"]) < 1e-5


def test_lift_train_mode_matches_reference(golden):
    """oracle.lift.lift_forward(train=True) == the reference LiFT module in train() (core/training/trainer.py:213-214 puts the
    frozen upsampler there): output, every running statistic / batch counter, and the gradient w.r.t. the LR features."""
    g = golden("lift_train_56x84")
    sd = {k: v.clone() for k, v in synth.lift_state_dict(384, seed=0).items()}
    img = (synth.image_batch(3, 56, 84, seed=5) - 0.45) / 0.225
    lr = synth.lr_features(3, 384, 4, 6, seed=6).requires_grad_(True)
    out = olift.lift_forward(sd, lr, img, train=True)
    (out * synth.lr_features(3, 384, 8, 12, seed=7)).sum().backward()
    assert _relerr(out.detach().numpy(), g["out"]) < 1e-4
    assert _relerr(lr.grad.numpy(), g["dsource"]) < 1e-4
    for k in g.files:
        if "running" in k:
            key = [n for n in sd if n.replace(".", "_") == k][0]
            assert _relerr(sd[key].numpy(), g[k]) < 1e-5, k
        elif "num_batches" in k:
            assert int(g[k]) == 1


def test_head_and_patch_embed_match_reference(golden):
    g = golden("head_20x28")
    sd = synth.convhead_state_dict(384, 2, 1, seed=0)
    x = synth.lr_features(2, 384, 20, 28, seed=4)
    with torch.no_grad():
        out = ohead.convhead_forward(sd, x)
        emb = ohead.patch_embed_forward(synth.patch_embed_state_dict(384, 14, 3, seed=0),
                                        synth.image_batch(2, 28, 42, seed=6))
    assert _relerr(out, g["out"]) < 1e-5
    assert _relerr(emb, g["patch_embed"]) < 1e-5


def test_vit_matches_reference(golden):
    g = golden("vit_56x84")
    sd = synth.vit_state_dict(384, depth=12, seed=0)
    img = (synth.image_batch(2, 56, 84, seed=1) - 0.45) / 0.225
    emb = synth.lr_features(2, 24, 384, 1, seed=7).squeeze(-1) * 0.1
    with torch.no_grad():
        out = ovit.dinov2_forward(sd, img, emb)
    assert _relerr(out, g["out"]) < 1e-4


def test_jbu_shape_contract_and_adaptive_conv_identities(golden):
    """JBU parity is UNPINNED (FeatUp is not in the reference tree); check the
    one contract the reference states (JBUFeatUp.py:36-45: x16, same channels)
    and internal consistency of the restatement."""
    g = golden("jbu_contract")
    sd = ojbu.init_state_dict(8, seed=0)
    src = synth.lr_features(1, 8, 3, 4, seed=2)
    gd = synth.image_batch(1, 48, 64, seed=1)
    with torch.no_grad():
        out = ojbu.jbu_stack_forward(sd, src, gd)
    assert tuple(out.shape) == (1, 8, 48, 64)
    assert list(g["out_shape"][2:] // g["source_shape"][2:]) == [16, 16]
    # adaptive conv with a delta filter at the centre is the identity on the unpadded input
    x = torch.randn(1, 3, 5 + 6, 6 + 6)
    f = torch.zeros(1, 5, 6, 7, 7)
    f[..., 3, 3] = 1
    assert torch.equal(ojbu.adaptive_conv(x, f), x[:, :, 3:-3, 3:-3])
    # grad_input is the adjoint of the forward: <A x, y> == <x, A^T y>
    f = torch.randn(1, 5, 6, 7, 7)
    y = torch.randn(1, 3, 5, 6)
    lhs = (ojbu.adaptive_conv(x, f) * y).sum()
    rhs = (x * ojbu.adaptive_conv_grad_input(y, f)).sum()
    assert abs(lhs - rhs) < 1e-3 * abs(lhs).clamp(min=1)
    # filters sum to ~1 before the fixup term => a constant image stays ~constant
    k = ojbu.range_kernel(sd, "up1", gd[:, :, :6, :8]) * ojbu.spatial_kernel(sd, "up1")
    k = k / k.sum(1, keepdim=True).clamp(1e-7)
    assert torch.allclose(k.sum(1), torch.ones(1, 6, 8), atol=1e-5)


def test_maskclip_matches_reference(golden):
    """Oracle restatement vs the unmodified maskclip VisionTransformer (patch_output path, with and
    without click-embedding injection; non-square input exercises the (w, h) ordering quirk)."""
    g = golden("maskclip_64x96")
    sd = synth.maskclip_state_dict(seed=0)
    img = (synth.image_batch(2, 64, 96, seed=1) - 0.45) / 0.225
    emb = synth.lr_features(2, 24, 768, 1, seed=7).squeeze(-1) * 0.1
    with torch.no_grad():
        plain = omc.maskclip_forward(sd, img)
        inj = omc.maskclip_forward(sd, img, emb)
        sq = omc.maskclip_forward(sd, (synth.image_batch(1, 64, 64, seed=2) - 0.45) / 0.225)
    for out, key in ((plain, "plain"), (inj, "injected"), (sq, "square")):
        want = torch.from_numpy(g[key])
        assert tuple(out.shape) == tuple(want.shape)
        assert float((out - want).abs().max() / want.abs().max()) < 1e-4, key


def test_next_points_matches_reference(golden):
    """Click simulation of the training loop (isegprobe_b200/training.get_next_points) vs the reference function's own
    source run on the same inputs and numpy RNG state."""
    import numpy as np
    pytest.importorskip("cv2")
    from isegprobe_b200.training import get_next_points
    g = golden("next_points")
    pred, gt = torch.from_numpy(g["pred"]), torch.from_numpy(g["gt"])
    points = torch.full((3, 12, 3), -1.0)
    np.random.seed(123)
    out1 = get_next_points(pred, gt, points, 1)
    out2 = get_next_points(pred.flip(3), gt, out1, 2)
    assert np.array_equal(out1.numpy(), g["out1"]) and np.array_equal(out2.numpy(), g["out2"])
    assert (out1[:, :, 2] == 1).sum() == 3 and (out2[:, :, 2] == 2).sum() == 3


def test_nfl_loss_matches_reference(golden):
    """The training step's loss restatement (isegprobe_b200/training.py) vs the reference's
    NormalizedFocalLossSigmoid(alpha=0.5, gamma=2): value and gradient, ignore-label rows included."""
    from isegprobe_b200.training import normalized_focal_loss
    g = golden("nfl_loss")
    pred = torch.from_numpy(g["pred"]).requires_grad_(True)
    out = normalized_focal_loss(pred, torch.from_numpy(g["label"]))
    out.mean().backward()
    assert float((out.detach() - torch.from_numpy(g["out"])).abs().max()) < 1e-6
    assert float((pred.grad - torch.from_numpy(g["grad"])).abs().max()) < 1e-7


def test_dino_vit_oracle_matches_reference_golden(golden):
    """DINO / timm ViT-S/16 adapter (DINO.py:529-611), 'key' and 'token' features, non-square image, click embedding
    injected before the blocks."""
    from oracle import dino as odino
    g = golden("dino_vit_64x96")
    sd = synth.dino_vit_state_dict(seed=0)
    img = (synth.image_batch(2, 64, 96, seed=1) - 0.45) / 0.225
    emb = synth.lr_features(2, 24, 384, 1, seed=7).squeeze(-1) * 0.1
    with torch.no_grad():
        for ft in ("key", "token"):
            out = odino.dino_vit_forward(sd, img, emb, feat_type=ft)
            want = torch.from_numpy(g[ft])
            assert float((out - want).abs().max() / want.abs().max()) < 1e-5, ft


def test_simple_vit_oracle_matches_reference_golden(golden):
    """SimpleViTFeaturizer (simple_ViT.py:96-146), the trainable click embedding of dinov2/simple-vit_noup.py."""
    from oracle import simple_vit as osv
    g = golden("simple_vit_56x84")
    sd = synth.simple_vit_state_dict(depth=2, seed=0)
    x = synth.image_batch(2, 56, 84, seed=4)
    with torch.no_grad():
        out = osv.simple_vit_forward(sd, x)
    want = torch.from_numpy(g["out"])
    assert float((out - want).abs().max() / want.abs().max()) < 1e-5
