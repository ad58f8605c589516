"""GPU parity: click-map encoding (bit-exact for disks) vs the oracle and the
reference's golden vectors."""
import numpy as np
import pytest
import torch

from oracle import distmaps as odm
from oracle import head as ohead
from oracle import synth
from tests.gpu_util import DEV

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def isp():
    import isegprobe_b200
    return isegprobe_b200


def test_golden_distmaps(isp, golden):
    g = golden("distmaps")
    img = torch.zeros(3, 3, 40, 56, device=DEV)
    for tag in ("int_p4", "f32_p4", "frac_p5", "frac_p24"):
        pts = torch.from_numpy(g[f"{tag}_points"]).to(DEV)
        out = isp.DistMaps(5, 1.0, False, True)(img, pts).cpu().numpy()
        assert np.array_equal(out, g[f"{tag}_disk"]), tag  # bit-exact
        out = isp.DistMaps(5, 1.0, False, False)(img, pts).cpu().numpy()
        np.testing.assert_allclose(out, g[f"{tag}_tanh"], rtol=0, atol=4e-7)  # tanhf vs ATen tanh: <= 2 ulp near 1


@pytest.mark.parametrize("B,P,H,W,frac", [(1, 1, 448, 448, False), (4, 24, 448, 448, True), (2, 20, 301, 517, True),
                                          (3, 5, 7, 9, False)])
def test_distmaps_vs_oracle(isp, B, P, H, W, frac):
    pts = synth.click_points(B, P, H, W, seed=B * 100 + P, frac=frac)
    if B > 1:
        pts[1] = -1  # an image with no clicks at all
    for disks in (True, False):
        out = isp.DistMaps(5, 1.0, False, disks)(torch.zeros(B, 3, H, W, device=DEV), pts.to(DEV)).cpu().numpy()
        ref = odm.distmaps(pts.numpy(), H, W, 5, 1.0, disks)
        if disks:
            assert np.array_equal(out, ref)
        else:
            np.testing.assert_allclose(out, ref, rtol=0, atol=4e-7)


def test_distmaps_empty_points(isp):
    out = isp.DistMaps(5, use_disks=True)(torch.zeros(2, 3, 16, 20, device=DEV), torch.zeros(2, 0, 3, device=DEV))
    assert out.shape == (2, 2, 16, 20) and float(out.abs().sum()) == 0.0


def test_distmaps_cython_semantics(isp, golden):
    g = golden("distmaps")
    if "bfs_points" not in g.files:
        pytest.skip("no cython golden")
    pts = torch.from_numpy(g["bfs_points"]).to(DEV)
    img = torch.zeros(2, 3, 24, 32, device=DEV)
    out = isp.DistMaps(5, 1.0, True, True)(img, pts).cpu().numpy()
    assert np.array_equal(out, g["bfs_disk"])
    out = isp.DistMaps(5, 1.0, True, False)(img, pts).cpu().numpy()
    np.testing.assert_allclose(out, g["bfs_tanh"], rtol=0, atol=4e-7)


def test_prepare_input_fused(isp):
    B, H, W, P = 3, 70, 90, 6
    img = torch.cat([synth.image_batch(B, H, W, seed=1), synth.image_batch(B, H, W, seed=9)[:, :1]], 1)
    pts = synth.click_points(B, P, H, W, seed=4, frac=True)
    norm, coord = isp.prepare_input(img.to(DEV), pts.to(DEV), 5, 1.0, True)
    ref_norm = ohead.normalize_image(img[:, :3])
    ref_maps = odm.distmaps(pts.numpy(), H, W, 5, 1.0, True)
    assert torch.allclose(norm.cpu(), ref_norm, rtol=0, atol=1e-6)
    assert torch.equal(coord[:, 0].cpu(), img[:, 3])
    assert np.array_equal(coord[:, 1:].cpu().numpy(), ref_maps)
    norm3, coord2 = isp.prepare_input(img[:, :3].to(DEV), pts.to(DEV), 5, 1.0, True)
    assert coord2.shape == (B, 2, H, W) and np.array_equal(coord2.cpu().numpy(), ref_maps)
