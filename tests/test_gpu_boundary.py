"""Drop-in proof on the GPU: the reference's OWN `iSegProbeModel.forward(image, points)`
(core/model/iseg_base_model.py:67-89, iseg_probe_model.py:110-134) is run twice on cuda:0 at 448x448 --

  (A) entirely with the reference's modules (fp32 eager torch: DistMaps, vendored DINOv2 ViT-S/14, LoftUp loaded from a
      checkpoint file in the upstream format, mmcv-style ConvSegHead), and
  (B) the same class, same config dictionaries, same checkpoint file, after `install_into_reference(featurizers=True)`
      -- every hot-path module is then this package's CUDA implementation --

and (B)'s logits / masks are compared with (A)'s and with `ISegPipeline` (the fused assembly bench.py times).
The only stand-ins are the two network downloads the reference would do: torch.hub's `dinov2_vits14` (replaced by the
reference's own vendored `vit_small`, random init) and the LoftUp checkpoint (written here in upstream's format).

Reads the staged copy under baseline/_ref (made by __graft_entry__.build(); git-ignored); skipped if absent."""
import os

import pytest
import torch

from oracle import ref_shim, synth
from tests.gpu_util import DEV, cosine
from tests.test_boundary_cpu import reference_cfgs

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.available(), reason="no reference tree staged under baseline/_ref")]

H = W = 448


def _loftup_checkpoint(path):
    """Upstream LoftUp checkpoint layout read by load_loftup_checkpoint (loftup/loftup.py:152-177)."""
    usd, cn = synth.loftup_state_dict(384, seed=0), synth.channelnorm_state_dict(384, seed=1)
    sd = {"upsampler." + k: v for k, v in usd.items()}
    sd.update({"model.1." + k: v for k, v in cn.items()})
    torch.save({"state_dict": sd}, path)


@pytest.fixture(scope="module")
def models(tmp_path_factory):
    ref_shim.install()
    ckpt = str(tmp_path_factory.mktemp("ckpt") / "loftup_synth.ckpt")
    _loftup_checkpoint(ckpt)
    vsd = synth.vit_state_dict(384, depth=12, seed=0)
    hsd = synth.convhead_state_dict(384, 2, 1, seed=0)
    psd = synth.patch_embed_state_dict(384, 14, 3, seed=0)

    import importlib

    import core.model.featurizers.DINOv2 as ref_dino
    import core.model.heads as ref_heads
    import core.model.iseg_base_model as ref_ibm
    import core.model.iseg_probe_model as ref_ipm
    import core.model.ops as ref_ops
    import core.model.upsamplers as ref_up
    import core.utils.model_builder as ref_mb
    # pristine reference state, whatever earlier tests installed into the shimmed modules
    for m in (ref_ops, ref_up, ref_heads, ref_ibm, ref_mb, ref_ipm):
        importlib.reload(m)

    def hub_load(repo, arch, *a, **k):  # the one network call on the path (DINOv2.py:491)
        assert arch == "dinov2_vits14"
        return ref_dino.vit_small(patch_size=14, img_size=518, init_values=1.0, block_chunks=0)

    real_hub = torch.hub.load
    torch.hub.load = hub_load
    try:
        # `from core.model.featurizers import *` is empty under the namespace shim (the package __init__ needs timm)
        ref_mb.DINOv2Featurizer = ref_dino.DINOv2Featurizer
        cfg = reference_cfgs("loftup")
        cfg["upsampler_cfg"]["params"]["upsampler_path"] = ckpt
        kw = dict(model_builder=ref_mb.ModelBuilder(), use_disks=True, norm_radius=5, with_prev_mask=True)
        ref_model = ref_ipm.iSegProbeModel(**cfg, **kw)
        assert type(ref_model.upsampler).__module__.startswith("core.model.upsamplers")
        assert type(ref_model.dist_maps).__module__ == "core.model.ops"

        import isegprobe_b200 as isp
        isp.install_into_reference(featurizers=True)
        our_model = ref_ipm.iSegProbeModel(**cfg, **kw)
    finally:
        torch.hub.load = real_hub
    for m in (ref_model, our_model):
        m.backbone.model.load_state_dict(vsd)
        m.head.load_state_dict(hsd)
        m.embed_coords.load_state_dict(psd)
        m.to(DEV).eval()
    pipe = isp.ISegPipeline("loftup", {"upsampler_path": ckpt, "n_dim": 384}).to(DEV).eval()
    pipe.backbone.model.load_state_dict(vsd)
    pipe.head.load_state_dict(hsd)
    pipe.embed_coords.load_state_dict(psd)
    return ref_model, our_model, pipe


def _inputs(B=2):
    image = torch.cat([synth.image_batch(B, H, W, seed=1), (synth.image_batch(B, H, W, seed=8)[:, :1] > 0.5).float()], 1)
    pts = synth.click_points(B, 3, H, W, seed=3)
    return image.to(DEV), pts.to(DEV)


def test_reference_model_with_our_modules_matches_the_reference(models):
    import isegprobe_b200 as isp
    ref_model, our_model, pipe = models
    assert type(our_model) is type(ref_model)
    assert type(our_model.dist_maps) is isp.DistMaps and type(our_model.upsampler) is isp.LoftUpUpsampler
    assert type(our_model.head) is isp.ConvSegHead
    image, pts = _inputs()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # the reference arm is fp32 (SURVEY: no TF32 / autocast anywhere)
    try:
        with torch.no_grad():
            want = ref_model(image, pts)["instances"].float().cpu()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    l0 = isp._lib.launch_count()
    with torch.no_grad():
        got = our_model(image, pts)["instances"].float().cpu()
        fused = pipe(image, pts)["instances"].float().cpu()
    assert isp._lib.launch_count() - l0 > 200  # the CUDA library did the work
    assert tuple(got.shape) == tuple(want.shape) == (2, 1, H, W)
    # click maps through the reference's own call site are bit-exact
    with torch.no_grad():
        a = ref_model.dist_maps(image[:, :3], pts)
        b = our_model.dist_maps(image[:, :3], pts)
    assert torch.equal(a, b)
    # bf16 tensor-core path vs fp32 reference
    assert cosine(got, want) > 0.999, cosine(got, want)
    agree = float(((got > 0) == (want > 0)).float().mean())  # RAW agreement over all 2 x 448^2 pixels
    print(f"reference-model drop-in: cosine {cosine(got, want):.6f}, raw mask agreement {agree:.6f}")
    assert agree >= 0.999, agree
    # the fused assembly (ISegPipeline: one prepare_input kernel, bf16 hand-over to the head) is the same computation
    assert cosine(fused, got) > 0.9999
    assert float(((fused > 0) == (got > 0)).float().mean()) >= 0.999


def test_reference_model_trains_through_our_modules(models):
    """trainer.py:451-459 with the reference's own model object: loss.backward() reaches the trainable parameters
    (head + click embedding through the frozen upsampler and backbone) and matches the reference's gradients."""
    from isegprobe_b200.training import normalized_focal_loss
    ref_model, our_model, _ = models
    image, pts = _inputs(B=1)
    gt = (image[:, 3:] > 0.5).float()
    grads = []
    for m in (ref_model, our_model):
        m.zero_grad(set_to_none=True)
        loss = normalized_focal_loss(m(image, pts)["instances"].float(), gt).mean()
        loss.backward()
        grads.append({k: p.grad.detach().float().cpu() for k, p in m.named_parameters() if p.requires_grad})
    assert set(grads[0]) == set(grads[1]) and "embed_coords.proj.weight" in grads[0]
    for k in grads[0]:
        c = cosine(grads[1][k], grads[0][k])
        assert c > 0.99, (k, c)
