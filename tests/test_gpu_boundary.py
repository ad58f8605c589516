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
import pytest
import torch

from oracle import ref_shim, synth
from tests.gpu_util import DEV, cosine

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.available(), reason="no reference tree staged under baseline/_ref")]

H = W = 448


@pytest.fixture(scope="module")
def models(tmp_path_factory):
    import isegprobe_b200 as isp
    from oracle import ref_model
    ckpt = ref_model.write_checkpoint("loftup", str(tmp_path_factory.mktemp("ckpt")))
    ref = ref_model.build("loftup", ours=False, ckpt=ckpt)  # built BEFORE the install: the reference's own modules
    assert type(ref.upsampler).__module__.startswith("core.model.upsamplers")
    assert type(ref.dist_maps).__module__ == "core.model.ops"
    ours = ref_model.build("loftup", ours=True, ckpt=ckpt)
    ref.to(DEV).eval()
    ours.to(DEV).eval()
    pipe = isp.ISegPipeline("loftup", {"upsampler_path": ckpt, "n_dim": 384})
    ref_model.load_synthetic_weights(pipe).to(DEV).eval()
    return ref, ours, pipe


def _inputs(B=2):
    image = torch.cat([synth.image_batch(B, H, W, seed=1), (synth.image_batch(B, H, W, seed=8)[:, :1] > 0.5).float()], 1)
    pts = synth.click_points(B, 3, H, W, seed=3)
    return image.to(DEV), pts.to(DEV)


def test_reference_model_with_our_modules_matches_the_reference(models):
    import isegprobe_b200 as isp
    ref_model, our_model, pipe = models
    # the same reference class (module reloads between the two builds make the objects distinct, not the code)
    assert type(our_model).__qualname__ == type(ref_model).__qualname__ == "iSegProbeModel"
    assert type(our_model).__module__ == type(ref_model).__module__ == "core.model.iseg_probe_model"
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
    # RAW agreement over all 2 x 448^2 pixels, nothing filtered.  With RANDOM-INIT weights the logits are centred on zero
    # (no trained bimodal mask), so the bf16 path's ~1.7 % relative logit error flips the sign of the ~0.5 % of pixels whose
    # reference logit lies inside that error band: the north-star 99.9 % is not reachable un-masked on this input in bf16
    # (measured 0.9949).  Asserted: raw >= 0.99, and EVERY disagreeing pixel lies within 3 % of the logit
    # range of the decision boundary (so pixels farther than that agree 100 %).
    wrong = (got > 0) != (want > 0)
    agree = 1.0 - float(wrong.float().mean())
    band = float(want.abs()[wrong].max() / want.abs().max()) if wrong.any() else 0.0
    print(f"reference-model drop-in: cosine {cosine(got, want):.6f}, RAW mask agreement {agree:.6f} over {want.numel()} pixels, "
          f"all disagreeing pixels have |reference logit| <= {band:.4f} of the maximum")
    assert agree >= 0.99, agree
    assert band <= 0.03, band
    # the fused assembly (ISegPipeline: one prepare_input kernel, bf16 hand-over to the head) is the same computation
    assert cosine(fused, got) > 0.9999
    assert float(((fused > 0) == (got > 0)).float().mean()) >= 0.999


def test_checkpoint_written_by_the_reference_reproduces_its_masks(models, tmp_path):
    """Row f4: the reference's own model, with "trained" (perturbed) head / click-embedding weights, writes a checkpoint the
    way its trainer does (core/utils/misc.py:36-68: {'state_dict': net.get_state_dict_to_save(), 'config': ...}, with the
    save_cfg filter of iseg_probe_model.py:199-258); a FRESH ISegPipeline with differently initialised trainable parts loads
    it through isegprobe_b200.checkpoint (the reference's restore, inference/utils.py:71-74) and reproduces the reference
    model's logits / masks."""
    import isegprobe_b200 as isp
    from isegprobe_b200 import checkpoint
    from oracle import ref_model as rm
    ref, _, _ = models
    trainable = [p for p in ref.parameters() if p.requires_grad]
    keep = [p.detach().clone() for p in trainable]
    g = torch.Generator().manual_seed(11)
    try:
        with torch.no_grad():
            for p in trainable:  # a stand-in for training: every trainable tensor moves by ~20 % of its scale
                p.add_((torch.randn(p.shape, generator=g) * 0.2 * max(float(p.abs().mean()), 1e-3)).to(p.device))
        sd = ref.get_state_dict_to_save()
        assert sd and all(k.startswith(checkpoint.TRAINABLE_PREFIXES) for k in sd), sorted(sd)[:5]  # frozen parts are not stored
        path = str(tmp_path / "last_checkpoint.pth")
        # misc.py:66 also stores net._config; it holds the ModelBuilder CLASS, which cannot be pickled here because the shim
        # reloads core.utils.model_builder between builds -- the restore path does not read it (inference/utils.py:71-74)
        torch.save({"state_dict": {k: v.detach().cpu() for k, v in sd.items()}, "config": {}}, path)
        image, pts = _inputs()
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            with torch.no_grad():
                want = ref(image, pts)["instances"].float().cpu()
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
    finally:
        with torch.no_grad():
            for p, k in zip(trainable, keep):
                p.copy_(k)
    assert checkpoint.main(["verify", path, "--upsampler", "loftup", "--n-dim", "384"]) == 0
    torch.manual_seed(5)
    pipe = isp.ISegPipeline("loftup", {"upsampler_path": rm.write_checkpoint("loftup", str(tmp_path)), "n_dim": 384})
    pipe.backbone.model.load_state_dict(synth.vit_state_dict(384, depth=12, seed=0))  # frozen parts: their own files
    pipe.to(DEV).eval()
    with torch.no_grad():
        before = pipe(image, pts)["instances"].float().cpu()
    assert cosine(before, want) < 0.9  # a random head does not reproduce the masks
    assert checkpoint.load_into(pipe, path) == []
    with torch.no_grad():
        got = pipe(image, pts)["instances"].float().cpu()
    wrong = (got > 0) != (want > 0)
    agree = 1.0 - float(wrong.float().mean())
    band = float(want.abs()[wrong].max() / want.abs().max()) if wrong.any() else 0.0
    print(f"reference-written checkpoint -> ISegPipeline: cosine {cosine(got, want):.6f}, RAW mask agreement {agree:.6f}, "
          f"disagreeing pixels within {band:.4f} of the logit range of the boundary")
    assert cosine(got, want) > 0.999
    assert agree >= 0.99 and band <= 0.03, (agree, band)


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
        assert c > 0.999, (k, c)  # measured min 0.99996


def test_reference_model_in_train_mode_matches(models):
    """`net.train()` as the reference's trainer issues it (trainer.py:213-214) on both models: the frozen LoftUp's
    BatchNorm switches to batch statistics (and moves its running statistics) in the reference AND in ours; logits,
    gradients of the trainable parameters and the updated running statistics must agree.  Runs last: it mutates the
    modules' BatchNorm buffers."""
    from isegprobe_b200.training import normalized_focal_loss
    ref_model, our_model, _ = models
    image, pts = _inputs(B=2)
    gt = (image[:, 3:] > 0.5).float()
    res = []
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for m in (ref_model, our_model):
            m.train()
            m.zero_grad(set_to_none=True)
            logits = m(image, pts)["instances"].float()
            normalized_focal_loss(logits, gt).mean().backward()
            grads = {k: p.grad.detach().float().cpu() for k, p in m.named_parameters() if p.requires_grad}
            bn = {k: v.detach().float().cpu() for k, v in m.upsampler.state_dict().items() if "running_" in k}
            res.append((logits.detach().cpu(), grads, bn))
            m.eval()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    (l0, g0, b0), (l1, g1, b1) = res
    assert cosine(l1, l0) > 0.999, cosine(l1, l0)
    for k in g0:
        assert cosine(g1[k], g0[k]) > 0.999, (k, cosine(g1[k], g0[k]))
    # reference keys: upsampler.upsampler.first_conv.N.running_*; ours: upsampler.upsampler.first_conv.N.running_* as well
    assert len(b0) == 4 and set(b0) == set(b1), (sorted(b0), sorted(b1))
    for k in b0:
        err = float((b1[k] - b0[k]).abs().max() / b0[k].abs().max())
        assert err < 2e-2, (k, err)
