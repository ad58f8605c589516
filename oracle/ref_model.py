"""ORACLE (test infrastructure) -- the reference's OWN assembled model, built from its own config dictionaries.

`build(...)` constructs `core.model.iseg_probe_model.iSegProbeModel` (core/model/iseg_probe_model.py:34-108) through
`core.utils.model_builder.ModelBuilder` (core/utils/model_builder.py:13-95) exactly like
`models/sbd/dinov2/patch-embed_loftup.py:91-112` (init_model), from the UNMODIFIED reference tree that
`oracle/ref_shim.py` points at (/root/reference here, the staged git-ignored baseline/_ref copy on the GPU box).

Only the reference's two network downloads are replaced:
  * torch.hub's `dinov2_vits14` (DINOv2.py:491) -> the reference's vendored `vit_small(patch_size=14, img_size=518,
    init_values=1.0, block_chunks=0)` (DINOv2.py:413-423), random init;
  * the LoftUp / LiFT checkpoint files -> files written here in the upstream formats (loftup/loftup.py:152-177,
    LiFT.py:125-136) from the seeded synthetic state dicts of oracle/synth.py.

Used by tests/test_gpu_boundary.py and by bench.py's `--impl reference` / reference-on-B200 legs.  Never by the product.
"""
import importlib
import os
import tempfile

import torch

from . import ref_shim, synth


def reference_cfgs(upsampler="loftup"):
    """The config dictionaries of models/sbd/dinov2/patch-embed_{loftup,lift,jbu}.py:27-88 (define_modules_cfg).
    Checkpoint paths are None here (= random init, this package's documented extension; the reference always
    torch.load()s); `build` fills in synthetic checkpoint files."""
    up = {"loftup": dict(type="loftup", params=dict(upsampler_path=None, n_dim=384)),
          "jbu_featup": dict(type="jbu_featup", params=dict(backbone_type="dinov2", use_norm=True)),
          "lift": dict(type="lift", params=dict(lift_path=None, n_dim=384, patch=14)),
          "bilinear": dict(type="bilinear", params=None)}[upsampler]
    return dict(
        backbone_cfg=dict(type="dinov2", params=dict(feats_injection_mode="before_backbone")),
        embed_coords_cfg=dict(type="patchEmbed", params=dict(img_size=(448, 448), patch_size=(14, 14), embed_dim=384)),
        head_cfg=dict(type="convhead", params=dict(in_channels=384, num_layers=2, num_classes=1)),
        upsampler_cfg=up, neck_cfg=None,
        save_cfg=dict(embed_coords=True, backbone=False, upsampler=False, head=True),
        architecture="backbone_upsampler_head")


def write_checkpoint(upsampler, directory=None):
    """Synthetic checkpoint file in the upstream layout the reference's loader reads; returns its path (None if the
    upsampler takes no file)."""
    directory = directory or tempfile.mkdtemp(prefix="isp_ckpt_")
    if upsampler == "loftup":
        usd, cn = synth.loftup_state_dict(384, seed=0), synth.channelnorm_state_dict(384, seed=1)
        sd = {"upsampler." + k: v for k, v in usd.items()}
        sd.update({"model.1." + k: v for k, v in cn.items()})
        path = os.path.join(directory, "loftup_synth.ckpt")
        torch.save({"state_dict": sd}, path)
        return path
    if upsampler == "lift":
        path = os.path.join(directory, "lift_synth.pth")
        torch.save({"module." + k: v for k, v in synth.lift_state_dict(384, seed=0).items()}, path)
        return path
    return None


def pristine():
    """Reload the shimmed reference modules so registries / class bindings are the reference's own again (undoes an
    earlier `install_into_reference()` in this process).  Returns the modules."""
    ref_shim.install()
    import core.model.heads as ref_heads
    import core.model.iseg_base_model as ref_ibm
    import core.model.iseg_probe_model as ref_ipm
    import core.model.ops as ref_ops
    import core.model.upsamplers as ref_up
    import core.utils.model_builder as ref_mb
    for m in (ref_ops, ref_up, ref_heads, ref_ibm, ref_mb, ref_ipm):
        importlib.reload(m)
    import core.model.featurizers.DINOv2 as ref_dino
    # `from core.model.featurizers import *` is empty under the namespace shim (the package __init__ needs timm)
    ref_mb.DINOv2Featurizer = ref_dino.DINOv2Featurizer
    return ref_ipm, ref_mb, ref_dino


def load_synthetic_weights(model):
    """Seeded weights for the parts that have no checkpoint file (same tensors the GPU tests load into ours)."""
    model.backbone.model.load_state_dict(synth.vit_state_dict(384, depth=12, seed=0))
    model.head.load_state_dict(synth.convhead_state_dict(384, 2, 1, seed=0))
    model.embed_coords.load_state_dict(synth.patch_embed_state_dict(384, 14, 3, seed=0))
    return model


def build(upsampler="loftup", ours=False, ckpt=None):
    """ours=False: the all-reference model.  ours=True: the same class / config after
    `isegprobe_b200.install_into_reference(featurizers=True)` (every hot-path module is then the CUDA implementation)."""
    assert ref_shim.available(), "no reference tree (run __graft_entry__.build() where /root/reference exists)"
    ref_ipm, ref_mb, ref_dino = pristine()
    cfg = reference_cfgs(upsampler)
    if ckpt is None:
        ckpt = write_checkpoint(upsampler)
    if upsampler == "loftup":
        cfg["upsampler_cfg"]["params"]["upsampler_path"] = ckpt
    elif upsampler == "lift":
        cfg["upsampler_cfg"]["params"]["lift_path"] = ckpt

    def hub_load(repo, arch, *a, **k):  # the one network call on the path (DINOv2.py:491)
        assert arch == "dinov2_vits14", arch
        return ref_dino.vit_small(patch_size=14, img_size=518, init_values=1.0, block_chunks=0)

    real_hub = torch.hub.load
    torch.hub.load = hub_load
    try:
        if ours:
            import isegprobe_b200 as isp
            isp.install_into_reference(featurizers=True)
        model = ref_ipm.iSegProbeModel(**cfg, model_builder=ref_mb.ModelBuilder(), use_disks=True, norm_radius=5,
                                       with_prev_mask=True)
    finally:
        torch.hub.load = real_hub
    return load_synthetic_weights(model)
