"""The drop-in boundary, exercised through the reference's OWN model assembly:
`install_into_reference()` followed by the unmodified `iSegProbeModel(... model_builder=ModelBuilder())`
(core/model/iseg_probe_model.py:34-108, core/utils/model_builder.py:13-95) must yield a model whose click-map encoder,
upsampler, head (and, with featurizers=True, backbone + click embedding) are this package's modules.

Needs a reference tree (/root/reference here, the staged baseline/_ref copy elsewhere); skipped without one.
Construction only -- no compute without a GPU (tests/test_gpu_boundary.py runs the forward)."""
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="no reference tree (run __graft_entry__.build() where /root/reference exists)")


def build_reference_model(upsampler="loftup"):
    """models/sbd/dinov2/patch-embed_loftup.py:91-112 (init_model) with the reference's own classes."""
    ref_shim.install()
    import isegprobe_b200 as isp
    isp.install_into_reference(featurizers=True)
    from core.model.iseg_probe_model import iSegProbeModel
    from core.utils.model_builder import ModelBuilder
    from oracle.ref_model import reference_cfgs
    return iSegProbeModel(**reference_cfgs(upsampler), model_builder=ModelBuilder(), use_disks=True, norm_radius=5,
                          with_prev_mask=True)


def test_distmaps_swap_reaches_the_module_that_constructs_it():
    """iseg_base_model.py:9 binds DistMaps at import time; the swap must rebind THAT name (round-1 bug)."""
    ref_shim.install()
    import core.model.iseg_base_model as ibm  # imported BEFORE the install, like `import core.model` does
    import core.model.ops as ref_ops
    import isegprobe_b200 as isp
    isp.install_into_reference()
    assert ibm.DistMaps is isp.DistMaps
    assert ref_ops.DistMaps is isp.DistMaps


@pytest.mark.parametrize("upsampler", ["loftup", "jbu_featup", "lift"])
def test_reference_model_is_built_from_our_modules(upsampler):
    import isegprobe_b200 as isp
    from isegprobe_b200 import featurizers
    model = build_reference_model(upsampler)
    assert type(model).__module__ == "core.model.iseg_probe_model"  # the reference's class, untouched
    assert type(model.dist_maps) is isp.DistMaps
    assert model.dist_maps.use_disks is True and model.dist_maps.norm_radius == 5 and model.dist_maps.cpu_mode is False
    assert type(model.upsampler) is isp.UPSAMPLER_REGISTRY[upsampler]
    assert type(model.head) is isp.ConvSegHead
    assert type(model.backbone) is featurizers.DINOv2Featurizer
    assert type(model.embed_coords) is featurizers.PatchEmbed
    # ModelBuilder(freeze=True) semantics survive: frozen backbone / upsampler, trainable head + click embedding
    assert not any(p.requires_grad for p in model.upsampler.parameters())
    assert not any(p.requires_grad for p in model.backbone.parameters())
    assert all(p.requires_grad for p in model.head.parameters())
    assert all(p.requires_grad for p in model.embed_coords.parameters())
    # the state-dict keys that leak into reference checkpoints (SURVEY 8b "names that leak")
    keys = set(model.state_dict())
    for k in ("head.convs.0.conv.weight", "head.convs.1.conv.bias", "head.classifier.weight",
              "embed_coords.proj.weight", "embed_coords.proj.bias"):
        assert k in keys, k
    # the reference's selective checkpoint filter works on our modules (iseg_probe_model.py:199-258)
    saved = model.get_state_dict_to_save()
    assert saved and all(k.startswith(("head.", "embed_coords.")) for k in saved)


def test_swap_modules_converts_a_prebuilt_reference_model():
    ref_shim.install()
    import importlib

    import isegprobe_b200 as isp
    from isegprobe_b200.pipeline import swap_modules
    import core.model.ops as ref_ops
    ref_cls = importlib.reload(ref_ops).DistMaps  # a pristine reference class, whatever earlier tests installed
    import torch.nn as nn

    class Stub(nn.Module):
        def __init__(self):
            super().__init__()
            self.dist_maps = ref_cls(norm_radius=5, spatial_scale=1.0, cpu_mode=False, use_disks=True)

    m = swap_modules(Stub())
    assert type(m.dist_maps) is isp.DistMaps and m.dist_maps.use_disks and m.dist_maps.norm_radius == 5
    isp.install_into_reference()  # leave the shimmed modules in the installed state
