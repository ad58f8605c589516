"""Checkpoint compatibility (SURVEY section 8b 'names that leak', 8f row f4): the reference stores only the trainable
parts of an IS model -- `head.*` and `embed_coords.*` (save_cfg upsampler=False, backbone=False,
core/model/iseg_probe_model.py:227-258) -- and restores them with `state_dict().update(ckpt)` + strict=False
(core/inference/utils.py:71-74).  The pipeline must expose exactly those key names and survive that round trip.
Host logic only: modules are constructed on the CPU, nothing is launched."""
import torch

import isegprobe_b200 as isp

REF_KEYS = {"head.convs.0.conv.weight", "head.convs.0.conv.bias", "head.convs.1.conv.weight", "head.convs.1.conv.bias",
            "head.classifier.weight", "head.classifier.bias", "embed_coords.proj.weight", "embed_coords.proj.bias"}


def _trainable_part(model):
    return {k: v.clone() for k, v in model.state_dict().items() if k.startswith(("head.", "embed_coords."))}


def test_is_checkpoint_keys_and_round_trip():
    torch.manual_seed(0)
    a = isp.ISegPipeline("loftup", {"upsampler_path": None, "n_dim": 384})
    ckpt = _trainable_part(a)
    assert set(ckpt) == REF_KEYS
    assert tuple(ckpt["head.convs.0.conv.weight"].shape) == (384, 384, 3, 3)
    assert tuple(ckpt["head.classifier.weight"].shape) == (1, 384, 1, 1)
    assert tuple(ckpt["embed_coords.proj.weight"].shape) == (384, 3, 14, 14)
    torch.manual_seed(1)
    b = isp.ISegPipeline("loftup", {"upsampler_path": None, "n_dim": 384})
    assert not torch.equal(b.head.classifier.weight, a.head.classifier.weight)
    cur = b.state_dict()
    cur.update(ckpt)                      # the reference's load path (inference/utils.py:71-74)
    b.load_state_dict(cur, strict=False)
    for k, v in ckpt.items():
        assert torch.equal(b.state_dict()[k], v), k
    # frozen parts keep their own (different) initialisation and are not part of the checkpoint
    assert not any(k.startswith(("backbone.", "upsampler.")) for k in ckpt)


def test_frozen_and_trainable_parameter_split():
    """ModelBuilder(freeze=True) semantics (core/utils/model_builder.py:59-76): backbone and upsampler parameters do
    not require grad, head and click embedding do; the optimizer still sees every named parameter."""
    m = isp.ISegPipeline("jbu_featup", {"backbone_type": "dinov2", "use_norm": True})
    req = {k.split(".")[0] for k, p in m.named_parameters() if p.requires_grad}
    assert req == {"head", "embed_coords"}
    assert all(not p.requires_grad for p in m.backbone.parameters())
    assert all(not p.requires_grad for p in m.upsampler.parameters())


def test_maskclip_pipeline_keys():
    m = isp.ISegPipeline("identity", {}, backbone="maskclip",
                         head_params={"in_channels": 512, "num_layers": 2, "num_classes": 1})
    ck = _trainable_part(m)
    assert set(ck) == REF_KEYS
    assert tuple(ck["embed_coords.proj.weight"].shape) == (768, 3, 16, 16)
    assert tuple(ck["head.convs.0.conv.weight"].shape) == (512, 512, 3, 3)
