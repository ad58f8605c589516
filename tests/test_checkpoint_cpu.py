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


def test_simple_vit_click_encoder_checkpoint_keys():
    """models/sbd/dinov2/simple-vit_noup.py saves `embed_coords.*` of a SimpleViTFeaturizer (save_cfg embed_coords=True):
    our module must expose the reference's keys and shapes (oracle/synth.simple_vit_state_dict lists them as produced by
    the reference class, oracle/make_golden.golden_simple_vit loads it strict=True) and the late-injection pipeline must
    keep the backbone frozen and the encoder trainable."""
    from oracle import synth
    ref = synth.simple_vit_state_dict(depth=6, seed=0)
    pipe = isp.ISegPipeline("identity", {}, embed_coords_type="simple_vit", feats_injection_mode="after_backbone")
    ours = {k[len("embed_coords."):]: v for k, v in pipe.state_dict().items() if k.startswith("embed_coords.")}
    assert set(ours) == set(ref)
    for k, v in ref.items():
        assert tuple(ours[k].shape) == tuple(v.shape), k
    pipe.embed_coords.load_state_dict(ref, strict=True)
    assert pipe.backbone.feats_injection_mode == "after_backbone"
    assert all(p.requires_grad for p in pipe.embed_coords.parameters())
    assert not any(p.requires_grad for p in pipe.backbone.parameters())
    assert pipe.embed_coords.reshape_feats_to_patches(torch.zeros(2, 1024, 384)).shape == (2, 384, 32, 32)
    try:
        isp.ISegPipeline("identity", {}, embed_coords_type="conv_stem")
    except ValueError as e:
        assert "Unsupported backbone type" in str(e)  # core/utils/model_builder.py:50-51
    else:
        raise AssertionError("unknown embed_coords type must raise")


def test_checkpoint_cli_verify_and_load(tmp_path):
    """isegprobe_b200.checkpoint: a reference-style IS checkpoint ({'state_dict': head.* + embed_coords.*}, DataParallel
    'module.' prefix tolerated) verifies against the matching configuration, loads the way the reference restores it
    (inference/utils.py:71-74), and a checkpoint of another head width is rejected."""
    from isegprobe_b200 import checkpoint
    torch.manual_seed(0)
    a = isp.ISegPipeline("loftup", {"upsampler_path": None, "n_dim": 384})
    path = str(tmp_path / "is.pth")
    torch.save({"state_dict": {"module." + k: v for k, v in _trainable_part(a).items()}, "config": {}}, path)
    assert checkpoint.main(["verify", path, "--upsampler", "loftup", "--n-dim", "384"]) == 0
    assert checkpoint.main(["inspect", path]) == 0
    torch.manual_seed(1)
    b = isp.ISegPipeline("loftup", {"upsampler_path": None, "n_dim": 384})
    assert checkpoint.load_into(b, path) == []
    for k, v in _trainable_part(a).items():
        assert torch.equal(b.state_dict()[k], v), k
    bad = {k: (v[:, :100] if k == "head.convs.0.conv.weight" else v) for k, v in _trainable_part(a).items()}
    import pytest
    with pytest.raises(ValueError):
        checkpoint.load_into(b, bad)
