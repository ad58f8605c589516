"""Checkpoint tooling: the reference stores only the trainable parts of an IS model -- `head.*`
and `embed_coords.*` (core/model/iseg_probe_model.py:227-258, save_cfg upsampler=False, backbone=False) -- and restores them
with `state_dict().update(ckpt)` + `load_state_dict(strict=False)` (core/inference/utils.py:71-74); the frozen upsamplers
load their own files (loftup.py:152-177 key remap, LiFT.py:129-132 `module.` prefix, FeatUp state dict).  No offline packing
step exists or is needed: every module repacks its weights for the tensor-core kernels lazily, keyed on the parameter
versions.  This module is the load path plus a verifier:

    python -m isegprobe_b200.checkpoint verify  CKPT.pth --upsampler loftup --n-dim 384
    python -m isegprobe_b200.checkpoint inspect CKPT.pth

`verify` builds the pipeline on the CPU (no kernel is launched), applies the checkpoint the way the reference does and reports
missing / unexpected / mis-shaped keys; exit status 1 if the checkpoint does not fit."""
import argparse
import sys

import torch

TRAINABLE_PREFIXES = ("head.", "embed_coords.")


def read_state_dict(path_or_dict):
    """A reference IS checkpoint is either the bare state dict or {'state_dict': ..., 'config': ...} (inference/utils.py:60-70)."""
    ck = torch.load(path_or_dict, map_location="cpu", weights_only=False) if isinstance(path_or_dict, str) else path_or_dict
    if isinstance(ck, dict) and "state_dict" in ck and isinstance(ck["state_dict"], dict):
        ck = ck["state_dict"]
    return {(k[7:] if k.startswith("module.") else k): v for k, v in ck.items()}


def check(pipeline, ckpt):
    """Compare a checkpoint with a pipeline's state dict: (missing trainable keys, unexpected keys, shape mismatches)."""
    cur = pipeline.state_dict()
    missing = sorted(k for k in cur if k.startswith(TRAINABLE_PREFIXES) and k not in ckpt)
    unexpected = sorted(k for k in ckpt if k not in cur)
    shapes = sorted((k, tuple(ckpt[k].shape), tuple(cur[k].shape)) for k in ckpt
                    if k in cur and tuple(ckpt[k].shape) != tuple(cur[k].shape))
    return missing, unexpected, shapes


def load_into(pipeline, path_or_dict, strict_trainable=True):
    """The reference's restore (inference/utils.py:71-74): update the current state dict with the checkpoint's entries and
    load non-strictly; raises if a trainable key is missing or a shape differs (the reference would fail later, or silently
    keep a random head)."""
    ckpt = read_state_dict(path_or_dict)
    missing, unexpected, shapes = check(pipeline, ckpt)
    if shapes or (strict_trainable and missing):
        raise ValueError(f"checkpoint does not fit the pipeline: missing {missing[:4]}, shape mismatches {shapes[:4]}")
    cur = pipeline.state_dict()
    cur.update({k: v for k, v in ckpt.items() if k in cur})
    pipeline.load_state_dict(cur, strict=False)
    return unexpected


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m isegprobe_b200.checkpoint")
    ap.add_argument("cmd", choices=["inspect", "verify"])
    ap.add_argument("ckpt")
    ap.add_argument("--upsampler", default="loftup")
    ap.add_argument("--n-dim", type=int, default=384)
    a = ap.parse_args(argv)
    ckpt = read_state_dict(a.ckpt)
    if a.cmd == "inspect":
        for k, v in ckpt.items():
            print(f"{k:60s} {tuple(v.shape)} {v.dtype}")
        print(f"{len(ckpt)} tensors, {sum(v.numel() for v in ckpt.values()) / 1e6:.2f} M parameters")
        return 0
    from .pipeline import ISegPipeline
    params = {"loftup": {"upsampler_path": None, "n_dim": a.n_dim}, "lift": {"lift_path": None, "n_dim": a.n_dim, "patch": 14},
              "jbu_featup": {"backbone_type": "dinov2"}}.get(a.upsampler, {})
    pipe = ISegPipeline(a.upsampler, params)
    missing, unexpected, shapes = check(pipe, ckpt)
    for k in missing:
        print("missing   ", k)
    for k in unexpected:
        print("unexpected", k)
    for k, got, want in shapes:
        print("shape     ", k, got, "!=", want)
    ok = not missing and not shapes
    print("OK: the checkpoint restores every trainable tensor" if ok else "the checkpoint does NOT fit this configuration")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
