"""isegprobe_b200 -- B200-native (sm_100a) drop-in for iSegProbe's dense hot path:
click-map encoding -> frozen ViT -> feature upsampler (LoftUp / FeatUp-JBU / LiFT) -> IS head,
behind the reference's own plugin API (UPSAMPLER_REGISTRY / HEAD_REGISTRY / DistMaps).
Host code is Python/PyTorch; all compute is hand-written CUDA in libisp_b200.so
reached through a C ABI (include/isp_b200.h).  No CPU fallback.
"""
from . import _lib  # noqa: F401
from .ops import BatchImageNormalize, DistMaps, prepare_input  # noqa: F401
from .upsamplers import (  # noqa: F401
    UPSAMPLER_REGISTRY,
    BaseUpsampler,
    BicubicUpsampler,
    BilinearUpsampler,
    IdentityUpsampler,
    JBUFeatUpUpsampler,
    NearestUpsampler,
)
from .loftup import LoftUpUpsampler  # noqa: E402,F401
from .lift import LiFTUpsampler  # noqa: E402,F401

UPSAMPLER_REGISTRY["loftup"] = LoftUpUpsampler
UPSAMPLER_REGISTRY["lift"] = LiFTUpsampler

from .heads import HEAD_REGISTRY, ConvSegHead, SimpleClassifierHead, SimpleConvSegHead  # noqa: E402,F401
from .featurizers import DINOFeaturizer, DINOv2Featurizer, MaskCLIPFeaturizer, PatchEmbed  # noqa: E402,F401
from .pipeline import ISegPipeline, install_into_reference  # noqa: E402,F401

__version__ = "0.1.0"
