"""Model assembly of the hot path, mirroring iSegBaseModel.forward
(core/model/iseg_base_model.py:67-110) and iSegProbeModel.backbone_forward
(core/model/iseg_probe_model.py:110-134) with this package's modules, plus the
registry hook that drops them into the reference's ModelBuilder
(core/utils/model_builder.py:59-95).

The reference's own `iSegProbeModel` keeps working unchanged once `install_into_reference()`
has swapped the registry entries; `ISegPipeline` is the same call sequence for users (and
bench.py) that do not have the reference's Python dependencies (omegaconf, mmcv, timm)."""
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .featurizers import DINOFeaturizer, DINOv2Featurizer, MaskCLIPFeaturizer, PatchEmbed
from .simple_vit import SimpleViTFeaturizer
from .heads import HEAD_REGISTRY
from .upsamplers import UPSAMPLER_REGISTRY, bilinear_align_corners_nhwc, to_nhwc_f32

FEATURIZER_REGISTRY = {"dinov2": DINOv2Featurizer, "maskclip": MaskCLIPFeaturizer, "patch_embedding": PatchEmbed,
                       "simple_vit": SimpleViTFeaturizer}


def install_into_reference(featurizers: bool = False) -> None:
    """Swap the reference's registry entries for the B200-native modules (same keys), so
    `ModelBuilder.load_upsampler/load_head` construct ours and `iSegBaseModel` builds our `DistMaps`.  Call before
    building a model (before or after importing the reference's `core.model` package: both orders work).
    featurizers=True also rebinds the featurizer classes `ModelBuilder.load_featurizer` and `iSegProbeModel` look up by
    name (core/utils/model_builder.py:27-49, core/model/iseg_probe_model.py:93-103): the frozen backbones and the two
    click encoders.  Ours keep the constructor keywords and the state-dict keys of `.model` but do not download
    weights: load the torch.hub / CLIP / timm state dict into `backbone.model` afterwards."""
    import sys

    import core.model.heads as ref_heads
    import core.model.upsamplers as ref_up

    ref_up.UPSAMPLER_REGISTRY.update(UPSAMPLER_REGISTRY)
    ref_heads.HEAD_REGISTRY.update(HEAD_REGISTRY)
    if featurizers:
        import core.model.iseg_probe_model as ref_model
        import core.utils.model_builder as ref_mb
        for cls in (DINOv2Featurizer, MaskCLIPFeaturizer, DINOFeaturizer, SimpleViTFeaturizer):
            setattr(ref_mb, cls.__name__, cls)
        ref_model.PatchEmbed = PatchEmbed
    # DistMaps: `iSegBaseModel.__init__` constructs the name it bound with `from core.model.ops import DistMaps` at
    # import time (core/model/iseg_base_model.py:9,60-65), so rebinding `core.model.ops.DistMaps` alone is not enough:
    # every already-imported module that holds the reference class under that name is rebound too.
    import core.model.iseg_base_model  # noqa: F401  (make sure the consumer exists before the scan)
    import core.model.ops as ref_ops
    ref_cls = ref_ops.DistMaps
    ref_ops.DistMaps = ops.DistMaps
    for name, mod in list(sys.modules.items()):
        if mod is not None and name.startswith("core.") and getattr(mod, "DistMaps", None) is ref_cls:
            mod.DistMaps = ops.DistMaps


def swap_modules(model: nn.Module) -> nn.Module:
    """For a reference model that was built BEFORE `install_into_reference()` (e.g. restored by
    core/utils/serialization.py:61-91): replace its `dist_maps` in place with ours (same constructor arguments,
    core/model/iseg_base_model.py:60-65).  Registry-built parts (upsampler / head) cannot be converted after the
    fact -- rebuild the model after installing."""
    dm = getattr(model, "dist_maps", None)
    if dm is not None and not isinstance(dm, ops.DistMaps):
        model.dist_maps = ops.DistMaps(norm_radius=dm.norm_radius, spatial_scale=dm.spatial_scale,
                                       cpu_mode=dm.cpu_mode, use_disks=dm.use_disks)
    return model


class ISegPipeline(nn.Module):
    """image [B,3|4,H,W] in [0,1] + points [B,2P,3] -> {"instances": [B,1,H,W] logits}.

    Constructor keys follow models/sbd/dinov2/patch-embed_*.py: upsampler/head `type` strings
    and `params` dicts, `use_disks`, `norm_radius`, `with_prev_mask`."""

    def __init__(self, upsampler_type: str = "loftup", upsampler_params: Optional[dict] = None,
                 head_type: str = "convhead", head_params: Optional[dict] = None, backbone_dim: int = 384,
                 patch: int = 14, use_disks: bool = True, norm_radius: int = 5, with_prev_mask: bool = True,
                 with_head: bool = True, backbone: str = "dinov2", embed_coords_type: str = "patch_embedding",
                 embed_coords_params: Optional[dict] = None, feats_injection_mode: str = "before_backbone"):
        super().__init__()
        if backbone not in ("dinov2", "maskclip", "vit"):
            raise ValueError(f"Unknown backbone type: {backbone}")
        if upsampler_type not in UPSAMPLER_REGISTRY:
            raise ValueError(f"Unknown upsampler type: {upsampler_type}")  # model_builder.py:64-65
        self.use_disks, self.norm_radius, self.with_prev_mask = use_disks, norm_radius, with_prev_mask
        self.upsampler_type = upsampler_type
        if embed_coords_type not in ("patch_embedding", "simple_vit"):
            raise ValueError(f"Unsupported backbone type: {embed_coords_type}")  # model_builder.py:50-51
        if backbone == "maskclip":  # models/sbd/maskclip/*.py: ViT-B/16, 768-wide tokens, 512-d features
            self.backbone = MaskCLIPFeaturizer("ViT-B/16", feats_injection_mode)
            patch, backbone_dim, embed_dim = 16, 512, 768
        elif backbone == "vit":  # models/sbd/vit/patch-embed_noup.py: timm ViT-S/16, key features
            self.backbone = DINOFeaturizer("vit_small_patch16_224", 16, "key", feats_injection_mode)
            patch, embed_dim = 16, backbone_dim
        else:
            self.backbone = DINOv2Featurizer("dinov2_vits14", feats_injection_mode)
            embed_dim = backbone_dim
        if embed_coords_type == "simple_vit":  # models/sbd/dinov2/simple-vit_noup.py (late injection: backbone_dim wide)
            ep = dict(img_size=[448, 448], patch_size=(patch, patch), embed_dim=embed_dim, depth=6, heads=8, mlp_dim=2048,
                      channels=3 if with_prev_mask else 2, dim_head=64)
            ep.update(embed_coords_params or {})
            self.embed_coords = SimpleViTFeaturizer(image_size=ep["img_size"], patch_size=ep["patch_size"], dim=ep["embed_dim"],
                                                    depth=ep["depth"], heads=ep["heads"], mlp_dim=ep["mlp_dim"],
                                                    channels=ep["channels"], dim_head=ep["dim_head"])
        else:
            self.embed_coords = PatchEmbed((448, 448), (patch, patch), 3 if with_prev_mask else 2, embed_dim)
        self.upsampler = UPSAMPLER_REGISTRY[upsampler_type](**(upsampler_params or {}))
        self.head = None
        if with_head:
            if head_type not in HEAD_REGISTRY:
                raise ValueError(f"Unknown head type: {head_type}")  # model_builder.py:81-82
            self.head = HEAD_REGISTRY[head_type](**(head_params or {"in_channels": backbone_dim, "num_layers": 2,
                                                                    "num_classes": 1}))
        if self.head is not None and hasattr(self.upsampler, "out_dtype") and hasattr(self.head, "_features_bf16"):
            self.upsampler.out_dtype = torch.bfloat16  # the head's own input format: no fp32 round trip in between
        for m in (self.backbone, self.upsampler):  # frozen, as ModelBuilder(freeze=True)
            for p in m.parameters():
                p.requires_grad = False

    def features(self, image: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
        """Up to the upsampler output, resized to the image size: [B,C,H,W] (NHWC memory)."""
        norm_img, coord = ops.prepare_input(image, points, self.norm_radius, 1.0, self.use_disks)
        emb = self.embed_coords(coord)
        lr = self.backbone(norm_img, emb)
        if hasattr(self.upsampler, "forward_resized"):  # upsampler + the resize below in one go (same result)
            hr = self.upsampler.forward_resized(lr, norm_img, tuple(norm_img.shape[2:]))
        else:
            hr = self.upsampler(source=lr, guidance=norm_img)
        if self.upsampler_type != "identity" and tuple(hr.shape[2:]) != tuple(norm_img.shape[2:]):
            hr = bilinear_align_corners_nhwc(to_nhwc_f32(hr), tuple(norm_img.shape[2:])).permute(0, 3, 1, 2)
        return hr

    def _graphed(self, kind: str, image: torch.Tensor, points: torch.Tensor, slot: int,
                 h2d_stream: Optional[torch.cuda.Stream]) -> torch.Tensor:
        """One CUDA graph per (kind, input shapes / dtypes, slot): static input buffers, replay, result in a buffer the
        graph owns.  Inputs may be (pinned) host tensors: they are copied straight into the static buffers."""
        dev = next(self.backbone.parameters()).device
        key = (kind, tuple(image.shape), tuple(points.shape), image.dtype, points.dtype, slot)
        graphs = self.__dict__.setdefault("_graphs", {})
        entry = graphs.get(key)
        if entry is None:
            from . import _lib
            fn = self.features if kind == "features" else (lambda i, p: self.forward(i, p)["instances"])
            s_img, s_pts = image.to(dev, copy=True), points.to(dev, copy=True)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():  # warm-up: weight packing, host-scalar caches, kernel attributes
                for _ in range(2):
                    fn(s_img, s_pts)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(graph), torch.no_grad():
                out = fn(s_img, s_pts)
            entry = (graph, s_img, s_pts, out, _lib.launch_count() - l0, torch.cuda.Event())
            graphs[key] = entry
            self.__dict__["_last_graph_launches"] = entry[4]
        graph, s_img, s_pts, out, _, done = entry
        cur = torch.cuda.current_stream()
        if h2d_stream is None:
            s_img.copy_(image, non_blocking=True)
            s_pts.copy_(points, non_blocking=True)
        else:
            h2d_stream.wait_event(done)  # the slot's previous replay has consumed its inputs
            with torch.cuda.stream(h2d_stream):
                s_img.copy_(image, non_blocking=True)
                s_pts.copy_(points, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(h2d_stream)
            cur.wait_event(ready)
        graph.replay()
        done.record(cur)
        return out

    def features_graphed(self, image: torch.Tensor, points: torch.Tensor, slot: int = 0,
                         h2d_stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """`features` replayed from a CUDA graph (one capture per input shape): the ~700 kernel launches of a
        step -- most of them tens of microseconds in the ViT -- are issued by the GPU front-end instead of
        ~700 Python/ctypes calls, which otherwise leave the device waiting between the small kernels.
        Inference only (no autograd); the result lives in a buffer owned by the graph and is overwritten by
        the next call with the same shapes and the same `slot`.
        Pipelined serving: alternate `slot` 0/1 and pass a copy stream -- the host->device copy of the next request then
        runs on `h2d_stream` while the previous request's graph executes (each slot has its own static input / output
        buffers; the copy waits for the slot's previous replay, the replay waits for the copy)."""
        assert not torch.is_grad_enabled() or not any(p.requires_grad for p in self.embed_coords.parameters()) or \
            not self.training, "features_graphed is an inference path"
        return self._graphed("features", image, points, slot, h2d_stream)

    def forward_graphed(self, image: torch.Tensor, points: torch.Tensor, slot: int = 0,
                        h2d_stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """`forward(image, points)["instances"]` replayed from a CUDA graph (one capture per input shape / dtype / slot).
        Inference only; the logits live in a buffer owned by the graph and are overwritten by the next call with the
        same shapes and slot.  Used by the NoC evaluation loop (FixedSizePredictor(use_graph=True): a click is ~430
        launches, most of them tens of microseconds) and by bench.py's end-to-end leg (host buffers in, logits out,
        requests pipelined two deep exactly like `features_graphed`)."""
        assert self.head is not None
        return self._graphed("forward", image, points, slot, h2d_stream)

    def graphed_launches(self) -> int:
        """Kernel launches inside the most recently captured graph (bench.py's gpu_launches claim)."""
        return self.__dict__.get("_last_graph_launches", 0)

    def forward(self, image: torch.Tensor, points: torch.Tensor) -> Dict:
        hr = self.features(image, points)
        if self.head is None:
            return {"features": hr, "instances": None, "instances_aux": None}
        logits = self.head(hr)
        if tuple(logits.shape[2:]) != tuple(image.shape[2:]):
            logits = bilinear_align_corners_nhwc(to_nhwc_f32(logits), tuple(image.shape[2:])).permute(0, 3, 1, 2)
        return {"instances": logits, "instances_aux": None}
