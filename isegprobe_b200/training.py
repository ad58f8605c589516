"""Training step of the IS model around the hot path (core/training/trainer.py:377-477, 213-226):

    features (click maps -> trainable click embedding -> frozen ViT -> frozen upsampler)
      -> ConvSegHead forward / backward on libisp_b200 (heads._ConvHeadFn)
      -> NormalizedFocalLossSigmoid (core/training/losses.py:42-109; stays torch)
      -> ONE all-reduce of the flat gradient arena (dist.FlatGradArena; DDP semantics: mean over ranks,
         core/utils/distributed.py + trainer.py:144-149) -> Adam(lr 5e-5) as models/defaults.py:103-107.

The click embedding (`embed_coords`) is trained THROUGH the frozen backbone and upsampler in the reference
(DINOv2.py:518-523).  Both backbones (DINOv2, MaskCLIP) and every upsampler (identity = the "noup" configs, bilinear,
torch's nearest / bicubic, the FeatUp JBU stack, LoftUp, LiFT) have their activation backward on libisp_b200, so the
step trains `embed_coords` like the reference.  `train_embedding=False` keeps it frozen (features under
no_grad, head forward / backward only).

Module modes.  The reference's trainer calls `self.net.train()` on the WHOLE model (trainer.py:213-214), which also flips
the frozen upsampler: LoftUp's first_conv BatchNorm then uses batch statistics and updates its running statistics
(SURVEY Q7).  `frozen_train_mode=True` reproduces that (`pipeline.train()`; LoftUpUpsampler honours `self.training`,
pinned to the reference module in train() by tests/golden/loftup_train_28x42.npz; JBUFeatUpUpsampler draws FeatUp's three
Dropout2d masks and applies them in forward and backward).  LiFT's BatchNorm batch statistics are NOT modelled (a warning is
issued; running statistics are used).  With
`frozen_train_mode=False` every frozen module runs in eval() and only the head (no mode-dependent layers) is in train()."""
import torch

from . import dist as idist


def normalized_focal_loss(pred: torch.Tensor, label: torch.Tensor, alpha: float = 0.5, gamma: float = 2.0,
                          eps: float = 1e-12, ignore_label: int = -1, weight: float = 1.0) -> torch.Tensor:
    """NormalizedFocalLossSigmoid.forward (losses.py:42-109) with its defaults (from_sigmoid=False,
    detach_delimeter=True, max_mult=-1, size_average=True): per-sample loss [B].  The reference's
    running `_k_sum` / `_m_max` statistics are logging only and need a host sync each step; omitted."""
    one_hot = label > 0.5
    sample_weight = label != ignore_label
    pred = torch.sigmoid(pred)
    alpha_t = torch.where(one_hot, alpha * sample_weight, (1 - alpha) * sample_weight)
    pt = torch.where(sample_weight, 1.0 - torch.abs(label - pred), torch.ones_like(pred))
    beta = (1 - pt) ** gamma
    sw_sum = torch.sum(sample_weight, dim=(-2, -1), keepdim=True)
    beta_sum = torch.sum(beta, dim=(-2, -1), keepdim=True)
    mult = (sw_sum / (beta_sum + eps)).detach()
    beta = beta * mult
    loss = -alpha_t * beta * torch.log(torch.clamp_max(pt + eps, 1.0))
    loss = weight * (loss * sample_weight)
    dims = tuple(range(1, loss.dim()))
    bsum = torch.sum(sample_weight, dim=dims)
    return torch.sum(loss, dim=dims) / (bsum + eps)


def get_next_points(pred: torch.Tensor, gt: torch.Tensor, points: torch.Tensor, click_indx: int,
                    pred_thresh: float = 0.49) -> torch.Tensor:
    """Click simulation of the training loop (core/training/trainer.py:577-618), host-side like the reference (OUT OF SCOPE
    to accelerate; in scope to run so that config 5 is measured as the reference runs it): per image, the
    chamfer distance transform (cv2 DIST_L2, 5x5 mask) of the zero-padded false-negative / false-positive regions of the
    current prediction, a uniformly random pixel of the inner half of the larger error region, written into the click
    slot `num_points - click_indx` (positive) or `2 * num_points - click_indx` (negative).  Draws from numpy's global RNG
    exactly like the reference."""
    import cv2
    import numpy as np
    assert click_indx > 0
    pred_np = pred.detach().float().cpu().numpy()[:, 0, :, :]
    gt_np = gt.detach().cpu().numpy()[:, 0, :, :] > 0.5
    fn_mask = np.logical_and(gt_np, pred_np < pred_thresh)
    fp_mask = np.logical_and(np.logical_not(gt_np), pred_np > pred_thresh)
    fn_mask = np.pad(fn_mask, ((0, 0), (1, 1), (1, 1)), "constant").astype(np.uint8)
    fp_mask = np.pad(fp_mask, ((0, 0), (1, 1), (1, 1)), "constant").astype(np.uint8)
    num_points = points.size(1) // 2
    points = points.clone()
    for b in range(fn_mask.shape[0]):
        fn_dt = cv2.distanceTransform(fn_mask[b], cv2.DIST_L2, 5)[1:-1, 1:-1]
        fp_dt = cv2.distanceTransform(fp_mask[b], cv2.DIST_L2, 5)[1:-1, 1:-1]
        fn_max, fp_max = np.max(fn_dt), np.max(fp_dt)
        is_positive = fn_max > fp_max
        dt = fn_dt if is_positive else fp_dt
        indices = np.argwhere(dt > max(fn_max, fp_max) / 2.0)
        if len(indices) > 0:
            coords = indices[np.random.randint(0, len(indices))]
            slot = (num_points if is_positive else 2 * num_points) - click_indx
            points[b, slot, 0] = float(coords[0])
            points[b, slot, 1] = float(coords[1])
            points[b, slot, 2] = float(click_indx)
    return points


class HeadTrainer:
    """One process per GPU; each rank steps on its own shard of the global batch."""

    def __init__(self, pipeline, lr: float = 5e-5, betas=(0.9, 0.999), eps: float = 1e-8, train_embedding: bool = True,
                 frozen_train_mode: bool = True):
        self.pipe = pipeline
        self.frozen_train_mode = frozen_train_mode
        assert pipeline.head is not None, "the pipeline was built without a head"
        self.train_embedding = train_embedding and (pipeline.upsampler_type in ("identity", "bilinear", "nearest", "bicubic", "jbu_featup", "loftup", "lift")
                                and hasattr(pipeline.backbone, "_backward_impl"))
        for p in pipeline.embed_coords.parameters():
            p.requires_grad = self.train_embedding  # see module docstring
        self.params = [p for p in pipeline.head.parameters() if p.requires_grad]
        if self.train_embedding:
            self.params += list(pipeline.embed_coords.parameters())
        self.arena = idist.FlatGradArena(self.params)
        self.comm_events = None  # bench.py: a list here makes every step record CUDA events around the all-reduce
        self.opt = torch.optim.Adam(self.params, lr=lr, betas=betas, eps=eps)

    def simulate_clicks(self, image: torch.Tensor, points: torch.Tensor, gt_mask: torch.Tensor, rounds: int):
        """The no-grad click-simulation rounds in front of the graded forward (trainer.py:399-431): the model in eval(),
        forward on (RGB || previous output), sigmoid, `get_next_points` on the host; returns the 4-channel image with the last
        prediction as previous mask and the extended click tensor.  `image`: [b,3|4,H,W] (a 4th channel is replaced: the
        reference starts from an all-zero previous output)."""
        pipe = self.pipe
        rgb = image[:, :3]
        prev = torch.zeros_like(rgb[:, :1])
        with torch.no_grad():
            for click_indx in range(rounds):
                pipe.eval()
                prev = torch.sigmoid(pipe(torch.cat((rgb, prev), dim=1), points)["instances"].float())
                points = get_next_points(prev, gt_mask, points, click_indx + 1).to(points.device)
        return torch.cat((rgb, prev), dim=1), points

    def step(self, image: torch.Tensor, points: torch.Tensor, gt_mask: torch.Tensor, click_sim_rounds: int = 0) -> torch.Tensor:
        """image [b,4,H,W] (RGB + previous mask), points [b,2P,3], gt_mask [b,1,H,W] in {0,1,-1}.
        click_sim_rounds > 0 runs that many click-simulation rounds first (the reference draws random.randint(0, 3) per batch,
        trainer.py:400-402).  Returns the (detached) mean loss of this rank's shard."""
        pipe = self.pipe
        if click_sim_rounds > 0:
            image, points = self.simulate_clicks(image, points, gt_mask, click_sim_rounds)
        if self.frozen_train_mode:
            pipe.train()  # trainer.py:213-214: the whole model, frozen modules included
        else:
            pipe.eval()
            pipe.head.train()
        if self.train_embedding:  # gradients flow head -> resize -> frozen ViT -> click embedding
            logits = pipe(image, points)["instances"]
        else:
            with torch.no_grad():
                feats = pipe.features(image, points)
            logits = pipe.head(feats)
        loss = normalized_focal_loss(logits, gt_mask).mean()
        self.arena.zero_()
        loss.backward()
        if self.comm_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.arena.all_reduce_mean()
            e1.record()
            self.comm_events.append((e0, e1))
        else:
            self.arena.all_reduce_mean()
        self.opt.step()
        return loss.detach()
