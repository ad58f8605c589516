"""NoC evaluation loop around the hot path (BASELINE config 4):
the click simulator, the predictor with its inference transforms, per-sample evaluation and the
NoC metric of the reference, restated so that the loop can run on this package's pipeline without
the reference's Python dependencies, and SHARDED over ranks (our addition: the reference evaluates
on one GPU, core/inference/utils.py:270-274).

  Clicker            core/inference/clicker.py:10-140       (host, cv2.distanceTransform)
  FixedSizePredictor core/inference/predictors/base_predictor.py:20-235 with the transform stack of
                     eval_mode="fixedNNN": ZoomIn(skip_clicks=-1, target_size=(N, N))
                     (core/inference/transforms/zoom_in.py:13-253, utils.py:300-318), SigmoidForPred
                     (base_transform.py:31-48) and AddHorizontalFlip (flip.py:13-45)
  evaluate_sample    core/inference/evaluation.py:43-88
  compute_noc_metric core/inference/utils.py:123-146

The model is any callable `net(image [B,4,H,W], points [B,2P,3]) -> {"instances": logits}` with a
`with_prev_mask` attribute: `ISegPipeline`, or the reference's own `iSegProbeModel`.  The crop / resize /
flip glue of FixedSizePredictor is torch on the model's device -- the structure of the reference's driver, kept as the
pinned host path; `DeviceNoCEvaluator` (below) is the device-resident loop of rows f1 / f2: transforms, IoU and the click
simulator run in libisp_b200 kernels and two samples are interleaved per rank."""
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import dist as idist


# ------------------------------------------------------------------ clicker (clicker.py)
class Click:
    def __init__(self, is_positive: bool, coords: Tuple[float, float], indx: Optional[int] = None) -> None:
        self.is_positive, self.coords, self.indx = is_positive, coords, indx

    @property
    def coords_and_indx(self):
        return (*self.coords, self.indx)

    def copy(self, **kwargs) -> "Click":
        c = Click(self.is_positive, self.coords, self.indx)
        for k, v in kwargs.items():
            setattr(c, k, v)
        return c


class Clicker:
    """Next click = the point of the largest error region farthest from its border
    (clicker.py:58-91): distance transform of the false-negative / false-positive masks (zero padded by
    one pixel), already-clicked pixels excluded, first maximum in row-major order."""

    def __init__(self, gt_mask: np.ndarray, ignore_label: int = -1, click_indx_offset: int = 0) -> None:
        import cv2  # host-side, as in the reference
        self._cv2 = cv2
        self.click_indx_offset = click_indx_offset
        self.gt_mask = gt_mask == 1
        self.not_ignore_mask = gt_mask != ignore_label
        self.reset_clicks()

    def reset_clicks(self) -> None:
        self.not_clicked_map = np.ones_like(self.gt_mask, dtype=bool)
        self.num_pos_clicks = self.num_neg_clicks = 0
        self.clicks_list: List[Click] = []

    def get_clicks(self, clicks_limit: Optional[int] = None) -> List[Click]:
        return self.clicks_list[:clicks_limit]

    def make_next_click(self, pred_mask: np.ndarray) -> None:
        cv2 = self._cv2
        fn = np.logical_and(np.logical_and(self.gt_mask, np.logical_not(pred_mask)), self.not_ignore_mask)
        fp = np.logical_and(np.logical_and(np.logical_not(self.gt_mask), pred_mask), self.not_ignore_mask)
        fn = np.pad(fn, ((1, 1), (1, 1)), "constant")
        fp = np.pad(fp, ((1, 1), (1, 1)), "constant")
        fn_dt = cv2.distanceTransform(fn.astype(np.uint8), cv2.DIST_L2, 0)[1:-1, 1:-1] * self.not_clicked_map
        fp_dt = cv2.distanceTransform(fp.astype(np.uint8), cv2.DIST_L2, 0)[1:-1, 1:-1] * self.not_clicked_map
        fn_max, fp_max = np.max(fn_dt), np.max(fp_dt)
        is_positive = fn_max > fp_max
        ys, xs = np.where(fn_dt == fn_max) if is_positive else np.where(fp_dt == fp_max)
        self.add_click(Click(is_positive=bool(is_positive), coords=(ys[0], xs[0])))

    def add_click(self, click: Click) -> None:
        click.indx = self.click_indx_offset + self.num_pos_clicks + self.num_neg_clicks
        if click.is_positive:
            self.num_pos_clicks += 1
        else:
            self.num_neg_clicks += 1
        self.clicks_list.append(click)
        self.not_clicked_map[click.coords[0], click.coords[1]] = False

    def __len__(self) -> int:
        return len(self.clicks_list)


# ------------------------------------------------------------------ bbox helpers (core/utils/misc.py:71-120)
def _bbox_from_mask(mask: np.ndarray):
    rows, cols = np.any(mask, axis=1), np.any(mask, axis=0)
    rmin, rmax = np.where(rows)[0][[0, -1]]
    cmin, cmax = np.where(cols)[0][[0, -1]]
    return rmin, rmax, cmin, cmax


def _expand_bbox(bbox, ratio: float, min_crop_size: Optional[float]):
    rmin, rmax, cmin, cmax = bbox
    rc, cc = 0.5 * (rmin + rmax), 0.5 * (cmin + cmax)
    h, w = ratio * (rmax - rmin + 1), ratio * (cmax - cmin + 1)
    if min_crop_size is not None:
        h, w = max(h, min_crop_size), max(w, min_crop_size)
    return int(round(rc - 0.5 * h)), int(round(rc + 0.5 * h)), int(round(cc - 0.5 * w)), int(round(cc + 0.5 * w))


def _segments_iou(s1, s2) -> float:
    (a, b), (c, d) = s1, s2
    return max(0, min(b, d) - max(a, c) + 1) / max(1e-6, max(b, d) - min(a, c) + 1)


def _bbox_iou(b1, b2) -> float:
    return _segments_iou(b1[:2], b2[:2]) * _segments_iou(b1[2:4], b2[2:4])


# ------------------------------------------------------------------ ZoomIn (zoom_in.py)
class ZoomIn:
    def __init__(self, target_size=400, skip_clicks: int = 1, expansion_ratio: float = 1.4, min_crop_size: int = 200,
                 recompute_thresh_iou: float = 0.5, prob_thresh: float = 0.50) -> None:
        self.target_size, self.min_crop_size, self.skip_clicks = target_size, min_crop_size, skip_clicks
        self.expansion_ratio, self.recompute_thresh_iou, self.prob_thresh = expansion_ratio, recompute_thresh_iou, prob_thresh
        self.reset()

    def reset(self) -> None:
        self._input_image_shape = None
        self._prev_probs = None
        self._object_roi = None
        self._roi_image = None
        self.image_changed = False

    def _object_roi_of(self, pred_mask: np.ndarray, clicks: Sequence[Click]):
        pred_mask = pred_mask.copy()
        for c in clicks:
            if c.is_positive:
                pred_mask[int(c.coords[0]), int(c.coords[1])] = 1
        bbox = _expand_bbox(_bbox_from_mask(pred_mask), self.expansion_ratio, self.min_crop_size)
        h, w = pred_mask.shape
        return max(0, bbox[0]), min(h - 1, bbox[1]), max(0, bbox[2]), min(w - 1, bbox[3])

    @staticmethod
    def _check_roi(roi, clicks: Sequence[Click]) -> bool:
        for c in clicks:
            if c.is_positive:
                if c.coords[0] < roi[0] or c.coords[0] >= roi[1] or c.coords[1] < roi[2] or c.coords[1] >= roi[3]:
                    return False
        return True

    def transform(self, image_nd: torch.Tensor, clicks_lists):
        imgs, out_clicks = [], []
        for b in range(len(clicks_lists)):
            im, cl = self._transform(image_nd[b:b + 1], clicks_lists[b])
            imgs.append(im)
            out_clicks.append(cl)
        return torch.cat(imgs, dim=0), out_clicks

    def _transform(self, image_nd: torch.Tensor, clicks: List[Click]):
        self.image_changed = False
        if len(clicks) <= self.skip_clicks:
            return image_nd, clicks
        self._input_image_shape = image_nd.shape
        current = None
        if self._prev_probs is not None:
            mask = (self._prev_probs > self.prob_thresh)[0, 0]
            if mask.sum() > 0:
                current = self._object_roi_of(mask, clicks)
        if current is None:
            if self.skip_clicks >= 0:
                return image_nd, clicks
            current = 0, image_nd.shape[2] - 1, 0, image_nd.shape[3] - 1
        update = (self._object_roi is None or not self._check_roi(self._object_roi, clicks)
                  or _bbox_iou(current, self._object_roi) < self.recompute_thresh_iou)
        if update:
            self._object_roi = current
            self.image_changed = True
        rmin, rmax, cmin, cmax = self._object_roi
        hh, ww = rmax - rmin + 1, cmax - cmin + 1
        if isinstance(self.target_size, tuple):
            nh, nw = self.target_size
        else:
            sc = self.target_size / max(hh, ww)
            nh, nw = int(round(hh * sc)), int(round(ww * sc))
        self._roi_image = F.interpolate(image_nd[:, :, rmin:rmax + 1, cmin:cmax + 1], size=(nh, nw), mode="bilinear",
                                        align_corners=True)
        ch, cw = self._roi_image.shape[2:]
        tclicks = [c.copy(coords=(ch * (c.coords[0] - rmin) / (rmax - rmin + 1), cw * (c.coords[1] - cmin) / (cmax - cmin + 1)))
                   for c in clicks]
        return self._roi_image, tclicks

    def inv_transform(self, prob_map: torch.Tensor) -> torch.Tensor:
        return torch.cat([self._inv_transform(prob_map[b:b + 1]) for b in range(prob_map.shape[0])], dim=0)

    def _inv_transform(self, prob_map: torch.Tensor) -> torch.Tensor:
        if self._object_roi is None:
            self._prev_probs = prob_map.cpu().numpy()
            return prob_map
        rmin, rmax, cmin, cmax = self._object_roi
        prob_map = F.interpolate(prob_map, size=(rmax - rmin + 1, cmax - cmin + 1), mode="bilinear", align_corners=True)
        if self._prev_probs is not None:
            new = torch.zeros(*self._prev_probs.shape, device=prob_map.device, dtype=prob_map.dtype)
            new[:, :, rmin:rmax + 1, cmin:cmax + 1] = prob_map
        else:
            new = prob_map
        self._prev_probs = new.cpu().numpy()
        return new

    def check_possible_recalculation(self) -> bool:
        if self._prev_probs is None or self._object_roi is not None or self.skip_clicks > 0:
            return False
        mask = (self._prev_probs > self.prob_thresh)[0, 0]
        if mask.sum() > 0:
            roi = self._object_roi_of(mask, [])
            full = (0, self._input_image_shape[2] - 1, 0, self._input_image_shape[3] - 1)
            if _bbox_iou(roi, full) < 0.50:
                return True
        return False


# ------------------------------------------------------------------ predictor (base_predictor.py)
class FixedSizePredictor:
    """BasePredictor with the transform stack [ZoomIn, SigmoidForPred, AddHorizontalFlip]."""

    def __init__(self, net, device, target_size=(448, 448), with_flip: bool = True, net_clicks_limit: Optional[int] = None,
                 zoom_in: Optional[ZoomIn] = "default", use_graph: bool = False, graph_clicks: int = 24) -> None:
        """use_graph: run the network through ISegPipeline.forward_graphed.  The click tensor is then padded to
        `graph_clicks` entries per polarity with (-1,-1,-1) rows -- invalid clicks, which DistMaps ignores
        (core/model/ops.py:44-52), so the prediction is unchanged -- to keep one graph for every click count."""
        self.net, self.device = net, device
        self.use_graph, self.graph_clicks = use_graph and hasattr(net, "forward_graphed"), graph_clicks
        self.with_flip, self.net_clicks_limit = with_flip, net_clicks_limit
        self.zoom_in = ZoomIn(skip_clicks=-1, target_size=tuple(target_size)) if zoom_in == "default" else zoom_in
        self.original_image = None
        self.prev_prediction = None

    def set_input_image(self, image) -> None:
        """image: HxWx3 uint8 / float array (ToTensor semantics: uint8 -> [0,1]) or a [3,H,W] / [1,3,H,W] tensor."""
        if not isinstance(image, torch.Tensor):
            arr = np.asarray(image)
            t = torch.from_numpy(np.ascontiguousarray(arr.transpose(2, 0, 1)))
            image = t.float().div(255) if arr.dtype == np.uint8 else t.float()
        if self.zoom_in is not None:
            self.zoom_in.reset()
        self.original_image = image.to(self.device)
        if self.original_image.dim() == 3:
            self.original_image = self.original_image.unsqueeze(0)
        self.prev_prediction = torch.zeros_like(self.original_image[:, :1])

    def get_points_nd(self, clicks_lists) -> torch.Tensor:
        num_pos = [sum(c.is_positive for c in cl) for cl in clicks_lists]
        num_neg = [len(cl) - p for cl, p in zip(clicks_lists, num_pos)]
        n = max(num_pos + num_neg)
        if self.net_clicks_limit is not None:
            n = min(self.net_clicks_limit, n)
        n = max(1, n)
        if self.use_graph and n <= self.graph_clicks:
            n = self.graph_clicks
        total = []
        for cl in clicks_lists:
            cl = cl[: self.net_clicks_limit]
            pos = [c.coords_and_indx for c in cl if c.is_positive]
            neg = [c.coords_and_indx for c in cl if not c.is_positive]
            total.append(pos + (n - len(pos)) * [(-1, -1, -1)] + neg + (n - len(neg)) * [(-1, -1, -1)])
        return torch.tensor(np.asarray(total, dtype=np.float64), device=self.device)

    def get_prediction(self, clicker: Clicker) -> np.ndarray:
        clicks = clicker.get_clicks()
        image = self.original_image
        if getattr(self.net, "with_prev_mask", False):
            image = torch.cat((image, self.prev_prediction), dim=1)
        clicks_lists = [clicks]
        if self.zoom_in is not None:
            image, clicks_lists = self.zoom_in.transform(image, clicks_lists)
        if self.with_flip:  # flip.py:14-30
            width = image.shape[3]
            image = torch.cat([image, torch.flip(image, dims=[3])], dim=0)
            clicks_lists = clicks_lists + [[c.copy(coords=(c.coords[0], width - c.coords[1] - 1)) for c in cl]
                                           for cl in clicks_lists]
        points = self.get_points_nd(clicks_lists)
        if self.use_graph and points.shape[1] == 2 * self.graph_clicks:
            logits = self.net.forward_graphed(image, points)
        else:
            logits = self.net(image, points)["instances"]
        pred = F.interpolate(logits.float(), mode="bilinear", align_corners=True, size=image.shape[2:])
        if self.with_flip:  # inverse transforms run in reverse order: flip, sigmoid, zoom-in
            n = pred.shape[0] // 2
            pred = 0.5 * (pred[:n] + torch.flip(pred[n:], dims=[3]))
        pred = torch.sigmoid(pred)
        if self.zoom_in is not None:
            pred = self.zoom_in.inv_transform(pred)
            if self.zoom_in.check_possible_recalculation():
                return self.get_prediction(clicker)
        self.prev_prediction = pred
        return pred.cpu().numpy()[0, 0]


# ------------------------------------------------------------------ evaluation (evaluation.py, utils.py)
def get_iou(gt_mask: np.ndarray, pred_mask: np.ndarray, ignore_label: int = -1) -> float:
    keep = gt_mask != ignore_label
    obj = gt_mask == 1
    inter = np.logical_and(np.logical_and(pred_mask, obj), keep).sum()
    union = np.logical_and(np.logical_or(pred_mask, obj), keep).sum()
    return inter / union


def evaluate_sample(image, gt_mask: np.ndarray, predictor: FixedSizePredictor, max_iou_thr: float, pred_thr: float = 0.49,
                    min_clicks: int = 1, max_clicks: int = 20, callback: Optional[Callable] = None):
    clicker = Clicker(gt_mask=gt_mask)
    pred_mask = np.zeros_like(gt_mask)
    ious = []
    with torch.no_grad():
        predictor.set_input_image(image)
        for click_indx in range(max_clicks):
            clicker.make_next_click(pred_mask)
            pred_probs = predictor.get_prediction(clicker)
            pred_mask = pred_probs > pred_thr
            if callback is not None:
                callback(image, gt_mask, pred_probs, click_indx, clicker.clicks_list)
            iou = get_iou(gt_mask, pred_mask)
            ious.append(iou)
            if iou >= max_iou_thr and click_indx + 1 >= min_clicks:
                break
    return clicker.clicks_list, np.array(ious, dtype=np.float32), pred_probs


def compute_noc_metric(all_ious: Sequence[np.ndarray], iou_thrs: Sequence[float], max_clicks: int = 20):
    def _noc(arr, thr):
        vals = arr >= thr
        return np.argmax(vals) + 1 if np.any(vals) else max_clicks

    noc, noc_std, over = [], [], []
    for thr in iou_thrs:
        scores = np.array([_noc(a, thr) for a in all_ious], dtype=np.int_)
        noc.append(scores.mean())
        noc_std.append(scores.std())
        over.append((scores == max_clicks).sum())
    return noc, noc_std, over


def evaluate_dataset_sharded(samples: Sequence[Tuple[np.ndarray, np.ndarray]], predictor,
                             max_iou_thr: float = 1.01, max_clicks: int = 20, pred_thr: float = 0.49):
    """evaluate_dataset (evaluation.py:23-40) with the samples sharded round-robin over the ranks
    (dist.shard_indices) and the per-sample IoU curves gathered at the end (dist.gather_sample_results):
    every rank returns the full list, in dataset order, ready for compute_noc_metric.  IoU curves shorter
    than max_clicks (early exit at max_iou_thr) are padded with NaN for the gather and cut back after it."""
    n = len(samples)
    mine = idist.shard_indices(n)
    rows = torch.full((len(mine), max_clicks + 1), float("nan"), dtype=torch.float32)
    if isinstance(predictor, DeviceNoCEvaluator):  # the device-resident loop: this rank's samples, `lanes` at a time
        curves = [r[1] for r in predictor.evaluate([samples[i] for i in mine], max_iou_thr=max_iou_thr, pred_thr=pred_thr,
                                                    max_clicks=max_clicks)]
    else:
        curves = [evaluate_sample(samples[i][0], samples[i][1], predictor, max_iou_thr=max_iou_thr, pred_thr=pred_thr,
                                  max_clicks=max_clicks)[1] for i in mine]
    for j, ious in enumerate(curves):
        rows[j, 0] = len(ious)
        rows[j, 1:1 + len(ious)] = torch.from_numpy(ious)
    dev = predictor.device if idist.world() > 1 and torch.distributed.get_backend() == "nccl" else "cpu"
    full = idist.gather_sample_results(rows.to(dev), n).cpu()
    return [full[i, 1:1 + int(full[i, 0])].numpy() for i in range(n)]


# ------------------------------------------------------------------ device-resident loop
class _Lane:
    """One sample in flight: its image / ground truth / probability map / masks live on the device for the whole click loop."""

    def __init__(self, index, image, gt_mask, device):
        arr = np.asarray(image)
        t = torch.from_numpy(np.ascontiguousarray(arr.transpose(2, 0, 1)))
        self.index = index
        self.image = (t.float().div(255) if arr.dtype == np.uint8 else t.float()).to(device).contiguous()
        self.H, self.W = int(self.image.shape[1]), int(self.image.shape[2])
        self.gt = torch.from_numpy(np.ascontiguousarray(gt_mask.astype(np.int32))).to(device)
        self.prob = None  # previous probability map [H, W] (None = zeros: first click)
        self.prob_next = torch.empty(self.H, self.W, dtype=torch.float32, device=device)
        self.pred_mask = torch.zeros(self.H, self.W, dtype=torch.uint8, device=device)
        self.clicked = torch.zeros(self.H, self.W, dtype=torch.uint8, device=device)
        self.work = torch.empty(2 * self.H * self.W, dtype=torch.int32, device=device)
        self.result_d = torch.zeros(6, dtype=torch.int64, device=device)  # [best_fn, best_fp, 4 x packed int32 stats]
        self.result_h = torch.zeros(6, dtype=torch.int64).pin_memory()
        self.event = torch.cuda.Event()
        self.clicks: List[Click] = []
        self.ious: List[float] = []
        self.roi = None
        self.have_stats = False
        self.done = False


class DeviceNoCEvaluator:
    """evaluate_sample (core/inference/evaluation.py:43-88) for eval_mode "fixedNNN" with flip TTA, with everything between
    two network calls on the device: ZoomIn crop + resize + flip (`isp_zoom_in_fwd`), flip-average + sigmoid + un-zoom + threshold
    + IoU counts + bounding box (`isp_unzoom_probs`), and the click simulator (`isp_noc_next_click`).  Per click the host reads
    back 48 bytes (next click, IoU counts, bounding box) and does the ZoomIn ROI bookkeeping (integer / float scalar logic of
    zoom_in.py:51-104, kept verbatim); `lanes` samples are interleaved on one stream so that this read-back and the host
    logic of one sample overlap the forward of another.  Click sequences / IoU curves: identical to the host path
    (FixedSizePredictor + Clicker), see tests/test_gpu_noc_device.py."""

    def __init__(self, net, device, target_size=(448, 448), with_flip: bool = True, lanes: int = 2, graph_clicks: int = 24,
                 expansion_ratio: float = 1.4, min_crop_size: int = 200, recompute_thresh_iou: float = 0.5,
                 prob_thresh: float = 0.5):
        self.net, self.device = net, torch.device(device)
        self.S0, self.S1 = int(target_size[0]), int(target_size[1])
        self.with_flip, self.lanes, self.graph_clicks = with_flip, max(1, lanes), graph_clicks
        self.expansion_ratio, self.min_crop_size = expansion_ratio, min_crop_size
        self.recompute_thresh_iou, self.prob_thresh = recompute_thresh_iou, prob_thresh
        self.use_graph = hasattr(net, "forward_graphed")

    # -- device steps ---------------------------------------------------------------------------------------------
    def _enqueue_click(self, ln: _Lane):
        from . import _lib
        st = _lib.stream_ptr()
        _lib.call("isp_noc_next_click", _lib.dptr(ln.gt), _lib.dptr(ln.pred_mask), _lib.dptr(ln.clicked), ln.H, ln.W,
                  _lib.dptr(ln.work), _lib.dptr(ln.result_d), st)
        ln.result_h.copy_(ln.result_d, non_blocking=True)
        ln.event.record()

    def _roi_for(self, ln: _Lane, stats) -> Tuple[int, int, int, int]:
        """ZoomIn._transform's ROI bookkeeping (zoom_in.py:51-104) with skip_clicks = -1."""
        current = None
        if ln.have_stats and stats[2] > 0:  # mask = prev_probs > prob_thresh is not empty
            rmin, rmax, cmin, cmax = stats[3], stats[4], stats[5], stats[6]
            for c in ln.clicks:  # _object_roi_of: positive clicks join the mask
                if c.is_positive:
                    y, x = int(c.coords[0]), int(c.coords[1])
                    rmin, rmax, cmin, cmax = min(rmin, y), max(rmax, y), min(cmin, x), max(cmax, x)
            bb = _expand_bbox((rmin, rmax, cmin, cmax), self.expansion_ratio, self.min_crop_size)
            current = max(0, bb[0]), min(ln.H - 1, bb[1]), max(0, bb[2]), min(ln.W - 1, bb[3])
        if current is None:
            current = 0, ln.H - 1, 0, ln.W - 1
        if (ln.roi is None or not ZoomIn._check_roi(ln.roi, ln.clicks)
                or _bbox_iou(current, ln.roi) < self.recompute_thresh_iou):
            ln.roi = current
        return ln.roi

    def _points(self, ln: _Lane, roi) -> torch.Tensor:
        rmin, rmax, cmin, cmax = roi
        tc_ = [c.copy(coords=(self.S0 * (c.coords[0] - rmin) / (rmax - rmin + 1), self.S1 * (c.coords[1] - cmin) / (cmax - cmin + 1)))
               for c in ln.clicks]
        lists = [tc_]
        if self.with_flip:
            lists.append([c.copy(coords=(c.coords[0], self.S1 - c.coords[1] - 1)) for c in tc_])
        num_pos = [sum(c.is_positive for c in cl) for cl in lists]
        num_neg = [len(cl) - p for cl, p in zip(lists, num_pos)]
        n = max(1, max(num_pos + num_neg))
        if self.use_graph and n <= self.graph_clicks:
            n = self.graph_clicks
        total = []
        for cl in lists:
            pos = [c.coords_and_indx for c in cl if c.is_positive]
            neg = [c.coords_and_indx for c in cl if not c.is_positive]
            total.append(pos + (n - len(pos)) * [(-1, -1, -1)] + neg + (n - len(neg)) * [(-1, -1, -1)])
        return torch.tensor(np.asarray(total, dtype=np.float64), dtype=torch.float32).pin_memory().to(self.device, non_blocking=True)

    def _enqueue_prediction(self, ln: _Lane, roi, pred_thr: float):
        from . import _lib
        st = _lib.stream_ptr()
        nb = 2 if self.with_flip else 1
        net_in = torch.empty(nb, 4, self.S0, self.S1, dtype=torch.float32, device=self.device)
        _lib.call("isp_zoom_in_fwd", _lib.dptr(ln.image), _lib.dptr(ln.prob), ln.H, ln.W, roi[0], roi[1], roi[2], roi[3],
                  _lib.dptr(net_in), self.S0, self.S1, int(self.with_flip), st)
        points = self._points(ln, roi)
        if self.use_graph and points.shape[1] == 2 * self.graph_clicks:
            logits = self.net.forward_graphed(net_in, points)
        else:
            logits = self.net(net_in, points)["instances"]
        logits = logits.float()
        if tuple(logits.shape[2:]) != (self.S0, self.S1):  # base_predictor.py:93-98
            logits = F.interpolate(logits, mode="bilinear", align_corners=True, size=(self.S0, self.S1))
        logits = logits.contiguous()
        stats = ln.result_d[2:].view(torch.int32)
        _lib.call("isp_unzoom_probs", _lib.dptr(logits), self.S0, self.S1, int(self.with_flip), ln.H, ln.W, roi[0], roi[1],
                  roi[2], roi[3], _lib.dptr(ln.prob_next), _lib.dptr(ln.gt), float(pred_thr), float(self.prob_thresh),
                  _lib.dptr(ln.pred_mask), _lib.dptr(stats), st)
        ln.prob, ln.prob_next = ln.prob_next, (ln.prob if ln.prob is not None else torch.empty_like(ln.prob_next))
        ln.have_stats = True

    # -- driver ---------------------------------------------------------------------------------------------------
    def _advance(self, ln: _Lane, max_iou_thr, pred_thr, min_clicks, max_clicks):
        """Consume the lane's read-back (IoU of the last prediction, next click) and enqueue the next click's work."""
        ln.event.synchronize()
        res = ln.result_h
        stats = res[2:].view(torch.int32).tolist()
        if ln.have_stats:
            iou = stats[0] / stats[1] if stats[1] else float("nan")
            ln.ious.append(iou)
            n = len(ln.ious)
            if (iou >= max_iou_thr and n >= min_clicks) or n >= max_clicks:
                ln.done = True
                return
        keys = [int(v) & 0xFFFFFFFFFFFFFFFF for v in res[:2].tolist()]
        dmax = [np.frombuffer(np.uint32(k >> 32).tobytes(), dtype=np.float32)[0] for k in keys]
        is_pos = bool(dmax[0] > dmax[1])
        idx = 0xFFFFFFFF - (keys[0 if is_pos else 1] & 0xFFFFFFFF)
        y, x = divmod(int(idx), ln.W)
        ln.clicks.append(Click(is_positive=is_pos, coords=(y, x), indx=len(ln.clicks)))
        ln.clicked[y, x] = 1
        roi = self._roi_for(ln, stats)
        self._enqueue_prediction(ln, roi, pred_thr)
        self._enqueue_click(ln)

    def evaluate(self, samples: Sequence[Tuple[np.ndarray, np.ndarray]], max_iou_thr: float = 1.01, pred_thr: float = 0.49,
                 min_clicks: int = 1, max_clicks: int = 20, return_probs: bool = False):
        """-> [(clicks_list, ious float32 array, final probability map or None)] in sample order."""
        out = [None] * len(samples)
        pending = list(range(len(samples)))[::-1]
        active: List[_Lane] = []
        with torch.no_grad():
            while pending or active:
                while pending and len(active) < self.lanes:
                    i = pending.pop()
                    ln = _Lane(i, samples[i][0], samples[i][1], self.device)
                    self._enqueue_click(ln)
                    active.append(ln)
                for ln in list(active):
                    self._advance(ln, max_iou_thr, pred_thr, min_clicks, max_clicks)
                    if ln.done:
                        probs = ln.prob.cpu().numpy() if return_probs else None
                        out[ln.index] = (ln.clicks, np.array(ln.ious, dtype=np.float32), probs)
                        active.remove(ln)
        return out


# ------------------------------------------------------------------ synthetic datasets
def synthetic_dataset(kind: str = "grabcut", n: Optional[int] = None, seed: int = 0):
    """GrabCut-shaped: 50 images, sides ~ U(300..640), one object, with a -1 ignore band around the
    object border (datasets/grabcut.py:37-38); DAVIS-shaped: 345 images of 480 x 854.  Image = smooth
    random field, mask = union of random ellipses.  Returns [(uint8 HxWx3 image, int32 HxW mask)]."""
    import cv2
    rng = np.random.RandomState(seed)
    count = n if n is not None else (50 if kind == "grabcut" else 345)
    out = []
    for _ in range(count):
        if kind == "grabcut":
            H, W = int(rng.randint(300, 641)), int(rng.randint(300, 641))
        else:
            H, W = 480, 854
        low = rng.rand(H // 32 + 2, W // 32 + 2, 3).astype(np.float32)
        img = np.clip(cv2.resize(low, (W, H), interpolation=cv2.INTER_CUBIC), 0, 1)
        img = (img * 255).astype(np.uint8)
        mask = np.zeros((H, W), np.uint8)
        cy, cx = rng.uniform(0.3, 0.7) * H, rng.uniform(0.3, 0.7) * W
        for _k in range(int(rng.randint(1, 4))):
            ay, ax = rng.uniform(0.08, 0.25) * H, rng.uniform(0.08, 0.25) * W
            cv2.ellipse(mask, (int(cx + rng.uniform(-0.1, 0.1) * W), int(cy + rng.uniform(-0.1, 0.1) * H)),
                        (int(ax), int(ay)), float(rng.uniform(0, 180)), 0, 360, 1, -1)
        gt = mask.astype(np.int32)
        if kind == "grabcut":
            k = np.ones((5, 5), np.uint8)
            band = cv2.dilate(mask, k) - cv2.erode(mask, k)
            gt[band > 0] = -1
        out.append((img, gt))
    return out
