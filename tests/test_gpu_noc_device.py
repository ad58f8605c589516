"""Device-resident NoC evaluation loop: the three kernels against the host code they replace
(torch F.interpolate / flip / sigmoid, numpy IoU, cv2.distanceTransform clicker), and the whole loop against (a) the
reference driver's golden click sequences (tests/golden/noc_driver.npz, produced by core/inference/* around oracle/stubnet.py)
and (b) this package's host driver (FixedSizePredictor + Clicker, itself pinned to the reference) on the real pipeline."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import DEV

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _call(name, *a):
    from isegprobe_b200 import _lib
    _lib.call(name, *[(_lib.dptr(x) if torch.is_tensor(x) else x) for x in a], _lib.stream_ptr())


@pytest.mark.parametrize("Hs,Ws,roi,S", [(300, 417, (0, 299, 0, 416), (96, 128)), (333, 301, (17, 250, 40, 288), (112, 112)),
                                         (64, 80, (5, 5, 7, 60), (32, 48))])
def test_zoom_in_matches_torch(Hs, Ws, roi, S):
    g = torch.Generator().manual_seed(Hs + Ws)
    image = torch.rand(3, Hs, Ws, generator=g).to(DEV)
    prev = torch.rand(Hs, Ws, generator=g).to(DEV)
    out = torch.empty(2, 4, S[0], S[1], device=DEV)
    _call("isp_zoom_in_fwd", image, prev, Hs, Ws, roi[0], roi[1], roi[2], roi[3], out, S[0], S[1], 1)
    x = torch.cat([image, prev[None]], 0)[None, :, roi[0]:roi[1] + 1, roi[2]:roi[3] + 1]
    want = F.interpolate(x, size=S, mode="bilinear", align_corners=True)
    want = torch.cat([want, torch.flip(want, dims=[3])], 0)
    assert float((out - want).abs().max()) < 2e-6
    out1 = torch.empty(1, 4, S[0], S[1], device=DEV)
    _call("isp_zoom_in_fwd", image, None, Hs, Ws, roi[0], roi[1], roi[2], roi[3], out1, S[0], S[1], 0)
    assert torch.equal(out1[0, :3], out[0, :3]) and float(out1[0, 3].abs().max()) == 0.0


@pytest.mark.parametrize("Hs,Ws,roi,S,flip", [(300, 417, (0, 299, 0, 416), (96, 128), 1), (333, 301, (17, 250, 40, 288), (112, 112), 1),
                                              (120, 90, (10, 100, 5, 80), (64, 64), 0)])
def test_unzoom_matches_torch_and_numpy(Hs, Ws, roi, S, flip):
    from isegprobe_b200 import evaluation as ev
    g = torch.Generator().manual_seed(Hs * 3 + Ws)
    logits = (torch.randn(2 if flip else 1, 1, S[0], S[1], generator=g) * 2).to(DEV)
    gt = torch.zeros(Hs, Ws, dtype=torch.int32)
    gt[Hs // 4:Hs // 2, Ws // 4:Ws // 2] = 1
    gt[Hs // 2:Hs // 2 + 4] = -1
    prob = torch.empty(Hs, Ws, device=DEV)
    mask = torch.empty(Hs, Ws, dtype=torch.uint8, device=DEV)
    stats = torch.zeros(8, dtype=torch.int32, device=DEV)
    _call("isp_unzoom_probs", logits, S[0], S[1], flip, Hs, Ws, roi[0], roi[1], roi[2], roi[3], prob, gt.to(DEV), 0.49, 0.5,
          mask, stats)
    p = logits
    if flip:
        p = 0.5 * (p[:1] + torch.flip(p[1:], dims=[3]))
    p = torch.sigmoid(p)
    p = F.interpolate(p, size=(roi[1] - roi[0] + 1, roi[3] - roi[2] + 1), mode="bilinear", align_corners=True)
    want = torch.zeros(Hs, Ws, device=DEV)
    want[roi[0]:roi[1] + 1, roi[2]:roi[3] + 1] = p[0, 0]
    assert float((prob - want).abs().max()) < 2e-6
    wm = (want > 0.49).cpu().numpy()
    # pixels whose probability sits within float noise of a threshold may legitimately differ; none do on this input
    assert np.array_equal(mask.cpu().numpy().astype(bool), wm)
    st = stats.cpu().tolist()
    gtn = gt.numpy()
    keep, obj = gtn != -1, gtn == 1
    assert st[0] == int((wm & obj & keep).sum()) and st[1] == int(((wm | obj) & keep).sum())
    assert abs(st[0] / st[1] - ev.get_iou(gtn, wm)) < 1e-12
    box = (want > 0.5).cpu().numpy()
    rows, cols = np.where(box.any(1))[0], np.where(box.any(0))[0]
    assert st[2] == int(box.sum()) and st[3:7] == [rows[0], rows[-1], cols[0], cols[-1]]


@pytest.mark.parametrize("H,W,seed", [(97, 131, 0), (300, 417, 1), (640, 333, 2), (33, 40, 3)])
def test_next_click_matches_cv2_clicker(H, W, seed):
    """Same next click (polarity, coordinates) as Clicker.make_next_click on random blobs with an ignore band, for several
    consecutive clicks (already-clicked pixels excluded)."""
    from isegprobe_b200 import evaluation as ev
    rng = np.random.RandomState(seed)
    gt = np.zeros((H, W), np.int32)
    cv2.ellipse(gt, (W // 2, H // 2), (W // 4, H // 5), 20, 0, 360, 1, -1)
    band = cv2.dilate((gt == 1).astype(np.uint8), np.ones((5, 5), np.uint8)) - cv2.erode((gt == 1).astype(np.uint8), np.ones((5, 5), np.uint8))
    gt[band > 0] = -1
    clicker = ev.Clicker(gt_mask=gt)
    gt_d = torch.from_numpy(gt).to(DEV)
    clicked = torch.zeros(H, W, dtype=torch.uint8, device=DEV)
    work = torch.empty(2 * H * W, dtype=torch.int32, device=DEV)
    best = torch.zeros(2, dtype=torch.int64, device=DEV)
    pred = np.zeros((H, W), bool)
    for step in range(5):
        clicker.make_next_click(pred)
        c = clicker.clicks_list[-1]
        _call("isp_noc_next_click", gt_d, torch.from_numpy(pred.astype(np.uint8)).to(DEV), clicked, H, W, work, best)
        keys = [int(v) & 0xFFFFFFFFFFFFFFFF for v in best.cpu().tolist()]
        dmax = [np.frombuffer(np.uint32(k >> 32).tobytes(), dtype=np.float32)[0] for k in keys]
        is_pos = bool(dmax[0] > dmax[1])
        idx = 0xFFFFFFFF - (keys[0 if is_pos else 1] & 0xFFFFFFFF)
        assert (is_pos, divmod(idx, W)) == (c.is_positive, (int(c.coords[0]), int(c.coords[1]))), step
        clicked[int(c.coords[0]), int(c.coords[1])] = 1
        # next "prediction": a random blob around the object, so that both error types occur
        pred = np.zeros((H, W), np.uint8)
        cv2.ellipse(pred, (W // 2 + rng.randint(-W // 8, W // 8), H // 2 + rng.randint(-H // 8, H // 8)),
                    (max(2, W // 4 + rng.randint(-W // 10, W // 10)), max(2, H // 5 + rng.randint(-H // 10, H // 10))),
                    float(rng.uniform(0, 180)), 0, 360, 1, -1)
        pred = pred.astype(bool)


class _StubNetDev(torch.nn.Module):
    """oracle/stubnet.py on the input's device (same arithmetic, vectorised over the clicks in the same order)."""
    with_prev_mask = True

    def forward(self, image, points):
        B, _, H, W = image.shape
        dev = image.device
        yy = torch.arange(H, dtype=torch.float32, device=dev).view(1, H, 1)
        xx = torch.arange(W, dtype=torch.float32, device=dev).view(1, 1, W)
        pts = points.to(torch.float32).cpu()
        P = pts.shape[1] // 2
        out = torch.full((B, 1, H, W), -1.0, device=dev)
        for b in range(B):
            acc = torch.zeros(1, H, W, device=dev)
            for k in range(2 * P):
                r, c, _ = pts[b, k]
                if max(float(r), float(c)) < 0:
                    continue
                sign = 1.0 if k < P else -1.0
                sigma = 0.12 * min(H, W) if k < P else 0.06 * min(H, W)
                acc = acc + sign * 3.0 * torch.exp(-((yy - float(r)) ** 2 + (xx - float(c)) ** 2) / (2 * sigma * sigma))
            out[b] = out[b] + acc + 0.5 * (image[b, 3:4] - 0.5) + 0.2 * (image[b, 0:1] - 0.5)
        return {"instances": out}


def test_device_loop_matches_reference_driver_golden(golden):
    """The reference's own driver (core/inference/evaluation.py + BasePredictor + ZoomIn + Clicker around the stub network):
    identical click sequences, IoU curves and final probability maps from the device-resident loop."""
    from isegprobe_b200 import evaluation as ev
    g = golden("noc_driver")
    samples = ev.synthetic_dataset("grabcut", n=3, seed=5)
    evl = ev.DeviceNoCEvaluator(_StubNetDev(), DEV, target_size=(96, 128), with_flip=True, lanes=2)
    res = evl.evaluate(samples, max_iou_thr=0.95, pred_thr=0.49, max_clicks=8, return_probs=True)
    for si, (clicks, ious, probs) in enumerate(res):
        got = np.array([[int(c.is_positive), c.coords[0], c.coords[1]] for c in clicks], dtype=np.int64)
        assert np.array_equal(got, g[f"clicks_{si}"]), (si, got.tolist(), g[f"clicks_{si}"].tolist())
        assert np.allclose(ious, g[f"ious_{si}"], atol=1e-6)
        assert np.allclose(probs, g[f"probs_{si}"], atol=1e-5)


def test_device_loop_matches_host_driver_on_the_pipeline():
    """MaskCLIP ViT-B/16 + LoftUp(512) + head (BASELINE config 4's model) at a small crop: the device-resident loop and the
    host driver (FixedSizePredictor: torch transforms, numpy IoU, cv2 clicker) produce the same clicks and IoU curves."""
    import isegprobe_b200 as isp
    from isegprobe_b200 import evaluation as ev
    torch.manual_seed(0)
    S = 112
    pipe = isp.ISegPipeline("loftup", {"upsampler_path": None, "n_dim": 512}, backbone="maskclip",
                            head_params={"in_channels": 512, "num_layers": 2, "num_classes": 1}).to(DEV).eval()
    pipe.embed_coords = isp.PatchEmbed((S, S), (16, 16), 3, 768).to(DEV).eval()
    samples = ev.synthetic_dataset("grabcut", n=3, seed=9)
    host = ev.FixedSizePredictor(pipe, torch.device(DEV), target_size=(S, S), with_flip=True, use_graph=True)
    want = [ev.evaluate_sample(img, gt, host, max_iou_thr=1.01, max_clicks=5) for img, gt in samples]
    got = ev.DeviceNoCEvaluator(pipe, DEV, target_size=(S, S), with_flip=True, lanes=2).evaluate(
        samples, max_iou_thr=1.01, max_clicks=5, return_probs=True)
    for (wc, wi, wp), (gc, gi, gp) in zip(want, got):
        assert [(c.is_positive, int(c.coords[0]), int(c.coords[1])) for c in gc] == \
               [(c.is_positive, int(c.coords[0]), int(c.coords[1])) for c in wc]
        assert np.allclose(gi, wi, atol=1e-4), (gi, wi)
        assert float(np.abs(gp - wp).max()) < 1e-4


def test_sharded_dataset_evaluation_accepts_the_device_loop():
    from isegprobe_b200 import evaluation as ev
    samples = ev.synthetic_dataset("grabcut", n=3, seed=5)
    host = ev.FixedSizePredictor(_StubNetDev(), torch.device(DEV), target_size=(96, 128), with_flip=True)
    a = ev.evaluate_dataset_sharded(samples, host, max_iou_thr=0.95, max_clicks=8)
    b = ev.evaluate_dataset_sharded(samples, ev.DeviceNoCEvaluator(_StubNetDev(), DEV, target_size=(96, 128), lanes=3),
                                    max_iou_thr=0.95, max_clicks=8)
    assert len(a) == len(b) == 3
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.allclose(x, y, atol=1e-6)
