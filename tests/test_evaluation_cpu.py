"""Evaluation driver (clicker, ZoomIn / flip transforms, predictor, NoC metric) vs the REFERENCE's own
driver: tests/golden/noc_driver.npz holds click sequences, IoU curves and final probability maps produced
by core/inference/evaluation.py + BasePredictor + ZoomIn + Clicker around oracle/stubnet.py
(oracle/make_golden.py::golden_noc_driver).  Runs on the CPU: the driver is device-agnostic host code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from isegprobe_b200 import evaluation as ev
from oracle.stubnet import StubNet

cv2 = pytest.importorskip("cv2")


def test_driver_matches_reference(golden):
    g = golden("noc_driver")
    samples = ev.synthetic_dataset("grabcut", n=3, seed=5)
    for si, (img, gt) in enumerate(samples):
        pred = ev.FixedSizePredictor(StubNet(), torch.device("cpu"), target_size=(96, 128), with_flip=True)
        clicks, ious, probs = ev.evaluate_sample(img, gt, pred, max_iou_thr=0.95, pred_thr=0.49, max_clicks=8)
        got = np.array([[int(c.is_positive), c.coords[0], c.coords[1]] for c in clicks], dtype=np.int64)
        assert np.array_equal(got, g[f"clicks_{si}"]), (si, got.tolist(), g[f"clicks_{si}"].tolist())
        assert np.allclose(ious, g[f"ious_{si}"], atol=1e-6)
        assert np.allclose(probs, g[f"probs_{si}"], atol=1e-5)
        # IoU by an independent expression (ignore band excluded)
        m = probs > 0.49
        keep = gt != -1
        assert abs(ious[-1] - ((m & (gt == 1) & keep).sum() / ((m | (gt == 1)) & keep).sum())) < 1e-6


def test_noc_metric():
    ious = [np.array([0.5, 0.86, 0.91], np.float32), np.array([0.2, 0.3], np.float32), np.array([0.95], np.float32)]
    noc, std, over = ev.compute_noc_metric(ious, [0.85, 0.90], max_clicks=20)
    assert noc == [np.mean([2, 20, 1]), np.mean([3, 20, 1])] and over == [1, 1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        samples = ev.synthetic_dataset("grabcut", n=3, seed=5)
        pred = ev.FixedSizePredictor(StubNet(), torch.device("cpu"), target_size=(96, 128), with_flip=True)
        all_ious = ev.evaluate_dataset_sharded(samples, pred, max_iou_thr=0.95, max_clicks=8)
        q.put((rank, [a.tolist() for a in all_ious]))
    finally:
        dist.destroy_process_group()


def test_sharded_evaluation_world2(golden):
    """Samples sharded round-robin over two gloo ranks; both ranks end with the full, ordered list of
    IoU curves, equal to the reference's single-process result."""
    g = golden("noc_driver")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for _, curves in res:
        assert len(curves) == 3
        for si in range(3):
            assert np.allclose(np.array(curves[si], np.float32), g[f"ious_{si}"], atol=1e-6)
