"""Where one NoC click goes (eval workload): host clicker, host transforms + H2D, GPU graph replay, D2H."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import isegprobe_b200 as isp
from isegprobe_b200 import evaluation as ev

dev = torch.device("cuda:0")
wl = bench.WORKLOADS["eval"]
pipe = isp.ISegPipeline(wl["upsampler"], wl["params"], backbone="maskclip",
                        head_params={"in_channels": 512, "num_layers": 2, "num_classes": 1}).to(dev).eval()
pipe.embed_coords = isp.PatchEmbed((448, 448), (16, 16), 3, 768).to(dev).eval()
pred = ev.FixedSizePredictor(pipe, dev, use_graph=True)
samples = ev.synthetic_dataset("grabcut", n=3, seed=0)
T = {"clicker": 0.0, "predict": 0.0, "iou": 0.0, "n": 0}
orig_click, orig_pred = ev.Clicker.make_next_click, ev.FixedSizePredictor.get_prediction


def t_click(self, m):
    t = time.perf_counter(); orig_click(self, m); T["clicker"] += time.perf_counter() - t


def t_pred(self, c):
    t = time.perf_counter(); r = orig_pred(self, c); T["predict"] += time.perf_counter() - t; T["n"] += 1
    return r


ev.evaluate_sample(*samples[0], pred, max_iou_thr=1.01, max_clicks=20)
ev.Clicker.make_next_click, ev.FixedSizePredictor.get_prediction = t_click, t_pred
t0 = time.perf_counter()
for img, gt in samples[1:]:
    ev.evaluate_sample(img, gt, pred, max_iou_thr=1.01, max_clicks=20)
tot = time.perf_counter() - t0
n = T["n"]
print("clicks %d  total %.2f ms/click  clicker %.2f  get_prediction %.2f  other %.2f" % (
    n, tot / n * 1e3, T["clicker"] / n * 1e3, T["predict"] / n * 1e3, (tot - T["clicker"] - T["predict"]) / n * 1e3))
# GPU time of one graph replay
img, gt = samples[1]
pred.set_input_image(img)
clk = ev.Clicker(gt_mask=gt)
clk.make_next_click(np.zeros_like(gt))
pred.get_prediction(clk)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
key = list(pipe._graphs.keys())[0] if hasattr(pipe, "_graphs") and pipe._graphs else None
print("graphs:", list(getattr(pipe, "_graphs", {}).keys())[:2], [k for k in pipe.__dict__ if "graph" in k])
