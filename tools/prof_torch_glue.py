"""Which torch (non-libisp) CUDA kernels run inside one training step, by total time (torch.profiler, batch 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from isegprobe_b200.pipeline import ISegPipeline
from isegprobe_b200.training import HeadTrainer

dev = torch.device("cuda:0")
cfg = bench.WORKLOADS["train"]
batch = int(os.environ.get("BATCH", "4"))
pipe = ISegPipeline(upsampler_type=cfg["upsampler"], upsampler_params=cfg["params"], with_head=True).to(dev).eval()
img, pts = bench.synth_inputs(batch, 1)
img, pts = img.to(dev), pts.to(dev)
tr = HeadTrainer(pipe)
gt = (torch.rand(batch, 1, bench.H, bench.W, device=dev) > 0.5).float()
for _ in range(2):
    tr.step(img, pts, gt)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(img, pts, gt)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[1] for r in rows if "isp::" not in r[0] and "Memcpy" not in r[0] or True)
rows.sort(key=lambda r: -r[1])
print("kernel / op (device ms, calls)")
for k, t, c in rows[:45]:
    print("%9.3f ms %5d  %s" % (t, c, k[:110]))
