"""One full training step (click-embedding gradient through LoftUp + ViT) at a small batch, for an ncu launch list."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import isegprobe_b200 as isp
from isegprobe_b200.training import HeadTrainer
B = int(os.environ.get("BATCH", "2"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
pipe = isp.ISegPipeline("loftup", {"upsampler_path": None, "n_dim": 384}).to(dev).eval()
tr = HeadTrainer(pipe)
img, pts = bench.synth_inputs(B, 1)
gt = (img[:, 3:] > 0.5).float()
for _ in range(int(os.environ.get("STEPS", "2"))):
    loss = tr.step(img.to(dev), pts.to(dev), gt.to(dev))
torch.cuda.synchronize()
print("ok", float(loss))
