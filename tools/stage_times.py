"""Per-call CUDA-event times of one warm pass of a workload's feature path (eager launches, every C-ABI call
bracketed by events on the launching stream).  WORKLOAD=loftup|jbu, BATCH=n.  Prints calls in launch order,
aggregated by (entry point, integer arguments)."""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from isegprobe_b200 import _lib
from isegprobe_b200.pipeline import ISegPipeline

wl = os.environ.get("WORKLOAD", "loftup")
batch = int(os.environ.get("BATCH", "4"))
cfg = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
torch.manual_seed(0)
train = wl == "train"
pipe = ISegPipeline(upsampler_type=cfg["upsampler"], upsampler_params=cfg["params"], with_head=train).to(dev).eval()
img, pts = bench.synth_inputs(batch, 1)
img, pts = img.to(dev), pts.to(dev)
if train:
    from isegprobe_b200.training import HeadTrainer
    trainer = HeadTrainer(pipe)
    gt = (torch.rand(batch, 1, bench.H, bench.W, device=dev) > 0.5).float()
    run = lambda: trainer.step(img, pts, gt)
else:
    run = lambda: pipe.features(img, pts)
rec, orig = [], _lib.call
SHAPE_ARGS = {"isp_conv3x3_dgrad_bf16_tc": (6, 7, 8, 9, 11), "isp_conv3x3_wgrad_bf16_tc": (5, 6, 7, 8, 9),
              "isp_gemm_bf16_tc": (13, 14, 15, 9, 12), "isp_gemm_bf16_tc_ex": (13, 14, 15, 9, 12),
              "isp_conv3x3_bf16_tc": (6, 7, 8, 9, 11), "isp_conv3x3_bf16_tc_ex": (6, 7, 8, 9, 11)}

def timed(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(name, *a)
    e1.record()
    idx = SHAPE_ARGS.get(name)
    key = tuple(a[i] for i in idx) if idx else tuple(x for x in a[:-1] if isinstance(x, int) and 0 <= x < 10**7)[:8]
    rec.append((name, key, e0, e1))

import contextlib
with (contextlib.nullcontext() if train else torch.no_grad()):
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    _lib.call = timed
    for m in list(sys.modules.values()):
        if getattr(m, "__name__", "").startswith("isegprobe_b200") and hasattr(m, "_lib"):
            pass  # modules call _lib.call through the module attribute, patched above
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    run()
    t1.record()
    torch.cuda.synchronize()
_lib.call = orig
agg = collections.OrderedDict()
for name, key, e0, e1 in rec:
    a = agg.setdefault((name, key), [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"# {wl} batch {batch}: {len(rec)} calls, sum of call times {tot:.3f} ms, pass {t0.elapsed_time(t1):.3f} ms")
for (name, key), (n, ms) in agg.items():
    print(f"{name:32s} {str(key):44s} n={n:4d} {ms:9.3f} ms {100*ms/tot:5.1f}%")
