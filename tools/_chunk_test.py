import sys, torch, json
sys.path.insert(0, ".")
import bench, isegprobe_b200 as isp
dev = torch.device("cuda:0")
torch.manual_seed(0)
pipe = isp.ISegPipeline("loftup", {"upsampler_path": None, "n_dim": 384}, with_head=True).to(dev).eval()
img, pts = bench.synth_inputs(32, 1)
img, pts = img.to(dev), pts.to(dev)
for ci, ff in ((4, False), (4, True)):
    pipe.upsampler.chunk_images = ci
    pipe.upsampler.fuse_ffn = ff
    pipe.__dict__.pop("_graphs", None)
    for _ in range(3):
        pipe.features_graphed(img, pts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        pipe.features_graphed(img, pts)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"chunk_images": ci, "fuse_ffn": ff, "ms_per_step": e0.elapsed_time(e1) / 8, "img_s": 32 * 8 / e0.elapsed_time(e1) * 1e3,
                      "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
