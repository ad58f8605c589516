#!/usr/bin/env python
"""Benchmark of the hot path (driver contract: see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W [--workload jbu|loftup] [--impl reference]

A step = one pass of  click maps -> click embedding -> DINOv2 ViT-S/14 -> upsampler (-> 448^2)
over one batch of synthetic 448x448 images with random-init weights.
  jbu    : BASELINE.json configs[1]  (FeatUp JBU stack / AdaptiveConv, batch 16 per GPU)  [default]
  loftup : BASELINE.json configs[2]  (LoftUp cross-attention to 448^2, bf16, batch 32 per GPU)
  eval   : BASELINE.json configs[3]  (20-click NoC evaluation loop, MaskCLIP ViT-B/16 + LoftUp(512) + head,
           eval_mode fixed448 with flip TTA, one synthetic GrabCut-shaped sample per rank and step; clicks/s)
  train  : BASELINE.json configs[4]  (IS training step as the reference runs it: click maps -> trainable click
           embedding -> frozen DINOv2-S/14 -> frozen LoftUp -> ConvSegHead, NFL loss, backward through the head AND
           through the frozen upsampler / backbone down to the click embedding, gradient all-reduce, Adam; GLOBAL
           batch 64 split over the ranks -- strong scaling, as the reference's batch_size // ngpus).
           `--train-head-only` freezes the click embedding (features under no_grad, head forward/backward only).
Multi-GPU: one process per GPU (torchrun), images sharded across ranks, no data-path
collective (weak scaling); time = max over ranks of the CUDA-event time of the K steps.
`--impl reference` times the CPU oracle port of the same path on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    "jbu": {"batch": 16, "upsampler": "jbu_featup", "params": {"backbone_type": "dinov2", "use_norm": True},
            "name": "DINOv2 ViT-S/14 + FeatUp JBU stack (AdaptiveConv, 32->448 px) forward, batch 16 at 448x448"},
    "loftup": {"batch": 32, "upsampler": "loftup", "params": {"upsampler_path": None, "n_dim": 384},
               "name": "DINOv2 ViT-S/14 + LoftUp cross-attention upsampler to 448x448, bf16, batch 32"},
    "eval": {"batch": 2, "upsampler": "loftup", "params": {"upsampler_path": None, "n_dim": 512},
             "name": "20-click NoC loop, MaskCLIP ViT-B/16 + LoftUp(512) + ConvSegHead, fixed448 + flip, synthetic "
                     "GrabCut-shaped samples sharded over the ranks"},
    "train": {"batch": 64, "upsampler": "loftup", "params": {"upsampler_path": None, "n_dim": 384},
              "name": "IS training step: frozen DINOv2-S/14 + LoftUp + ConvSegHead fwd/bwd, global batch 64 at 448x448"},
}
H = W = 448
P_CLICKS = 24


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def synth_inputs(batch, seed):
    from oracle import synth  # input generator only (seeded tensors), shared with the tests
    img = torch.cat([synth.image_batch(batch, H, W, seed=seed),
                     (synth.image_batch(batch, H, W, seed=seed + 77)[:, :1] > 0.7).float()], 1)
    pts = synth.click_points(batch, P_CLICKS, H, W, seed=seed + 5)
    return img, pts


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_port_step(wl, batch, seed=1):
    """One pass of the same path with the CPU oracle (torch fp32, all host threads)."""
    from oracle import distmaps as odm, head as ohead, jbu as ojbu, loftup as oloft, synth, vit as ovit
    torch.manual_seed(0)
    img, pts = synth_inputs(batch, seed)
    vsd = synth.vit_state_dict(384, 12, seed=0)
    psd = synth.patch_embed_state_dict(384, 14, 3, seed=0)
    if wl == "jbu":
        usd = ojbu.init_state_dict(384, seed=0)
    else:
        usd, cn = synth.loftup_state_dict(384, seed=0), synth.channelnorm_state_dict(384, seed=1)
    t0 = time.perf_counter()
    with torch.no_grad():
        nimg = ohead.normalize_image(img[:, :3])
        maps = torch.from_numpy(odm.distmaps(pts.numpy(), H, W, 5, 1.0, True))
        emb = ohead.patch_embed_forward(psd, torch.cat([img[:, 3:], maps], 1))
        lr = ovit.dinov2_forward(vsd, nimg, emb)
        if wl == "jbu":
            hr = ohead.bilinear_align_corners(ojbu.jbu_stack_forward(usd, lr, nimg), (H, W))
        else:
            hr = oloft.loftup_forward(usd, lr, nimg, cn["norm.weight"], cn["norm.bias"])
        chk = float(hr.mean())
    return time.perf_counter() - t0, chk


def run_reference(args):
    """Reference arm: the CPU port of the reference path (oracle/) on the host cores.  The
    reference itself cannot run on this box (no omegaconf/mmcv/timm, hub downloads); see DESIGN.md."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    wl = args.workload
    sample_b = 1
    for _ in range(min(args.warmup, 1)):
        cpu_port_step(wl, sample_b)
    ts = [cpu_port_step(wl, sample_b)[0] for _ in range(args.steps)]
    t = sum(ts) / len(ts)
    v = sample_b / t
    line = {"impl": "reference", "metric": "images/sec @448^2 DINOv2-S/14 + upsampler forward", "value": v,
            "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOADS[wl]["name"], "batch_per_step": sample_b},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{sample_b} image per step of the same workload, torch CPU fp32 oracle"},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_eval(args, rank, world, dev, dist):
    """Config 4: every rank runs the 20-click loop on its own samples (no collective in the timed region;
    the IoU curves are gathered once at the end).  value = clicks/s over all ranks; a click = one
    predictor call = one batch-2 (image + flip) forward of the whole model + the host-side click simulator."""
    import numpy as np
    import isegprobe_b200 as isp
    from isegprobe_b200 import _lib, evaluation as ev
    from isegprobe_b200 import dist as idist
    wl = WORKLOADS["eval"]
    torch.manual_seed(0)
    pipe = isp.ISegPipeline("loftup", wl["params"], backbone="maskclip",
                            head_params={"in_channels": 512, "num_layers": 2, "num_classes": 1}).to(dev).eval()
    pipe.embed_coords = isp.PatchEmbed((448, 448), (16, 16), 3, 768).to(dev).eval()
    n_w, n_t = max(args.warmup, 3), args.steps
    samples = ev.synthetic_dataset("grabcut", n=(n_w + n_t) * world, seed=0)
    mine = [samples[i] for i in idist.shard_indices(len(samples), world, rank)]
    pred = ev.FixedSizePredictor(pipe, dev, target_size=(448, 448), with_flip=True, use_graph=True)
    for img, gt in mine[:n_w]:
        ev.evaluate_sample(img, gt, pred, max_iou_thr=1.01, max_clicks=3)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    if sampler:
        sampler.start()
    e0.record()
    curves = [ev.evaluate_sample(img, gt, pred, max_iou_thr=1.01, max_clicks=20)[1] for img, gt in mine[n_w:n_w + n_t]]
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.stop_flag = True
        sampler.join()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0  # eager launches (none when every click replays the graph) ...
    fwd_graphs = pipe.__dict__.get("_fwd_graphs", {})
    if fwd_graphs:                         # ... plus the kernels inside each replayed graph
        launches += list(fwd_graphs.values())[-1][4] * 20 * n_t
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    rows = torch.tensor(np.stack(curves), device=dev)
    allr = idist.gather_sample_results(rows, n_t * world).cpu().numpy()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    noc, _, over = ev.compute_noc_metric(list(allr), [0.85, 0.90], max_clicks=20)
    clicks = 20 * n_t * world
    v = clicks / (ms / 1e3)
    line = {"metric": "clicks/sec, 20-click NoC loop @448^2 MaskCLIP ViT-B/16 + LoftUp + head (flip TTA)", "value": v,
            "unit": "clicks/s", "n_gpus": world, "steps": n_t, "warmup": n_w, "ms_per_step": ms / n_t,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["name"], "samples_per_rank": n_t, "clicks_per_sample": 20, "weights": "random init (seed 0)",
                       "l2": "inputs change every click; every forward streams > 2 GB of intermediates",
                       "parallelism": f"dp{world} (samples sharded round-robin, final IoU gather only)"},
            "e2e": {"value": v, "unit": "clicks/s", "h2d_bytes_per_step": int(20 * 2 * 4 * 448 * 448 * 4) * world,
                    "d2h_bytes_per_step": int(20 * 448 * 448 * 4) * world, "ms_per_step": ms / n_t,
                    "note": "the loop is end-to-end by construction: image / clicks go host->device and the "
                            "probability map comes back to the host clicker every click"},
            "gpu_launches": int(launches), "clocks": sampler.summary() if sampler else None, "roofline": None,
            "noc": {"NoC@85": float(noc[0]), "NoC@90": float(noc[1]), ">=20@85": int(over[0]), ">=20@90": int(over[1]),
                    "note": "random-init weights: the values only show the metric path runs"}}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="jbu")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-head-only", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import isegprobe_b200 as isp
    from isegprobe_b200 import _lib, upsamplers

    if args.workload == "eval":
        return run_eval(args, rank, world, dev, dist)
    wl = WORKLOADS[args.workload]
    train = args.workload == "train"
    B = wl["batch"] // world if train else wl["batch"]  # train: global batch split over ranks (trainer.py:66-68)
    torch.manual_seed(0)
    pipe = isp.ISegPipeline(wl["upsampler"], wl["params"], with_head=train).to(dev).eval()
    img_h, pts_h = synth_inputs(B, seed=1 + rank)
    img_h, pts_h = img_h.pin_memory(), pts_h.pin_memory()
    img_d, pts_d = img_h.to(dev), pts_h.to(dev)
    if train:
        from isegprobe_b200.training import HeadTrainer
        trainer = HeadTrainer(pipe, train_embedding=not args.train_head_only)
        gt_h = (img_h[:, 3:] > 0.5).float().pin_memory()  # synthetic instance masks (the prev-mask channel's blobs)
        gt_d = gt_h.to(dev)

        def step_device():
            return trainer.step(img_d, pts_d, gt_d)

        def step_e2e():
            return trainer.step(img_h.to(dev, non_blocking=True), pts_h.to(dev, non_blocking=True),
                                gt_h.to(dev, non_blocking=True)).cpu()  # loss read back each step
    else:
        # the step is replayed from a CUDA graph (ISegPipeline.features_graphed); the e2e variant copies the
        # pinned host inputs straight into the graph's static input buffers
        def step_device():
            return pipe.features_graphed(img_d, pts_d)

        # e2e: requests are pipelined two deep -- the pinned host inputs of step i+1 are copied (copy stream, second
        # set of static graph buffers) while the graph of step i runs, and the centre-pixel feature vectors of step i are read
        # back (async D2H into pinned memory) while step i+1 runs; the host waits for result i-1 before issuing i+1.
        h2d = torch.cuda.Stream()
        res_h = [torch.empty(B, 384, dtype=torch.float32).pin_memory() for _ in range(2)]
        res_ev = [torch.cuda.Event(), torch.cuda.Event()]
        e2e_i = [0]

        def step_e2e():
            i = e2e_i[0]
            slot = i & 1
            out = pipe.features_graphed(img_h, pts_h, slot=slot, h2d_stream=h2d)
            # the features stay on the device for the head; what is read back each step is each image's centre-pixel
            # feature vector (a full-tensor checksum would add a 4.9 GB reduction pass that is not part of the path)
            res_h[slot].copy_(out[:, :, H // 2, W // 2], non_blocking=True)
            res_ev[slot].record()
            if i > 0:
                res_ev[slot ^ 1].synchronize()  # result of the previous step is on the host now
            e2e_i[0] = i + 1
            return res_h[slot]

        def step_eager():
            with torch.no_grad():
                return pipe.features(img_d, pts_d)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    upsamplers.KERNEL_TIMERS.clear()
    upsamplers.KERNEL_TIMING = train  # events cannot be recorded inside a graph replay: the graphed workloads
    l0 = _lib.launch_count()          # time their dominant kernel in a few eager steps right after the timed region
    ms = timed(step_device, args.steps)
    launches = (_lib.launch_count() - l0) if train else pipe.graphed_launches() * args.steps
    if not train:
        upsamplers.KERNEL_TIMING = True
        for _ in range(2):
            step_eager()
    upsamplers.KERNEL_TIMING = False
    torch.cuda.synchronize()
    ktimes = {k: [a.elapsed_time(b) for a, b in v] for k, v in upsamplers.KERNEL_TIMERS.items()}
    if sampler:
        sampler.stop_flag = True
        sampler.join()
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk, pk_kind = peaks()
    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    C = 384
    roofline = None
    if args.workload == "jbu" and ktimes.get("adaptive_conv_512"):
        t = sum(ktimes["adaptive_conv_512"]) / len(ktimes["adaptive_conv_512"])
        alg = 4.0 * B * (C * 518 * 518 + 49 * 512 * 512 + C * 512 * 512)  # padded input + filters + output, fp32
        ach = alg / (t * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("adaptive_conv_512_bytes_per_image")
            traffic = traffic * B if traffic else None
        roofline = {"kernel": "adaptive_conv_v3_kernel (JBU stage 512)", "bound": "hbm", "achieved": ach,
                    "peak": pk["hbm_gbs"], "peak_kind": f"{pk_kind} hbm copy", "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "traffic": traffic, "ms_per_launch": t, "algorithmic_bytes_per_launch": alg}
    elif args.workload in ("loftup", "train") and ktimes.get("loftup_attention"):
        t = sum(ktimes["loftup_attention"]) / len(ktimes["loftup_attention"])
        imgs = ktimes.get("_loftup_attention_images", [pipe.upsampler.chunk_images])[0]
        flops = 2.0 * 2 * 4 * 200704 * 1024 * 101 * pipe.upsampler.chunk_images  # QK^T + PV, un-padded head dim
        ach = flops / (t * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("loftup_attention_bytes_per_image")
            traffic = traffic * pipe.upsampler.chunk_images if traffic else None
        roofline = {"kernel": "attention_kernel<2,7,112> (LoftUp cross-attention, per layer call)", "bound": "tensor",
                    "achieved": ach, "peak": pk["bf16_tflops_sustained"], "peak_kind": f"{pk_kind} cuBLAS bf16 sustained",
                    "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"], "traffic": traffic, "ms_per_launch": t,
                    "algorithmic_flops_per_launch": flops}
    line = {
        "metric": (("images/sec @448^2 IS training step (frozen DINOv2-S/14 + LoftUp, head fwd/bwd"
                    + (")" if args.train_head_only else " + click-embedding gradient through the frozen path)")) if train
                   else "images/sec @448^2 DINOv2-S/14 + upsampler forward"), "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if train else "weak", "vs_baseline": None,
        "dtype": "f32 (JBU SIMT kernels) + bf16 tcgen05 (ViT, 1x1 conv)" if args.workload == "jbu" else
                 ("bf16 (fp32 parameter gradients, loss and optimizer)" if train else "bf16"),
        "data": "synthetic",
        "config": {"workload": wl["name"], "batch_per_gpu": B, "global_batch": B * world, "image": "448x448",
                   "clicks_per_polarity": P_CLICKS, "weights": "random init (seed 0)",
                   "launch": "eager" if train else "CUDA graph replay (one graph per step)",
                   "l2": "no explicit flush: every step streams > 10 GB of intermediates (>> 126 MB L2)",
                   "parallelism": f"dp{world} (images sharded, no data-path collective)"},
        "e2e": {"value": e2e, "unit": "images/s",
                "h2d_bytes_per_step": int(img_h.numel() * 4 + pts_h.numel() * 4 + (gt_h.numel() * 4 if train else 0)) * world,
                "d2h_bytes_per_step": (4 if train else 4 * B * 384) * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": sampler.summary() if sampler else None,
        "roofline": roofline,
        "kernel_ms": {k: round(sum(v) / len(v), 4) for k, v in ktimes.items() if v},
    }
    if not args.no_cpu_baseline and world == 1 and not train:
        torch.set_num_threads(os.cpu_count())
        t, _ = cpu_port_step(args.workload, 1)
        line["cpu_baseline"] = {"value": 1.0 / t, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "1 image of the same workload (one step at batch 1), torch CPU fp32 oracle"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
