#!/usr/bin/env python
"""Benchmark of the hot path (driver contract: see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W [--workload all|loftup|jbu|train|eval] [--impl reference]

A step = one pass of  click maps -> click embedding -> DINOv2 ViT-S/14 -> upsampler (-> 448^2)
over one batch of synthetic 448x448 images with random-init weights.

  loftup : BASELINE.json configs[2]  (LoftUp cross-attention to 448^2, bf16, batch 32 per GPU)   <- the headline line
  jbu    : BASELINE.json configs[1]  (FeatUp JBU stack / AdaptiveConv, batch 16 per GPU; JBU parity is UNPINNED)
  train  : BASELINE.json configs[4]  (IS training step as the reference runs it: click maps -> trainable click
           embedding -> frozen DINOv2-S/14 -> frozen LoftUp -> ConvSegHead, NFL loss, backward through the head AND
           through the frozen upsampler / backbone down to the click embedding, gradient all-reduce INSIDE the timed step
           (its own CUDA-event time is reported), Adam; GLOBAL batch 64 split over the ranks -- strong scaling, as the
           reference's batch_size // ngpus).  The whole model is in train() like the reference's trainer (the frozen LoftUp's
           BatchNorm uses batch statistics); --train-eval-mode-frozen keeps the frozen modules in eval().
  eval   : BASELINE.json configs[3]  (20-click NoC evaluation loop, MaskCLIP ViT-B/16 + LoftUp(512) + head,
           eval_mode fixed448 with flip TTA, synthetic GrabCut-shaped samples sharded over the ranks; clicks/s)
  all    : (default) the loftup line, with the other three as sub-records under "workloads" (fewer steps each).

`value`  : device-timed, inputs resident in HBM, CUDA-graph replay (forward workloads).
`e2e`    : the same step through the public API with HOST buffers: pinned host image + clicks -> device, the whole
           model INCLUDING the IS head, and the [B,1,448,448] logits copied back to the host every step.
Multi-GPU: one process per GPU (torchrun), images sharded across ranks, no data-path collective for the forward
workloads (weak scaling); time = max over ranks of the CUDA-event time of the K steps.
`--impl reference` times the reference's own model (`iSegProbeModel.forward`, unmodified, from the staged
baseline/_ref tree; kind "reference") -- or the CPU oracle port when no tree is staged (kind "port") -- on the host cores.
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
METRIC_FWD = "images/sec @448^2 DINOv2-S/14 + upsampler forward"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def workload_config(wl, world):
    """Definition of the workload -- identical in both arms (the b200 line and `--impl reference`); how an arm executes it
    (graph replay, sample size) is reported outside `config`."""
    w = WORKLOADS[wl]
    train = wl == "train"
    per_gpu = w["batch"] // world if train else w["batch"]
    return {"workload": w["name"], "batch_per_gpu": per_gpu, "global_batch": per_gpu * world, "image": "448x448",
            "clicks_per_polarity": P_CLICKS, "weights": "random init (seed 0)",
            "result": "features [B,384,448,448] (value); mask logits [B,1,448,448] after the IS head (e2e)",
            "l2": "no explicit flush: every step streams > 10 GB of intermediates (>> 126 MB L2)",
            "parallelism": f"dp{world} (images sharded over the ranks"
                           + (", one gradient all-reduce per step)" if train else ", no data-path collective)")}


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

    def finish(self):
        self.stop_flag = True
        self.join()
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(self.rows)}


class Ctx:
    """Process placement: one process per GPU."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.dist is None:
            return ms
        t = torch.tensor([ms], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def timed(self, fn, steps):
        """EXACTLY `steps` calls bracketed by barrier + synchronize; CUDA events on the launching stream; max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- CPU legs
def cpu_port_step(wl, batch, seed=1, with_head=False):
    """One pass of the same path with the CPU oracle (torch fp32, all host threads): (seconds to the features,
    seconds to the logits or None)."""
    from oracle import distmaps as odm, head as ohead, jbu as ojbu, loftup as oloft, synth, vit as ovit
    torch.manual_seed(0)
    img, pts = synth_inputs(batch, seed)
    vsd = synth.vit_state_dict(384, 12, seed=0)
    psd = synth.patch_embed_state_dict(384, 14, 3, seed=0)
    hsd = synth.convhead_state_dict(384, 2, 1, seed=0)
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
        float(hr.mean())
        t_feat = time.perf_counter() - t0
        t_all = None
        if with_head:
            float(ohead.convhead_forward(hsd, hr).mean())
            t_all = time.perf_counter() - t0
    return t_feat, t_all


class ReferenceCPU:
    """The reference's own assembled model (oracle/ref_model.py: unmodified `iSegProbeModel` built by `ModelBuilder` from
    the models/sbd/dinov2/patch-embed_loftup.py config) on the host cores, fp32, eval, no_grad -- BASELINE.md section 4.
    The time to the upsampler output is taken by a forward hook on `model.upsampler` (the LoftUp output already is 448^2,
    so no resize follows), the time to the logits at the return of `model(image, points)`."""

    def __init__(self, wl):
        from oracle import ref_model
        self.model = ref_model.build(WORKLOADS[wl]["upsampler"]).eval()
        self.t_up = None
        self.model.upsampler.register_forward_hook(lambda *a: setattr(self, "t_up", time.perf_counter()))

    def step(self, batch, seed=1):
        img, pts = synth_inputs(batch, seed)
        t0 = time.perf_counter()
        with torch.no_grad():
            out = self.model(img, pts)["instances"]
            float(out.mean())
        return self.t_up - t0, time.perf_counter() - t0


def reference_available(wl):
    from oracle import ref_shim
    return wl in ("loftup", "train") and ref_shim.available()


def cpu_baseline_record(wl):
    """Bounded sample for the b200 line: ONE image of the same workload on all host threads."""
    torch.set_num_threads(os.cpu_count())
    if reference_available(wl):
        ref = ReferenceCPU(wl)
        ref.step(1)  # warm-up (allocator, thread pool)
        t_feat, t_all = ref.step(1)
        kind = "reference"
        what = "the reference's own iSegProbeModel.forward (unmodified modules staged under baseline/_ref), torch CPU fp32"
    else:
        t_feat, t_all = cpu_port_step(wl, 1, with_head=True)
        kind, what = "port", "torch CPU fp32 oracle port (oracle/*.py)"
    return {"value": 1.0 / t_feat, "unit": "images/s", "cores": os.cpu_count(), "kind": kind,
            "sample": f"1 image of the same workload (one step at batch 1), {what}; value = to the upsampler output, "
                      "to_logits = through the IS head",
            "to_logits": 1.0 / t_all}


def run_reference(args):
    """Reference arm (rank 0 only): same metric / unit / config as the b200 line; every step is a bounded sample
    (one image) of the workload on all host threads."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    torch.set_num_threads(os.cpu_count())
    wl = "loftup" if args.workload == "all" else args.workload
    if wl in ("train", "eval"):
        print(json.dumps({"impl": "reference", "unavailable": f"the reference arm times the forward workloads (loftup, jbu); got {wl}"}))
        return
    if reference_available(wl):
        ref = ReferenceCPU(wl)
        step = lambda: ref.step(1)
        kind = "reference"
        what = "the reference's own iSegProbeModel.forward (unmodified modules, staged baseline/_ref tree), torch CPU fp32"
    else:
        step = lambda: cpu_port_step(wl, 1, with_head=True)
        kind = "port"
        what = "torch CPU fp32 oracle port (no reference tree staged" + ("; FeatUp's JBU is not in the reference tree)" if wl == "jbu" else ")")
    for _ in range(args.warmup):
        step()
    ts = [step() for _ in range(args.steps)]
    t_feat = sum(t[0] for t in ts) / len(ts)
    t_all = sum(t[1] for t in ts) / len(ts)
    v, v_all = 1.0 / t_feat, 1.0 / t_all
    line = {"impl": "reference", "metric": METRIC_FWD, "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_feat * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, args.gpus),
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": kind,
                             "sample": f"1 image per step of the same workload (per-image cost does not depend on the batch: "
                                       f"the only batch-coupled op is MinMaxScaler's min/max), {what}"},
            # value = to the upsampler output (what the b200 line's `value` times); e2e = through the IS head to the logits
            # (what the b200 line's `e2e` times) -- both measured in the same passes, so each ratio compares like with like
            "e2e": {"value": v_all, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "ms_per_step": t_all * 1e3}}
    print(json.dumps(line))


def reference_on_b200(dev, batch=2, steps=3):
    """Context number: the reference's own fp32 eager CUDA path (cuDNN / cuBLAS, TF32 off as in the reference)
    on the same GPU, through `iSegProbeModel.forward`.  Batch 2 = what the reference's evaluation feeds it (image + flip)."""
    from oracle import ref_model
    model = ref_model.build("loftup").to(dev).eval()
    img, pts = synth_inputs(batch, seed=1)
    img, pts = img.to(dev), pts.to(dev)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            model(img, pts)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model(img, pts)
            e1.record()
            torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    ms = e0.elapsed_time(e1) / steps
    del model
    torch.cuda.empty_cache()
    return {"value": batch / (ms / 1e3), "unit": "images/s", "batch": batch, "ms_per_step": ms,
            "what": "reference iSegProbeModel.forward (DINOv2-S/14 + LoftUp + ConvSegHead -> logits), fp32 eager torch on this "
                    "B200 (unmodified modules from baseline/_ref); compare with e2e / to_logits, not with value"}


# ----------------------------------------------------------------------------------------------- GPU workloads
def attention_standalone(dev, images, iters=10):
    """The LoftUp cross-attention kernel alone on chunk-shaped operands (for the roofline against the BURST peak)."""
    from isegprobe_b200 import _lib
    nh, HP, KP, T, HW = 4, 112, 128, 1024, H * W
    bf = torch.bfloat16
    hd = 101  # unit-variance scores (q, k ~ N(0, hd^-1/2) on the 101 real columns of every head), zero padding as in LoftUp
    Q = torch.zeros(images * HW, nh * HP, device=dev, dtype=bf)
    Q.view(-1, nh, HP)[:, :, :hd] = (torch.randn(images * HW, nh, hd, device=dev) * hd ** -0.25).to(bf)
    Kp = torch.zeros(images, nh, T, KP, device=dev, dtype=bf)
    Kp[..., :hd] = (torch.randn(images, nh, T, hd, device=dev) * hd ** -0.25).to(bf)
    Vt = torch.zeros(images, nh, HP, T, device=dev, dtype=bf)
    Vt[:, :, :hd] = torch.randn(images, nh, hd, T, device=dev).to(bf)
    Vt[:, :, hd] = 1.0  # the ones-row that accumulates the softmax denominator (as LoftUpUpsampler calls the kernel)
    O = torch.empty(images * HW, nh * HP, dtype=bf, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        _lib.call("isp_attention_bf16_tc_opt", Q.data_ptr(), nh * HP, HP, Kp.data_ptr(), Vt.data_ptr(), O.data_ptr(),
                  nh * HP, HP, images, HW, nh, T, 1, None, hd, 0, st)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    # "timed alone": the burst peak it is held against was measured on an idle GPU, so let the board leave the power-capped
    # state of the preceding steps, then time single launches (one event pair each) and take the median
    time.sleep(1.5)
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def run_forward(wl_name, args, ctx, steps, warmup, headline):
    import isegprobe_b200 as isp
    from isegprobe_b200 import upsamplers
    wl = WORKLOADS[wl_name]
    B, dev, world = wl["batch"], ctx.dev, ctx.world
    torch.manual_seed(0)
    pipe = isp.ISegPipeline(wl["upsampler"], wl["params"], with_head=True).to(dev).eval()
    img_h, pts_h = synth_inputs(B, seed=1 + ctx.rank)
    img_h, pts_h = img_h.pin_memory(), pts_h.pin_memory()
    img_d, pts_d = img_h.to(dev), pts_h.to(dev)

    def step_device():  # `value`: inputs resident, one CUDA graph replay (ISegPipeline.features_graphed)
        return pipe.features_graphed(img_d, pts_d)

    def step_logits():  # device-timed counterpart of the e2e leg
        return pipe.forward_graphed(img_d, pts_d, slot=2)

    # e2e: requests are pipelined two deep -- the pinned host inputs of step i+1 are copied (copy stream, second set of
    # static graph buffers) while the graph of step i runs, and the logits of step i are read back (async D2H into pinned
    # memory) while step i+1 runs; the host waits for result i-1 before issuing i+1.
    h2d = torch.cuda.Stream()
    res_h = [torch.empty(B, 1, H, W, dtype=torch.float32).pin_memory() for _ in range(2)]
    res_ev = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_i = [0]

    def step_e2e():
        i = e2e_i[0]
        slot = i & 1
        logits = pipe.forward_graphed(img_h, pts_h, slot=slot, h2d_stream=h2d)
        res_h[slot].copy_(logits, non_blocking=True)  # the step's result: every image's mask logits
        res_ev[slot].record()
        if i > 0:
            res_ev[slot ^ 1].synchronize()  # result of the previous step is on the host now
        e2e_i[0] = i + 1
        return res_h[slot]

    def step_eager():
        with torch.no_grad():
            return pipe.features(img_d, pts_d)

    for _ in range(max(warmup, 3)):
        step_device()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    if sampler:
        sampler.start()
    ms = ctx.timed(step_device, steps)
    launches = pipe.graphed_launches() * steps
    # events cannot be recorded inside a graph replay: the dominant kernel is timed in eager steps right after
    upsamplers.KERNEL_TIMERS.clear()
    upsamplers.KERNEL_TIMING = True
    for _ in range(2):
        step_eager()
    upsamplers.KERNEL_TIMING = False
    torch.cuda.synchronize()
    ktimes = {k: [a.elapsed_time(b) for a, b in v] for k, v in upsamplers.KERNEL_TIMERS.items()}
    clocks = sampler.finish() if sampler else None
    for _ in range(2):
        step_logits()
    ms_logits = ctx.timed(step_logits, steps)
    for _ in range(2):
        step_e2e()
    ms_e2e = ctx.timed(step_e2e, steps)
    if ctx.rank != 0:
        del pipe
        torch.cuda.empty_cache()
        return None
    pk, pk_kind = peaks()
    C = 384
    roofline = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    traffic_tab = json.load(open(tp)) if os.path.exists(tp) else {}
    if wl_name == "jbu":
        key = next((k for k in ("jbu_stage_512", "adaptive_conv_512") if ktimes.get(k)), None)
        if key:
            t = sum(ktimes[key]) / len(ktimes[key])
            alg = 4.0 * B * (C * 518 * 518 + 49 * 512 * 512 + C * 512 * 512)  # padded input + filters + output, fp32
            ach = alg / (t * 1e-3) / 1e9
            traffic = traffic_tab.get("adaptive_conv_512_bytes_per_image")
            roofline = {"kernel": "AdaptiveConv, JBU stage 512 (" + key + ")", "bound": "hbm", "achieved": ach,
                        "peak": pk["hbm_gbs"], "peak_kind": f"{pk_kind} hbm copy", "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                        "traffic": traffic * B if traffic else None,
                        "traffic_source": traffic_tab.get("adaptive_conv_512_source"),
                        "ms_per_launch": t, "algorithmic_bytes_per_launch": alg,
                        "note": "op-level bytes of the stand-alone AdaptiveConv call (SURVEY 8d): padded input + 7x7 filters + output"}
    elif ktimes.get("loftup_attention"):
        ci = pipe.upsampler.chunk_images
        t = sum(ktimes["loftup_attention"]) / len(ktimes["loftup_attention"])
        flops = 2.0 * 2 * 4 * 200704 * 1024 * 101 * ci  # QK^T + PV, un-padded head dim, per launch (one layer, one chunk)
        ach = flops / (t * 1e-3) / 1e12
        traffic = traffic_tab.get("loftup_attention_bytes_per_image")
        roofline = {"kernel": "attention_pair_kernel<2,7,112,LSUM> (LoftUp cross-attention, one layer of one chunk per launch)",
                    "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"],
                    "peak_kind": f"{pk_kind} cuBLAS bf16 sustained (kernel timed inside the step)", "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_tflops_sustained"], "traffic": traffic * ci if traffic else None,
                    "traffic_source": traffic_tab.get("loftup_attention_source"), "ms_per_launch": t,
                    "algorithmic_flops_per_launch": flops, "images_per_launch": ci}
        if headline:
            ts = attention_standalone(dev, ci)
            roofline["standalone"] = {"ms_per_launch": ts, "achieved": flops / (ts * 1e-3) / 1e12, "peak": pk["bf16_tflops"],
                                      "peak_kind": f"{pk_kind} cuBLAS bf16 burst (kernel timed alone)",
                                      "frac": flops / (ts * 1e-3) / 1e12 / pk["bf16_tflops"]}
    n_img = world * B * steps
    line = {
        "metric": METRIC_FWD, "value": n_img / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": steps,
        "warmup": max(warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32 (JBU SIMT kernels) + bf16 tcgen05 (ViT, 1x1 conv)" if wl_name == "jbu" else "bf16",
        "data": "synthetic", "config": workload_config(wl_name, world),
        "launch": "CUDA graph replay (one graph per step)",
        "to_logits": {"value": n_img / (ms_logits / 1e3), "unit": "images/s", "ms_per_step": ms_logits / steps,
                      "what": "device-timed like `value`, but through the IS head to the logits (the e2e leg's computation)"},
        "e2e": {"value": n_img / (ms_e2e / 1e3), "unit": "images/s",
                "h2d_bytes_per_step": int(img_h.numel() * 4 + pts_h.numel() * 4) * world,
                "d2h_bytes_per_step": int(B * H * W * 4) * world, "ms_per_step": ms_e2e / steps,
                "what": "ISegPipeline.forward_graphed: pinned host image + clicks -> device, click maps, ViT, upsampler, "
                        "ConvSegHead, [B,1,448,448] fp32 logits -> pinned host; requests pipelined two deep"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "kernel_ms": {k: round(sum(v) / len(v), 4) for k, v in ktimes.items() if v},
    }
    line["features_dtype"] = (str(getattr(pipe.upsampler, "out_dtype", torch.float32)).replace("torch.", "")
                              + " NHWC (`value` ends at the upsampled features in the format the pipeline's head consumes; "
                              "arithmetic dtype of the path: `dtype`)")
    if wl_name == "jbu":
        line["parity"] = "unpinned (FeatUp is not in the reference tree; oracle/jbu.py restates the published algorithm)"
    del pipe
    torch.cuda.empty_cache()
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_record(wl_name)
    return line


def run_train(args, ctx, steps, warmup):
    import isegprobe_b200 as isp
    from isegprobe_b200 import _lib, upsamplers
    from isegprobe_b200.training import HeadTrainer
    wl = WORKLOADS["train"]
    dev, world = ctx.dev, ctx.world
    B = wl["batch"] // world  # global batch split over the ranks (trainer.py:66-68)
    torch.manual_seed(0)
    pipe = isp.ISegPipeline(wl["upsampler"], wl["params"], with_head=True).to(dev).eval()
    img_h, pts_h = synth_inputs(B, seed=1 + ctx.rank)
    img_h, pts_h = img_h.pin_memory(), pts_h.pin_memory()
    gt_h = (img_h[:, 3:] > 0.5).float().pin_memory()  # synthetic instance masks (the prev-mask channel's blobs)
    img_d, pts_d, gt_d = img_h.to(dev), pts_h.to(dev), gt_h.to(dev)
    trainer = HeadTrainer(pipe, train_embedding=not args.train_head_only,
                          frozen_train_mode=not args.train_eval_mode_frozen)

    def step_device():
        return trainer.step(img_d, pts_d, gt_d)

    import random
    sim_rng, sim_rounds = random.Random(1234), []

    def step_e2e():
        # as the reference's batch_forward runs it (trainer.py:399-431): random.randint(0, 3) no-grad click-simulation rounds
        # (eval() forward, sigmoid, host-side cv2 next-click) in front of the graded forward; loss read back each step
        sim_rounds.append(sim_rng.randint(0, 3))
        return trainer.step(img_h.to(dev, non_blocking=True), pts_h.to(dev, non_blocking=True),
                            gt_h.to(dev, non_blocking=True), click_sim_rounds=sim_rounds[-1]).cpu()

    for _ in range(max(warmup, 3)):
        step_device()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    if sampler:
        sampler.start()
    upsamplers.KERNEL_TIMERS.clear()
    upsamplers.KERNEL_TIMING = True
    trainer.comm_events = []
    l0 = _lib.launch_count()
    ms = ctx.timed(step_device, steps)
    launches = _lib.launch_count() - l0
    upsamplers.KERNEL_TIMING = False
    comm = [a.elapsed_time(b) for a, b in trainer.comm_events]
    trainer.comm_events = None
    ktimes = {k: [a.elapsed_time(b) for a, b in v] for k, v in upsamplers.KERNEL_TIMERS.items()}
    clocks = sampler.finish() if sampler else None
    step_e2e()
    del sim_rounds[:]
    ms_e2e = ctx.timed(step_e2e, steps)
    comm_ms = ctx.max_over_ranks(sum(comm) / max(len(comm), 1))
    n_params = sum(p.numel() for p in trainer.params)
    del trainer, pipe
    torch.cuda.empty_cache()
    if ctx.rank != 0:
        return None
    n_img = world * B * steps
    return {
        "metric": ("images/sec @448^2 IS training step (frozen DINOv2-S/14 + LoftUp, head fwd/bwd"
                   + (")" if args.train_head_only else " + click-embedding gradient through the frozen path)")),
        "value": n_img / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16 (fp32 parameter gradients, loss and optimizer)", "data": "synthetic",
        "config": workload_config("train", world), "launch": "eager",
        "frozen_modules": ("eval() (running-statistics BatchNorm) -- NOT the reference trainer's semantics"
                           if args.train_eval_mode_frozen else
                           "net.train() on the whole model as the reference's trainer does (trainer.py:213-214): the frozen LoftUp's "
                           "BatchNorm runs on batch statistics of the rank's local batch and updates its running statistics"),
        "allreduce": {"ms_per_step": comm_ms, "bytes": 4 * n_params, "ranks": world, "inside_timed_step": True,
                      "what": "ONE NCCL all-reduce (sum, then / world) of the flat fp32 gradient arena: head + click embedding "
                              "(core/training/trainer.py:141-149 DDP semantics); CUDA events around the call, max over ranks"
                              + ("" if world > 1 else "; world size 1: no collective is issued")},
        "e2e": {"value": n_img / (ms_e2e / 1e3), "unit": "images/s",
                "h2d_bytes_per_step": int(img_h.numel() * 4 + pts_h.numel() * 4 + gt_h.numel() * 4) * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / steps,
                "click_simulation_rounds": list(sim_rounds),
                "what": "HeadTrainer.step from pinned host tensors incl. the reference's click-simulation rounds (random.randint(0, 3) "
                        "per step, seeded: eval() forward + sigmoid + host cv2 next-click each, trainer.py:399-431), loss read back; "
                        "`value` is the graded step alone (0 rounds)"},
        "gpu_launches": int(launches), "clocks": clocks,
        "kernel_ms": {k: round(sum(v) / len(v), 4) for k, v in ktimes.items() if v},
    }


def run_eval(args, ctx, steps, warmup):
    """Config 4: every rank runs the 20-click loop on its own samples (no collective in the timed region;
    the IoU curves are gathered once at the end).  value = clicks/s over all ranks; a click = one
    predictor call = one batch-2 (image + flip) forward of the whole model + the click simulator."""
    import numpy as np
    import isegprobe_b200 as isp
    from isegprobe_b200 import _lib, evaluation as ev
    from isegprobe_b200 import dist as idist
    wl = WORKLOADS["eval"]
    rank, world, dev, dist = ctx.rank, ctx.world, ctx.dev, ctx.dist
    torch.manual_seed(0)
    pipe = isp.ISegPipeline("loftup", wl["params"], backbone="maskclip",
                            head_params={"in_channels": 512, "num_layers": 2, "num_classes": 1}).to(dev).eval()
    pipe.embed_coords = isp.PatchEmbed((448, 448), (16, 16), 3, 768).to(dev).eval()
    n_w, n_t = max(warmup, 3), steps
    samples = ev.synthetic_dataset("grabcut", n=(n_w + n_t) * world, seed=0)
    mine = [samples[i] for i in idist.shard_indices(len(samples), world, rank)]
    # device-resident loop (ZoomIn / flip / un-zoom / IoU / click simulator on the GPU, two samples interleaved per rank);
    # --eval-host-driver runs the host driver instead (torch transforms, numpy IoU, cv2 clicker: the reference's structure)
    if args.eval_host_driver:
        pred = ev.FixedSizePredictor(pipe, dev, target_size=(448, 448), with_flip=True, use_graph=True)
        run = lambda ss, nclk: [ev.evaluate_sample(img, gt, pred, max_iou_thr=1.01, max_clicks=nclk)[1] for img, gt in ss]
    else:
        evl = ev.DeviceNoCEvaluator(pipe, dev, target_size=(448, 448), with_flip=True, lanes=2)
        run = lambda ss, nclk: [r[1] for r in evl.evaluate(ss, max_iou_thr=1.01, max_clicks=nclk)]
    run(mine[:n_w], 3)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0.record()
    curves = run(mine[n_w:n_w + n_t], 20)
    e1.record()
    ctx.barrier()
    clocks = sampler.finish() if sampler else None
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - l0 + pipe.graphed_launches() * 20 * n_t  # eager launches + the replayed graphs' kernels
    rows = torch.tensor(np.stack(curves), device=dev)
    allr = idist.gather_sample_results(rows, n_t * world).cpu().numpy()
    del run, pipe
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    noc, _, over = ev.compute_noc_metric(list(allr), [0.85, 0.90], max_clicks=20)
    clicks = 20 * n_t * world
    v = clicks / (ms / 1e3)
    return {"metric": "clicks/sec, 20-click NoC loop @448^2 MaskCLIP ViT-B/16 + LoftUp + head (flip TTA)", "value": v,
            "unit": "clicks/s", "n_gpus": world, "steps": n_t, "warmup": n_w, "ms_per_step": ms / n_t,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["name"], "samples_per_rank": n_t, "clicks_per_sample": 20, "weights": "random init (seed 0)",
                       "driver": ("host (torch transforms, numpy IoU, cv2 clicker)" if args.eval_host_driver else
                                  "device-resident (isp_zoom_in_fwd / isp_unzoom_probs / isp_noc_next_click, 2 samples interleaved)"),
                       "l2": "inputs change every click; every forward streams > 2 GB of intermediates",
                       "parallelism": f"dp{world} (samples sharded round-robin, final IoU gather only)"},
            "e2e": {"value": v, "unit": "clicks/s",
                    "h2d_bytes_per_step": int((20 * 2 * 4 * 448 * 448 * 4) if args.eval_host_driver else (640 * 640 * 16 + 20 * 2 * 48 * 12)) * world,
                    "d2h_bytes_per_step": int((20 * 448 * 448 * 4) if args.eval_host_driver else 20 * 48) * world, "ms_per_step": ms / n_t,
                    "note": "the loop is end-to-end by construction: the sample (image + ground truth) goes host->device once, "
                            "every click reads the next click / IoU counts / bounding box (48 bytes) back and sends the click list"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": None,
            "noc": {"NoC@85": float(noc[0]), "NoC@90": float(noc[1]), ">=20@85": int(over[0]), ">=20@90": int(over[1]),
                    "note": "random-init weights: the values only show the metric path runs"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["all"], default="all")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-context", action="store_true", help="skip the reference-on-B200 (fp32 eager CUDA) context leg")
    ap.add_argument("--train-head-only", action="store_true")
    ap.add_argument("--eval-host-driver", action="store_true", help="eval workload: the host driver instead of the device-resident loop")
    ap.add_argument("--train-eval-mode-frozen", action="store_true",
                    help="train workload: keep the frozen backbone / upsampler in eval() (running-statistics BatchNorm); the default "
                         "is net.train() on the whole model like the reference's trainer (batch-statistics BatchNorm in LoftUp)")
    ap.add_argument("--sub-steps", type=int, default=None, help="steps of the train / eval sub-records of --workload all")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    ctx = Ctx()
    try:
        if args.workload == "all":
            def done(name):  # progress on stderr, after a device sync: a kernel fault is then attributed to its workload
                torch.cuda.synchronize()
                if ctx.rank == 0:
                    print(f"[bench] workload {name} done", file=sys.stderr, flush=True)

            line = run_forward("loftup", args, ctx, args.steps, args.warmup, headline=True)
            done("loftup")
            sub = args.sub_steps or min(args.steps, 5)
            subs = {}

            def sub_record(name, fn):
                # one GPU: a failing sub-workload must not cost the headline line (several ranks: let it propagate, the
                # other ranks would otherwise wait at the next barrier)
                try:
                    subs[name] = fn()
                    done(name)
                except Exception as e:  # noqa: BLE001
                    if ctx.world > 1:
                        raise
                    subs[name] = {"error": repr(e)[:300]}
                    print(f"[bench] workload {name} FAILED: {e!r}", file=sys.stderr, flush=True)

            sub_record("jbu", lambda: run_forward("jbu", args, ctx, args.steps, args.warmup, headline=False))
            sub_record("train", lambda: run_train(args, ctx, sub, 3))
            sub_record("eval", lambda: run_eval(args, ctx, min(sub, 4), 3))
            if line is not None:
                line["workloads"] = subs
        elif args.workload == "train":
            line = run_train(args, ctx, args.steps, args.warmup)
        elif args.workload == "eval":
            line = run_eval(args, ctx, args.steps, args.warmup)
        else:
            line = run_forward(args.workload, args, ctx, args.steps, args.warmup, headline=True)
        if line is not None and ctx.world == 1 and not args.no_context and args.workload in ("all", "loftup"):
            from oracle import ref_shim
            if ref_shim.available():
                try:
                    line["reference_on_b200"] = reference_on_b200(ctx.dev)
                except Exception as e:  # context only: never lose the bench line over it
                    line["reference_on_b200"] = {"unavailable": repr(e)[:200]}
        if line is not None:
            print(json.dumps(line))
    finally:
        ctx.close()


if __name__ == "__main__":
    main()
