"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules
(imported from /root/reference through oracle/ref_shim.py) on seeded synthetic
inputs and the seeded weights of oracle/synth.py.

Run in the authoring container only:  python -m oracle.make_golden
The reference cannot travel to the GPU box, so the outputs are committed as
fixtures; weights/inputs are regenerated from seeds on both sides.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


def golden_distmaps():
    from core.model.ops import DistMaps
    cases = {}
    img = torch.zeros(3, 3, 40, 56)
    specs = [("int_p4", 4, False, torch.int64), ("f32_p4", 4, False, torch.float32),
             ("frac_p5", 5, True, torch.float32), ("frac_p24", 24, True, torch.float32)]
    for tag, p, frac, dt in specs:
        pts = synth.click_points(3, p, 40, 56, seed=10 + p, frac=frac)
        pts[2, p:] = -1  # one image with no negative click at all
        pts[1, 0] = torch.tensor([-1.0, 7.0, 0.0])  # one negative coord only -> still VALID (ops.py:40)
        pts = pts.to(dt)
        for disks in (True, False):
            out = DistMaps(norm_radius=5, spatial_scale=1.0, cpu_mode=False, use_disks=disks)(img, pts.clone())
            cases[f"{tag}_{'disk' if disks else 'tanh'}"] = out
            cases[f"{tag}_points"] = pts
    # Cython BFS path (demo-only, a3): integer clicks, both modes
    try:
        pts = synth.click_points(2, 3, 24, 32, seed=5).float()
        for disks in (True, False):
            out = DistMaps(norm_radius=5, spatial_scale=1.0, cpu_mode=True, use_disks=disks)(
                torch.zeros(2, 3, 24, 32), pts.clone())
            cases[f"bfs_{'disk' if disks else 'tanh'}"] = out
        cases["bfs_points"] = pts
    except Exception as e:  # pyximport build problem: record, do not fail the rest
        print("cython path unavailable:", e)
    save("distmaps", **cases)


def golden_loftup():
    from core.model.upsamplers.loftup.layers import ChannelNorm
    from core.model.upsamplers.loftup.loftup import LoftUp, UpsamplerwithChannelNorm
    sd = synth.loftup_state_dict(384, seed=0)
    cn = synth.channelnorm_state_dict(384, seed=1)
    up, chn = LoftUp(384, lr_pe_type="sine"), ChannelNorm(384)
    up.load_state_dict(sd, strict=True)
    chn.load_state_dict(cn, strict=True)
    model = UpsamplerwithChannelNorm(up, chn).eval()
    img = (synth.image_batch(2, 28, 42, seed=1) - 0.45) / 0.225
    lr = synth.lr_features(2, 384, 2, 3, seed=2)
    with torch.no_grad():
        out = model(lr, img)
        q = up.first_conv(up.fourier_feat(img))
        ff = up.fourier_feat(img)
    save("loftup_28x42", out=out, first_conv=q[:, :, ::3, ::3], fourier=ff[:, :, ::3, ::3])
    # the same module in train() mode -- how the reference's trainer runs the frozen upsampler (trainer.py:213-214):
    # BatchNorm batch statistics, running statistics updated.  Three images so that batch statistics differ from a pair.
    model.train()
    img3 = (synth.image_batch(3, 28, 42, seed=5) - 0.45) / 0.225
    lr3 = synth.lr_features(3, 384, 2, 3, seed=6)
    with torch.no_grad():
        out_t = model(lr3, img3)
    stats = {k.replace(".", "_"): v.detach().clone() for k, v in up.state_dict().items() if "running" in k or "num_batches" in k}
    save("loftup_train_28x42", out=out_t, **stats)


def golden_lift():
    from core.model.upsamplers.LiFT import LiFT
    sd = synth.lift_state_dict(384, seed=0)
    m = LiFT(384, 14)
    m.load_state_dict(sd, strict=True)
    m.eval()
    img = (synth.image_batch(2, 56, 84, seed=1) - 0.45) / 0.225
    lr = synth.lr_features(2, 384, 4, 6, seed=2)
    with torch.no_grad():
        out = m(img, lr)
    save("lift_56x84", out=out)
    # the same module in train() (core/training/trainer.py:213-214): its five BatchNorm layers use batch statistics and move
    # their running statistics; the gradient w.r.t. the LR features (what the click embedding is trained through) included
    m.train()
    img3 = (synth.image_batch(3, 56, 84, seed=5) - 0.45) / 0.225
    lr3 = synth.lr_features(3, 384, 4, 6, seed=6).requires_grad_(True)
    out_t = m(img3, lr3)
    gout = synth.lr_features(3, 384, 8, 12, seed=7)
    (out_t * gout).sum().backward()
    stats = {k.replace(".", "_"): v.detach().clone() for k, v in m.state_dict().items() if "running" in k or "num_batches" in k}
    save("lift_train_56x84", out=out_t.detach(), dsource=lr3.grad.detach(), **stats)


def golden_head():
    from core.model.heads.conv_heads import ConvSegHead
    from core.model.featurizers.utils.patch_embed import PatchEmbed
    sd = synth.convhead_state_dict(384, 2, 1, seed=0)
    m = ConvSegHead(384, 2, 1)
    m.load_state_dict(sd, strict=True)
    x = synth.lr_features(2, 384, 20, 28, seed=4)
    with torch.no_grad():
        out = m(x)
    pe_sd = synth.patch_embed_state_dict(384, 14, 3, seed=0)
    pe = PatchEmbed((28, 42), (14, 14), 3, 384)
    pe.load_state_dict(pe_sd, strict=True)
    maps = synth.image_batch(2, 28, 42, seed=6)
    with torch.no_grad():
        emb = pe(maps)
    save("head_20x28", out=out, patch_embed=emb)


def golden_vit():
    from core.model.featurizers.DINOv2 import vit_small
    sd = synth.vit_state_dict(384, depth=12, seed=0)
    m = vit_small(patch_size=14, img_size=518, init_values=1.0, block_chunks=0)
    missing = m.load_state_dict(sd, strict=True)
    m.eval()
    img = (synth.image_batch(2, 56, 84, seed=1) - 0.45) / 0.225
    emb = synth.lr_features(2, 24, 384, 1, seed=7).squeeze(-1) * 0.1
    with torch.no_grad():  # DINOv2Featurizer.forward 'before_backbone' (DINOv2.py:518-546)
        x = m.patch_embed(img)
        x = x + emb
        x = torch.cat((m.cls_token.expand(2, -1, -1), x), dim=1)
        x = x + m.interpolate_pos_encoding(x, 56, 84)
        for blk in m.blocks:
            x = blk(x)
        f = m.norm(x)[:, 1:]
        f = f.reshape(-1, 4, 6, 384).permute(0, 3, 1, 2)
    save("vit_56x84", out=f)


def golden_dino_vit():
    """DINO.py imports timm and core.utils.log at module level: timm is stubbed (only used by the hub-loading
    constructor, which is bypassed) and DINOFeaturizer.forward is called unbound on a stand-in carrying the
    attributes it reads (model, patch_size, feat_type, feats_injection_mode)."""
    import types
    sys.modules.setdefault("timm", types.ModuleType("timm"))
    log = types.ModuleType("core.utils.log")
    log.logger = __import__("logging").getLogger("ref")
    sys.modules.setdefault("core.utils.log", log)
    from core.model.featurizers import DINO as ref
    m = ref.vit_small(patch_size=16, num_classes=0)
    sd = synth.dino_vit_state_dict(seed=0)
    m.load_state_dict(sd, strict=True)
    m.eval()
    img = (synth.image_batch(2, 64, 96, seed=1) - 0.45) / 0.225
    emb = synth.lr_features(2, 24, 384, 1, seed=7).squeeze(-1) * 0.1  # [2, 24, 384]
    outs = {}
    with torch.no_grad():
        for ft in ("key", "token"):
            me = types.SimpleNamespace(model=m, patch_size=16, feat_type=ft, feats_injection_mode="before_backbone")
            outs[ft] = ref.DINOFeaturizer.forward(me, img.clone(), emb.clone())
    save("dino_vit_64x96", key=outs["key"], token=outs["token"])


def golden_simple_vit():
    from core.model.featurizers.simple_ViT import SimpleViTFeaturizer
    m = SimpleViTFeaturizer(image_size=[56, 84], patch_size=(14, 14), dim=384, depth=2, heads=8, mlp_dim=2048, channels=3,
                            dim_head=64)
    sd = synth.simple_vit_state_dict(depth=2, seed=0)
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = synth.image_batch(2, 56, 84, seed=4)
    with torch.no_grad():
        out = m(x.clone())
    save("simple_vit_56x84", out=out)


def golden_maskclip():
    """maskclip/model.py imported BY PATH (the package __init__ pulls the CLIP tokenizer, which needs ftfy)."""
    import importlib.util
    base = os.path.join(ref_shim.REFERENCE_ROOT, "core", "model", "featurizers", "maskclip")
    pkg = __import__("types").ModuleType("maskclip_ref")
    pkg.__path__ = [base]
    sys.modules["maskclip_ref"] = pkg
    for name in ("interpolate", "model"):
        spec = importlib.util.spec_from_file_location(f"maskclip_ref.{name}", os.path.join(base, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"maskclip_ref.{name}"] = mod
        spec.loader.exec_module(mod)
    VT = sys.modules["maskclip_ref.model"].VisionTransformer
    m = VT(224, 16, 768, 12, 12, 512)
    sd = synth.maskclip_state_dict(seed=0)
    m.load_state_dict(sd, strict=True)
    m.eval()
    img = (synth.image_batch(2, 64, 96, seed=1) - 0.45) / 0.225
    emb = synth.lr_features(2, 24, 768, 1, seed=7).squeeze(-1) * 0.1
    with torch.no_grad():
        plain = m(img, patch_output=True)  # get_patch_encodings (MaskCLIP.py:70)
        x = m.conv1(img)                   # 'before_backbone' branch (MaskCLIP.py:52-66)
        x = x.reshape(2, 768, -1).permute(0, 2, 1) + emb
        inj = m.forward_without_patch_embed(x, (64, 96), patch_output=True)
        sq = m((synth.image_batch(1, 64, 64, seed=2) - 0.45) / 0.225, patch_output=True)
    save("maskclip_64x96", plain=plain.reshape(2, 4, 6, 512).permute(0, 3, 1, 2),
         injected=inj.reshape(2, 4, 6, 512).permute(0, 3, 1, 2), square=sq.reshape(1, 4, 4, 512).permute(0, 3, 1, 2))


def golden_nfl():
    """NormalizedFocalLossSigmoid(alpha=0.5, gamma=2) (core/training/losses.py:42-109, models/defaults.py:24)
    value and gradient.  losses.py imports core.utils.misc, which imports the whole model package; only
    `get_dims_with_exclusion` (misc.py:28-33) is used, so that one helper is stubbed."""
    import types
    tr = types.ModuleType("core.training")
    tr.__path__ = [os.path.join(ref_shim.REFERENCE_ROOT, "core", "training")]
    sys.modules["core.training"] = tr
    m = types.ModuleType("core.utils.misc")
    m.get_dims_with_exclusion = lambda dim, exclude=None: [d for d in range(dim) if d != exclude]
    sys.modules["core.utils.misc"] = m
    sys.modules["core.utils"].misc = m
    from core.training.losses import NormalizedFocalLossSigmoid
    g = torch.Generator().manual_seed(11)
    pred = torch.randn(3, 1, 24, 40, generator=g) * 3
    label = (torch.rand(3, 1, 24, 40, generator=g) > 0.6).float()
    label[1, :, :4] = -1
    pr = pred.clone().requires_grad_(True)
    out = NormalizedFocalLossSigmoid(alpha=0.5, gamma=2)(pr, label)
    out.mean().backward()
    save("nfl_loss", pred=pred, label=label, out=out.detach(), grad=pr.grad)


def golden_next_points():
    """`get_next_points` (core/training/trainer.py:577-618), the click simulation of the training loop.  trainer.py cannot be
    imported here (core.data needs albumentations), so the UNMODIFIED source of that one function is cut out of the reference
    file with `ast` and executed with the names it uses (cv2, np, torch)."""
    import ast
    import cv2
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, "core", "training", "trainer.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_next_points")
    ns = {"cv2": cv2, "np": np, "torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "trainer.py:get_next_points", "exec"), ns)
    g = torch.Generator().manual_seed(11)
    B, H, W, P = 3, 60, 84, 6
    gt = torch.zeros(B, 1, H, W)
    gt[0, 0, 10:40, 20:60] = 1
    gt[1, 0, 5:55, 5:30] = 1
    gt[2, 0, 30:50, 40:80] = 1
    pred = torch.rand(B, 1, H, W, generator=g)
    pred[0, 0, 15:35, 25:50] += 0.6   # partly right: false negatives around it
    pred[1, 0, 20:50, 40:70] += 0.7   # a false-positive blob
    points = torch.full((B, 2 * P, 3), -1.0)
    np.random.seed(123)
    out1 = ns["get_next_points"](pred, gt, points, 1)
    out2 = ns["get_next_points"](pred.flip(3), gt, out1, 2)
    save("next_points", pred=pred, gt=gt, out1=out1, out2=out2)


def golden_noc_driver():
    """Click sequences and IoU curves of the reference's evaluate_sample (core/inference/evaluation.py:43-88)
    with BasePredictor + ZoomIn(skip_clicks=-1) + flip (the eval_mode='fixedNNN' stack) around oracle/stubnet.py.
    Real reference code: clicker.py, transforms/*, predictors/base_predictor.py, evaluation.py.  Stubs:
    core.utils.misc's model import, core.inference.utils.get_iou (same 3-line formula; checked again in the
    test with an independent expression), dataset / logging imports."""
    import importlib.util
    import types
    from isegprobe_b200.evaluation import synthetic_dataset
    from oracle.stubnet import StubNet
    root = ref_shim.REFERENCE_ROOT
    sys.modules["core.model"].iSegBaseModel = object
    log = types.ModuleType("core.utils.log")
    log.logger = __import__("logging").getLogger("ref")
    sys.modules["core.utils.log"] = log
    sys.modules.pop("core.utils.misc", None)
    for name, rel in (("core.inference", "core/inference"), ("core.inference.predictors", "core/inference/predictors"),
                      ("core.data", "core/data")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(root, rel)]
        sys.modules[name] = m
    bd = types.ModuleType("core.data.base_dataset")
    bd.iSegBaseDataset = object
    sys.modules["core.data.base_dataset"] = bd
    iu = types.ModuleType("core.inference.utils")

    def get_iou(gt_mask, pred_mask, ignore_label=-1):
        keep, obj = gt_mask != ignore_label, gt_mask == 1
        return (np.logical_and(np.logical_and(pred_mask, obj), keep).sum()
                / np.logical_and(np.logical_or(pred_mask, obj), keep).sum())
    iu.get_iou = get_iou
    sys.modules["core.inference.utils"] = iu
    sys.modules["core.inference"].utils = iu
    from core.inference.predictors.base_predictor import BasePredictor  # by path: the package __init__ pulls the BRS code
    sys.modules["core.inference.predictors"].BasePredictor = BasePredictor
    from core.inference.evaluation import evaluate_sample
    from core.inference.transforms import ZoomIn
    arrays = {}
    samples = synthetic_dataset("grabcut", n=3, seed=5)
    for si, (img, gt) in enumerate(samples):
        pred = BasePredictor(StubNet(), torch.device("cpu"), zoom_in=ZoomIn(skip_clicks=-1, target_size=(96, 128)),
                             with_flip=True)
        clicks, ious, probs = evaluate_sample(img, gt, pred, max_iou_thr=0.95, pred_thr=0.49, max_clicks=8)
        arrays[f"clicks_{si}"] = np.array([[int(c.is_positive), c.coords[0], c.coords[1]] for c in clicks], dtype=np.int64)
        arrays[f"ious_{si}"] = ious
        arrays[f"probs_{si}"] = probs.astype(np.float32)
    save("noc_driver", **arrays)


def golden_jbu_shape():
    """The only anchor the reference holds for JBU is the shape contract of
    JBUFeatUp.py:36-45; record it (parity unpinned, see oracle/jbu.py)."""
    save("jbu_contract", source_shape=np.array([1, 384, 14, 14]), guidance_shape=np.array([1, 3, 224, 224]),
         out_shape=np.array([1, 384, 224, 224]))


if __name__ == "__main__":
    assert ref_shim.available(), "reference tree not mounted"
    ref_shim.install()
    torch.set_num_threads(os.cpu_count())
    golden_distmaps()
    golden_loftup()
    golden_lift()
    golden_head()
    golden_vit()
    golden_dino_vit()
    golden_simple_vit()
    golden_maskclip()
    golden_nfl()
    golden_next_points()
    golden_noc_driver()
    golden_jbu_shape()
