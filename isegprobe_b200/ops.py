"""Click-map encoding and input normalisation -- mirrors core/model/ops.py of the
reference (DistMaps :8-80, BatchImageNormalize :96-105) with the same
constructor arguments and call signatures, computed by libisp_b200 kernels."""
import torch
from torch import nn

from . import _lib


def _points_f32(points: torch.Tensor) -> torch.Tensor:
    # reference: int64 points are promoted by `points * spatial_scale` (ops.py:55)
    return points.detach().to(torch.float32).contiguous()


class DistMaps(nn.Module):
    """Same interface as the reference DistMaps (core/model/ops.py:8-80).

    cpu_mode=True selected the Cython BFS in the reference (ops.py:21-34); here it
    selects the GPU kernel with that path's semantics (rounded click coordinates,
    row-only validity) -- there is no host path."""

    def __init__(self, norm_radius, spatial_scale=1.0, cpu_mode=False, use_disks=False):
        super().__init__()
        self.spatial_scale = spatial_scale
        self.norm_radius = norm_radius
        self.cpu_mode = cpu_mode
        self.use_disks = use_disks

    def get_coord_features(self, points, batchsize, rows, cols):
        pts = _points_f32(points)
        if pts.dim() != 3 or pts.shape[0] != batchsize or pts.shape[2] != 3 or pts.shape[1] % 2:
            raise ValueError(f"points must be [B, 2P, 3], got {tuple(points.shape)} for batch {batchsize}")
        P = pts.shape[1] // 2
        out = torch.empty(batchsize, 2, rows, cols, dtype=torch.float32, device=pts.device)
        if self.cpu_mode:
            nd = 1.0 if self.use_disks else self.spatial_scale * self.norm_radius  # ops.py:24-26
            _lib.call("isp_distmaps_rounded_sqdist_fwd", _lib.dptr(pts), _lib.dptr(out), batchsize, P, rows, cols,
                      float(nd), _lib.stream_ptr())
            if self.use_disks:  # ops.py:72-75 shared final step
                return (out <= (self.norm_radius * self.spatial_scale) ** 2).float()
            return out.sqrt_().mul_(2).tanh_()
        _lib.call("isp_distmaps_fwd", _lib.dptr(pts), _lib.dptr(out), batchsize, P, rows, cols,
                  float(self.norm_radius), float(self.spatial_scale), int(bool(self.use_disks)), _lib.stream_ptr())
        return out

    def forward(self, x, coords):
        return self.get_coord_features(coords, x.shape[0], x.shape[2], x.shape[3])


class BatchImageNormalize:
    """(x - mean) / std per channel on a new tensor (core/model/ops.py:96-105)."""

    def __init__(self, mean, std, dtype=torch.float):
        self.mean = torch.as_tensor(mean, dtype=dtype)[None, :, None, None]
        self.std = torch.as_tensor(std, dtype=dtype)[None, :, None, None]

    def __call__(self, tensor):
        return (tensor - self.mean.to(tensor.device)) / self.std.to(tensor.device)


def prepare_input(image, points, norm_radius=5, spatial_scale=1.0, use_disks=True,
                  mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """One-kernel fusion of iSegBaseModel.prepare_input + get_coord_features
    (core/model/iseg_base_model.py:91-110): image [B,3|4,H,W] in [0,1] ->
    (normalised RGB [B,3,H,W], coord features [B,2|3,H,W] = [prev_mask,] pos, neg)."""
    import ctypes

    image = image.detach().float().contiguous()
    B, Cin, H, W = image.shape
    pts = _points_f32(points)
    P = pts.shape[1] // 2
    norm = torch.empty(B, 3, H, W, dtype=torch.float32, device=image.device)
    coord = torch.empty(B, 3 if Cin == 4 else 2, H, W, dtype=torch.float32, device=image.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    _lib.call("isp_prepare_input_fwd", _lib.dptr(image), _lib.dptr(pts), _lib.dptr(norm), _lib.dptr(coord),
              B, Cin, P, H, W, ctypes.cast(m, ctypes.c_void_p), ctypes.cast(s, ctypes.c_void_p),
              float(norm_radius), float(spatial_scale), int(bool(use_disks)), _lib.stream_ptr())
    return norm, coord
