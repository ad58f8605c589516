"""ORACLE (test infrastructure) -- a tiny deterministic interactive-segmentation 'model' on the CPU
(torch) used to pin the evaluation DRIVER (clicker, zoom-in / flip transforms, NoC loop) against the
reference's own driver: both sides call the same StubNet, so any difference in the click sequence or
the IoU curve comes from the driver code, not from a network."""
import torch


class StubNet(torch.nn.Module):
    with_prev_mask = True

    def forward(self, image, points):
        B, _, H, W = image.shape
        yy = torch.arange(H, dtype=torch.float32).view(1, H, 1)
        xx = torch.arange(W, dtype=torch.float32).view(1, 1, W)
        pts = points.to(torch.float32)
        P = pts.shape[1] // 2
        out = torch.full((B, 1, H, W), -1.0)
        for b in range(B):
            acc = torch.zeros(1, H, W)
            for k in range(2 * P):
                r, c, o = pts[b, k]
                if max(float(r), float(c)) < 0:
                    continue
                sign = 1.0 if k < P else -1.0
                sigma = 0.12 * min(H, W) if k < P else 0.06 * min(H, W)
                acc = acc + sign * 3.0 * torch.exp(-((yy - r) ** 2 + (xx - c) ** 2) / (2 * sigma * sigma))
            out[b] = out[b] + acc + 0.5 * (image[b, 3:4] - 0.5) + 0.2 * (image[b, 0:1] - 0.5)
        return {"instances": out}
