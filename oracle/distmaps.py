"""ORACLE (test infrastructure) -- click-map encoding, numpy restatement.

Follows the reference torch path `DistMaps.get_coord_features`
(/root/reference/core/model/ops.py:35-77) and the Cython BFS
(/root/reference/core/utils/cython/_get_dist_maps.pyx:18-64).

All arithmetic is float32 with separate multiply and add (the reference does
`coords.mul_(coords); coords[:,0] += coords[:,1]`, ops.py:61-63), so the result
is bit-comparable with the CUDA kernel, which uses __fmul_rn/__fadd_rn.
"""
import numpy as np

INVALID_D2 = np.float32(1e6)  # ops.py:66


def squared_distance_maps(points, rows, cols, norm_radius, spatial_scale=1.0,
                          use_disks=False):
    """points: [B, 2P, 3] (row, col, order) any real dtype.  Returns float32
    [B, 2, rows, cols]: min over valid clicks of the squared distance
    (ops.py:35-70); 1e6 where a polarity has no valid click."""
    pts = np.asarray(points)
    B, P2, _ = pts.shape
    P = P2 // 2
    # ops.py:55 `points * self.spatial_scale` -> float32 (python-float scale)
    rc = (pts[..., :2].astype(np.float32) * np.float32(spatial_scale)).astype(np.float32)
    # ops.py:40 invalid iff max(row, col) < 0 -- tested on the UNSCALED values
    invalid = pts[..., :2].astype(np.float32).max(axis=-1) < 0
    rr = np.arange(rows, dtype=np.float32)[:, None]
    cc = np.arange(cols, dtype=np.float32)[None, :]
    out = np.full((B, 2, rows, cols), INVALID_D2, dtype=np.float32)
    div = np.float32(norm_radius * spatial_scale)
    for b in range(B):
        for s in range(2):
            best = np.full((rows, cols), INVALID_D2, dtype=np.float32)
            for p in range(s * P, (s + 1) * P):
                if invalid[b, p]:
                    continue  # ops.py:66 fill 1e6 -> never below `best`
                dr = (rr - rc[b, p, 0]).astype(np.float32)
                dc = (cc - rc[b, p, 1]).astype(np.float32)
                if not use_disks:  # ops.py:59-60
                    dr = (dr / div).astype(np.float32)
                    dc = (dc / div).astype(np.float32)
                d2 = ((dr * dr).astype(np.float32) + (dc * dc).astype(np.float32)).astype(np.float32)
                best = np.minimum(best, d2)
            out[b, s] = best
    return out


def distmaps(points, rows, cols, norm_radius, spatial_scale=1.0, use_disks=False):
    """Full DistMaps.forward (ops.py:72-77): disks -> {0,1}; else tanh(2*sqrt(d2))."""
    d2 = squared_distance_maps(points, rows, cols, norm_radius, spatial_scale, use_disks)
    if use_disks:
        thr = np.float32((norm_radius * spatial_scale) ** 2)  # ops.py:73
        return (d2 <= thr).astype(np.float32)
    return np.tanh(np.float32(2.0) * np.sqrt(d2)).astype(np.float32)


def bfs_squared_distance_maps(points, rows, cols, norm_delimeter):
    """Pure-python restatement of the Cython multi-source BFS
    (_get_dist_maps.pyx:18-64) for ONE image.  points: float32 [2P, 3].
    Small cases only (python loops)."""
    pts = np.asarray(points, dtype=np.float32)
    n = pts.shape[0]
    dist = np.full((2, rows, cols), INVALID_D2, dtype=np.float32)
    q = []
    nd = np.float32(norm_delimeter)
    for i in range(n):
        x, y = int(round(float(pts[i, 0]))), int(round(float(pts[i, 1])))  # pyx:31
        if x >= 0:  # pyx:32 (only the row is tested)
            layer = 1 if i >= n / 2 else 0
            q.append((x, y, layer, x, y))
            dist[layer, x, y] = 0
    head = 0
    while head < len(q):
        r, c, layer, r0, c0 = q[head]
        head += 1
        for dx, dy in ((-1, 0), (0, -1), (0, 1), (1, 0)):
            x, y = r + dx, c + dy
            a = np.float32(np.float32(x - r0) / nd)
            b = np.float32(np.float32(y - c0) / nd)
            nd2 = np.float32(a * a + b * b)  # C float arithmetic, pyx:51
            if 0 <= x < rows and 0 <= y < cols and dist[layer, x, y] > nd2:
                q.append((x, y, layer, r0, c0))
                dist[layer, x, y] = nd2
    return dist
