"""numpy restatement of the Gaussian target splat of src/lib/utils/image.py
(TEST INFRASTRUCTURE): gaussian_radius :95-115, gaussian2D :118-124,
draw_umich_gaussian :126-141, and the per-object loop of
datasets/sample/ctdet.py:99-127 that calls them."""
import math
import numpy as np


def gaussian_radius(det_size, min_overlap=0.7):
    """utils/image.py:95-115 (float64 arithmetic, returns min of the three roots)."""
    height, width = det_size
    a1 = 1
    b1 = height + width
    c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + math.sqrt(b1 ** 2 - 4 * a1 * c1)) / 2
    a2 = 4
    b2 = 2 * (height + width)
    c2 = (1 - min_overlap) * width * height
    r2 = (b2 + math.sqrt(b2 ** 2 - 4 * a2 * c2)) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (height + width)
    c3 = (min_overlap - 1) * width * height
    r3 = (b3 + math.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return min(r1, r2, r3)


def gaussian2d(diameter, sigma):
    """utils/image.py:118-124: float64 exp, values below eps*max zeroed."""
    m = (diameter - 1.) / 2.
    y, x = np.ogrid[-m:m + 1, -m:m + 1]
    h = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h


def draw_umich_gaussian(heatmap, center, radius, k=1):
    """utils/image.py:126-141: heatmap[...] = max(heatmap, gaussian) on the clipped window."""
    diameter = 2 * radius + 1
    g = gaussian2d(diameter, diameter / 6)
    x, y = int(center[0]), int(center[1])
    height, width = heatmap.shape[0:2]
    left, right = min(x, radius), min(width - x, radius + 1)
    top, bottom = min(y, radius), min(height - y, radius + 1)
    mh = heatmap[y - top:y + bottom, x - left:x + right]
    mg = g[radius - top:radius + bottom, radius - left:radius + right]
    if min(mg.shape) > 0 and min(mh.shape) > 0:
        np.maximum(mh, mg * k, out=mh)
    return heatmap


def splat_objects(obj_cls, obj_cx, obj_cy, obj_radius, obj_valid, C, H, W):
    """Dense target maps from object lists: the draw_gaussian calls of
    datasets/sample/ctdet.py:111-117 for every valid object.  [B,M] lists -> [B,C,H,W] f32."""
    B, M = np.asarray(obj_cls).shape
    hm = np.zeros((B, C, H, W), dtype=np.float32)
    for b in range(B):
        for m in range(M):
            if not obj_valid[b, m]:
                continue
            draw_umich_gaussian(hm[b, int(obj_cls[b, m])], (int(obj_cx[b, m]), int(obj_cy[b, m])),
                                int(obj_radius[b, m]))
    return hm
