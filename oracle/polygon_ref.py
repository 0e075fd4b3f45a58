"""TEST INFRASTRUCTURE — CPU restatement of the reference's mask-target rasterisation.

`project_masks_on_boxes_ref` follows modeling/roi_heads/mask_head/loss.py:37-75; `fr_poly` follows
pycoco/maskApi.c:166-206 (`rleFrPoly`) line by line in Python ints / floats (C `double`), `rle_decode` :47-51,
union as rleMerge :53-75.  Pinned by tests/golden/mask_targets.npz (reference + its own pycocotools build).
"""
import math

import numpy as np


def _cint(v):
    """C (int) cast of a double: truncation toward zero."""
    return int(v)


def fr_poly(xy, h, w):
    """xy: flat [x0,y0,...] doubles.  Returns the RLE counts list (column-major)."""
    k = len(xy) // 2
    scale = 5.0
    x = [_cint(scale * xy[2 * j] + .5) for j in range(k)]
    y = [_cint(scale * xy[2 * j + 1] + .5) for j in range(k)]
    x.append(x[0])
    y.append(y[0])
    u, v = [], []
    for j in range(k):
        xs, xe, ys, ye = x[j], x[j + 1], y[j], y[j + 1]
        dx, dy = abs(xe - xs), abs(ys - ye)
        flip = (dx >= dy and xs > xe) or (dx < dy and ys > ye)
        if flip:
            xs, xe, ys, ye = xe, xs, ye, ys
        if dx >= dy:
            s = (ye - ys) / dx if dx != 0 else float("nan")
            for d in range(dx + 1):
                t = dx - d if flip else d
                u.append(t + xs)
                val = ys + s * t + .5
                v.append(_cint(val) if val == val else 0)
        else:
            s = (xe - xs) / dy
            for d in range(dy + 1):
                t = dy - d if flip else d
                v.append(t + ys)
                u.append(_cint(xs + s * t + .5))
    a = []
    for j in range(1, len(u)):
        if u[j] != u[j - 1]:
            xd = float(u[j] if u[j] < u[j - 1] else u[j] - 1)
            xd = (xd + .5) / scale - .5
            if math.floor(xd) != xd or xd < 0 or xd > w - 1:
                continue
            yd = float(v[j] if v[j] < v[j - 1] else v[j - 1])
            yd = (yd + .5) / scale - .5
            if yd < 0:
                yd = 0.0
            elif yd > h:
                yd = float(h)
            yd = math.ceil(yd)
            a.append(int(xd) * h + int(yd))
    a.append(h * w)
    a.sort()
    p, diffs = 0, []
    for t in a:
        diffs.append(t - p)
        p = t
    b = [diffs[0]]
    j = 1
    while j < len(diffs):
        if diffs[j] > 0:
            b.append(diffs[j])
            j += 1
        else:
            j += 1
            if j < len(diffs):
                b[-1] += diffs[j]
                j += 1
    return b


def rle_decode(counts, h, w):
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, val = 0, 0
    for c in counts:
        flat[pos:pos + c] = val
        pos += c
        val ^= 1
    return flat.reshape(w, h).T.copy()  # column-major -> [h, w]


def polygons_to_mask(polys, h, w):
    m = np.zeros((h, w), dtype=np.uint8)
    for p in polys:
        m |= rle_decode(fr_poly([float(t) for t in p], h, w), h, w)
    return m


def project_masks_on_boxes_ref(instances, boxes, M):
    """instances: list (per ROI) of lists of flat fp32 polygons in image coordinates; boxes [K,4] fp32 xyxy.
    crop (structures/segmentation_mask.py:94-108) and resize (:110-125) happen in fp32 like torch does."""
    boxes = np.asarray(boxes, dtype=np.float32)
    out = np.zeros((len(instances), M, M), dtype=np.float32)
    for i, (polys, box) in enumerate(zip(instances, boxes)):
        w = max(np.float32(box[2] - box[0]), np.float32(1))
        h = max(np.float32(box[3] - box[1]), np.float32(1))
        rw, rh = np.float32(float(M) / float(w)), np.float32(float(M) / float(h))
        scaled = []
        for p in polys:
            q = np.asarray(p, dtype=np.float32).copy()
            q[0::2] = (q[0::2] - box[0]) * rw
            q[1::2] = (q[1::2] - box[1]) * rh
            scaled.append(q)
        out[i] = polygons_to_mask(scaled, M, M)
    return out
