"""Synthetic 2-class cytology crops in the exact form the reference's loaders hand to the trainer
(SURVEY §8d): a 1000x1000 crop goes through the reference transform contract — Resize(min 800 / max 1333)
(data/transforms/transforms.py:28-59 -> 800x800), ToTensor, BGR*255, minus PIXEL_MEAN (config/defaults.py:48-52) —
so what reaches `MTtrainer.forward_source` is a fp32 [B,3,800,800] tensor plus per-image BoxLists with
`labels` (1 = cytoplasm, 2 = nucleus) and `masks` (SegmentationMask of 16-gon polygons).  The unlabeled batch is
AUG_K + AUG_S colour-jittered views of one other crop batch (brightness / contrast, data/transforms/build.py:21-33).
No dataset or network access is needed; everything is seeded."""
import math

import torch

from ..structures.bounding_box import BoxList
from ..structures.segmentation_mask import SegmentationMask

PIXEL_MEAN = (102.9801, 115.9465, 122.7717)


def _polygon(g, cx, cy, r, n=16):
    ang = torch.arange(n, dtype=torch.float32) / n * (2 * math.pi) + float(torch.rand(1, generator=g))
    rad = r * (0.8 + 0.4 * torch.rand(n, generator=g))
    return torch.stack([cx + rad * torch.cos(ang), cy + rad * torch.sin(ang)], 1)


def make_targets(g, size, n_cyto=20, n_nuc=20, crop=1000):
    """Objects are placed in the `crop` frame and scaled to `size` like Resize does to BoxList / polygons."""
    s = size / float(crop)
    boxes, labels, polys = [], [], []
    for label, count, (r0, r1) in ((1, n_cyto, (30., 80.)), (2, n_nuc, (8., 20.))):
        for _ in range(count):
            r = r0 + (r1 - r0) * float(torch.rand(1, generator=g))
            cx = r + (crop - 2 * r) * float(torch.rand(1, generator=g))
            cy = r + (crop - 2 * r) * float(torch.rand(1, generator=g))
            p = (_polygon(g, cx, cy, r) * s).clamp(0, size - 1)
            lo, hi = p.min(0)[0], p.max(0)[0]
            boxes.append(torch.cat([lo, hi]))
            labels.append(label)
            polys.append([p.reshape(-1)])
    t = BoxList(torch.stack(boxes), (size, size), mode="xyxy")
    t.add_field("labels", torch.tensor(labels, dtype=torch.int64))
    t.add_field("masks", SegmentationMask(polys, (size, size)))
    return t


def make_image(g, target, size):
    """uint8-range noise + brighter blobs where the objects are, then BGR255 - PIXEL_MEAN."""
    img = torch.rand(3, size, size, generator=g) * 60 + 150
    yy, xx = torch.meshgrid(torch.arange(size, dtype=torch.float32), torch.arange(size, dtype=torch.float32), indexing="ij")
    for b, l in zip(target.bbox, target.get_field("labels")):
        cx, cy = (b[0] + b[2]) / 2, (b[1] + b[3]) / 2
        rx, ry = (b[2] - b[0]) / 2 + 1, (b[3] - b[1]) / 2 + 1
        inside = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 < 1
        img[:, inside] -= 40.0 if int(l) == 1 else 90.0
    return img.clamp(0, 255) - torch.tensor(PIXEL_MEAN).view(3, 1, 1)


def colour_view(g, img, strength=0.15):
    """brightness / contrast jitter of an already-normalised image (one of the K+S views)."""
    mean = torch.tensor(PIXEL_MEAN).view(3, 1, 1)
    raw = img + mean
    b = 1 + (float(torch.rand(1, generator=g)) * 2 - 1) * strength
    c = 1 + (float(torch.rand(1, generator=g)) * 2 - 1) * strength
    raw = (raw * b - raw.mean()) * c + raw.mean()
    return raw.clamp(0, 255) - mean


def make_batch(seed, batch=2, size=800, n_views=3, n_cyto=20, n_nuc=20, pin=False):
    """Returns (images [B,3,S,S], targets list[BoxList], unlabeled list of n_views tensors [B,3,S,S])."""
    g = torch.Generator().manual_seed(seed)
    targets = [make_targets(g, size, n_cyto, n_nuc) for _ in range(batch)]
    images = torch.stack([make_image(g, t, size) for t in targets])
    u_targets = [make_targets(g, size, n_cyto, n_nuc) for _ in range(batch)]
    u_base = [make_image(g, t, size) for t in u_targets]
    views = [torch.stack([colour_view(g, im) for im in u_base]) for _ in range(n_views)]
    if pin:
        images = images.pin_memory()
        views = [v.pin_memory() for v in views]
    return images, targets, views
