"""FPN anchors (reference modeling/rpn/anchor_generator.py:34-265): per level one size x 3 aspect ratios,
cell anchors built with the Detectron rounding recipe, shifted over the feature grid in (y, x, anchor) order."""
import numpy as np
import torch
from torch import nn

from ..structures.bounding_box import BoxList


def _mk(ws, hs, cx, cy):
    ws, hs = ws[:, None], hs[:, None]
    return np.hstack((cx - 0.5 * (ws - 1), cy - 0.5 * (hs - 1), cx + 0.5 * (ws - 1), cy + 0.5 * (hs - 1)))


def _whc(a):
    w, h = a[2] - a[0] + 1, a[3] - a[1] + 1
    return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)


def generate_anchors(stride=16, sizes=(32, 64, 128, 256, 512), aspect_ratios=(0.5, 1, 2)):
    scales = np.array(sizes, dtype=float) / stride
    ratios = np.array(aspect_ratios, dtype=float)
    base = np.array([1, 1, stride, stride], dtype=float) - 1
    w, h, cx, cy = _whc(base)
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    per_ratio = _mk(ws, hs, cx, cy)
    out = []
    for a in per_ratio:
        w, h, cx, cy = _whc(a)
        out.append(_mk(w * scales, h * scales, cx, cy))
    return torch.from_numpy(np.vstack(out))


class BufferList(nn.Module):
    """Buffers addressable by index, so the state-dict keys read `cell_anchors.0 ...` like the reference's."""

    def __init__(self, buffers):
        super().__init__()
        for i, b in enumerate(buffers):
            self.register_buffer(str(i), b)

    def __len__(self):
        return len(self._buffers)

    def __iter__(self):
        return iter(self._buffers.values())


class AnchorGenerator(nn.Module):
    def __init__(self, sizes=(128, 256, 512), aspect_ratios=(0.5, 1.0, 2.0), anchor_strides=(8, 16, 32),
                 straddle_thresh=0):
        super().__init__()
        if len(anchor_strides) == 1:
            cells = [generate_anchors(anchor_strides[0], sizes, aspect_ratios).float()]
        else:
            if len(anchor_strides) != len(sizes):
                raise RuntimeError("FPN should have #anchor_strides == #sizes")
            cells = [generate_anchors(s, (size,), aspect_ratios).float() for s, size in zip(anchor_strides, sizes)]
        self.strides = anchor_strides
        self.cell_anchors = BufferList(cells)
        self.straddle_thresh = straddle_thresh
        self._cache = {}

    def num_anchors_per_location(self):
        return [len(c) for c in self.cell_anchors]

    def grid_anchors(self, grid_sizes):
        key = (tuple(tuple(g) for g in grid_sizes), str(next(iter(self.cell_anchors)).device))
        if key not in self._cache:
            out = []
            for (gh, gw), stride, base in zip(grid_sizes, self.strides, self.cell_anchors):
                dev = base.device
                sx = torch.arange(0, gw * stride, step=stride, dtype=torch.float32, device=dev)
                sy = torch.arange(0, gh * stride, step=stride, dtype=torch.float32, device=dev)
                yy, xx = torch.meshgrid(sy, sx, indexing="ij")
                xx, yy = xx.reshape(-1), yy.reshape(-1)
                shifts = torch.stack((xx, yy, xx, yy), dim=1)
                out.append((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
            self._cache = {key: out}
        return self._cache[key]

    def add_visibility_to(self, boxlist):
        W, H = boxlist.size
        a = boxlist.bbox
        if self.straddle_thresh >= 0:
            t = self.straddle_thresh
            inside = (a[..., 0] >= -t) & (a[..., 1] >= -t) & (a[..., 2] < W + t) & (a[..., 3] < H + t)
        else:
            inside = torch.ones(a.shape[0], dtype=torch.bool, device=a.device)
        boxlist.add_field("visibility", inside)

    def forward(self, image_list, grid_sizes):
        """grid_sizes: list of (h, w) of the feature maps (layout independent).  The BoxLists (with their `visibility`
        field) only depend on the grid and image sizes: cached, fixed-size crops re-use them every iteration."""
        key = (tuple(tuple(g) for g in grid_sizes), tuple(tuple(s) for s in image_list.image_sizes),
               str(next(iter(self.cell_anchors)).device))
        if getattr(self, "_fwd_cache", None) is not None and self._fwd_cache[0] == key:
            return [list(per_image) for per_image in self._fwd_cache[1]]
        per_level = self.grid_anchors(grid_sizes)
        anchors = []
        for (ih, iw) in image_list.image_sizes:
            in_image = []
            for a in per_level:
                bl = BoxList(a, (iw, ih), mode="xyxy")
                self.add_visibility_to(bl)
                in_image.append(bl)
            anchors.append(in_image)
        self._fwd_cache = (key, anchors)
        return [list(per_image) for per_image in anchors]


def make_anchor_generator(config):
    r = config.MODEL.RPN
    if r.USE_FPN:
        assert len(r.ANCHOR_STRIDE) == len(r.ANCHOR_SIZES), "FPN should have len(ANCHOR_STRIDE) == len(ANCHOR_SIZES)"
    else:
        assert len(r.ANCHOR_STRIDE) == 1, "Non-FPN should have a single ANCHOR_STRIDE"
    return AnchorGenerator(r.ANCHOR_SIZES, r.ASPECT_RATIOS, r.ANCHOR_STRIDE, r.STRADDLE_THRESH)
