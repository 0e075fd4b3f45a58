"""SegmentationMask / Polygons — per-image instance masks as polygon lists, same container contract as the
reference's `structures/segmentation_mask.py:53-247`.  The reference rasterises polygons on the HOST through
pycocotools for every positive ROI (modeling/roi_heads/mask_head/loss.py:37-75); here the polygons of an image
are packed once into flat device arrays (`packed()`) and rasterised by the CUDA kernel
`mmt_rasterize_polygons_f32` (bit-exact port of pycoco/maskApi.c:166-206 `rleFrPoly` + decode + merge)."""
import torch

from .bounding_box import FLIP_LEFT_RIGHT, FLIP_TOP_BOTTOM


class Polygons(object):
    """All polygons of ONE instance: list of flat [x0,y0,x1,y1,...] fp32 tensors."""

    def __init__(self, polygons, size, mode=None):
        if isinstance(polygons, Polygons):
            polygons = polygons.polygons
        self.polygons = [torch.as_tensor(p, dtype=torch.float32) for p in polygons]
        self.size = size
        self.mode = mode

    def transpose(self, method):
        if method not in (FLIP_LEFT_RIGHT, FLIP_TOP_BOTTOM):
            raise NotImplementedError("Only FLIP_LEFT_RIGHT and FLIP_TOP_BOTTOM implemented")
        W, H = self.size
        dim, idx = (W, 0) if method == FLIP_LEFT_RIGHT else (H, 1)
        out = []
        for poly in self.polygons:
            p = poly.clone()
            p[idx::2] = dim - poly[idx::2] - 1
            out.append(p)
        return Polygons(out, self.size, self.mode)

    def crop(self, box):
        w, h = max(box[2] - box[0], 1), max(box[3] - box[1], 1)
        out = []
        for poly in self.polygons:
            p = poly.clone()
            p[0::2] -= box[0]
            p[1::2] -= box[1]
            out.append(p)
        return Polygons(out, (w, h), self.mode)

    def resize(self, size, *args, **kwargs):
        rw, rh = (float(s) / float(o) for s, o in zip(size, self.size))
        out = []
        for poly in self.polygons:
            p = poly.clone()
            p[0::2] *= rw
            p[1::2] *= rh
            out.append(p)
        return Polygons(out, size, self.mode)

    def __len__(self):
        return len(self.polygons)

    def __repr__(self):
        return "Polygons(num_polygons=%d, image_width=%s, image_height=%s, mode=%s)" % (
            len(self.polygons), self.size[0], self.size[1], self.mode)


class SegmentationMask(object):
    """All instances of one image."""

    def __init__(self, polygons, size, mode=None):
        assert isinstance(polygons, list)
        self.polygons = [p if isinstance(p, Polygons) else Polygons(p, size, mode) for p in polygons]
        self.size = size
        self.mode = mode
        self._packed = {}

    def transpose(self, method):
        return SegmentationMask([p.transpose(method) for p in self.polygons], self.size, self.mode)

    def crop(self, box):
        return SegmentationMask([p.crop(box) for p in self.polygons], (box[2] - box[0], box[3] - box[1]), self.mode)

    def resize(self, size, *args, **kwargs):
        return SegmentationMask([p.resize(size, *args, **kwargs) for p in self.polygons], size, self.mode)

    def to(self, *args, **kwargs):
        return self

    def __getitem__(self, item):
        if isinstance(item, (int, slice)):
            sel = self.polygons[item] if isinstance(item, slice) else [self.polygons[item]]
        else:
            if isinstance(item, torch.Tensor):
                if item.dtype in (torch.uint8, torch.bool):
                    item = item.nonzero().flatten()
                item = item.tolist()
            sel = [self.polygons[i] for i in item]
        return SegmentationMask(list(sel), self.size, self.mode)

    def __iter__(self):
        return iter(self.polygons)

    def __len__(self):
        return len(self.polygons)

    def packed(self, device):
        """(xy [2*V] fp32, poly_offsets [n_poly+1] int32 in vertices, inst_offsets [n_inst+1] int32 in polygons)
        on `device`; built once per image and cached."""
        key = str(device)
        if key not in self._packed:
            xy, poly_off, inst_off = [], [0], [0]
            for inst in self.polygons:
                for p in inst.polygons:
                    xy.append(p.reshape(-1))
                    poly_off.append(poly_off[-1] + p.numel() // 2)
                inst_off.append(len(poly_off) - 1)
            flat = torch.cat(xy) if xy else torch.zeros(0)
            host = (flat, torch.tensor(poly_off, dtype=torch.int32), torch.tensor(inst_off, dtype=torch.int32))
            if torch.device(device).type == "cuda":
                # through pinned staging: a host -> device copy from PAGEABLE memory first synchronises the stream it is
                # enqueued on — the caller (MTtrainer.stage) would sit out the rest of the running iteration here
                host = tuple(t.pin_memory() if (t.device.type == "cpu" and t.numel()) else t for t in host)
            self._packed[key] = tuple(t.to(device, non_blocking=True) for t in host)
        return self._packed[key]

    def __repr__(self):
        return "SegmentationMask(num_instances=%d, image_width=%s, image_height=%s)" % (
            len(self.polygons), self.size[0], self.size[1])
