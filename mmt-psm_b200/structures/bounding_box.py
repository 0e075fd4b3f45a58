"""BoxList — same container contract as the reference's `structures/bounding_box.py:9-266`:
`bbox [n,4]` fp32, `size=(width,height)`, `mode in {"xyxy","xywh"}`, a dict of per-box extra fields that
are indexed together with the boxes, and the legacy "+1" pixel-extent convention (TO_REMOVE = 1)."""
import torch

FLIP_LEFT_RIGHT = 0
FLIP_TOP_BOTTOM = 1
_ONE = 1  # legacy inclusive-pixel convention


class BoxList(object):
    def __init__(self, bbox, image_size, mode="xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2 or bbox.size(-1) != 4:
            raise ValueError("bbox should be [n,4], got %s" % (tuple(bbox.shape),))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size  # (width, height)
        self.mode = mode
        self.extra_fields = {}

    # ---- fields
    def add_field(self, field, data):
        self.extra_fields[field] = data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def _copy_extra_fields(self, other):
        for k, v in other.extra_fields.items():
            self.extra_fields[k] = v

    def _derive(self, bbox, size=None, mode="xyxy", field_fn=None):
        out = BoxList(bbox, self.size if size is None else size, mode=mode)
        for k, v in self.extra_fields.items():
            if field_fn is not None and not isinstance(v, torch.Tensor):
                v = field_fn(v)
            out.add_field(k, v)
        return out

    # ---- coordinate systems
    def _split_into_xyxy(self):
        if self.mode == "xyxy":
            return self.bbox.split(1, dim=-1)
        x, y, w, h = self.bbox.split(1, dim=-1)
        return x, y, x + (w - _ONE).clamp(min=0), y + (h - _ONE).clamp(min=0)

    def convert(self, mode):
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        if mode == self.mode:
            return self
        x1, y1, x2, y2 = self._split_into_xyxy()
        if mode == "xyxy":
            box = torch.cat((x1, y1, x2, y2), dim=-1)
        else:
            box = torch.cat((x1, y1, x2 - x1 + _ONE, y2 - y1 + _ONE), dim=-1)
        out = BoxList(box, self.size, mode=mode)
        out._copy_extra_fields(self)
        return out

    def resize(self, size, *args, **kwargs):
        rw, rh = (float(s) / float(o) for s, o in zip(size, self.size))
        if rw == rh:
            out = self._derive(self.bbox * rw, size, self.mode, lambda v: v.resize(size, *args, **kwargs))
            return out
        x1, y1, x2, y2 = self._split_into_xyxy()
        box = torch.cat((x1 * rw, y1 * rh, x2 * rw, y2 * rh), dim=-1)
        return self._derive(box, size, "xyxy", lambda v: v.resize(size, *args, **kwargs)).convert(self.mode)

    def transpose(self, method):
        if method not in (FLIP_LEFT_RIGHT, FLIP_TOP_BOTTOM):
            raise NotImplementedError("Only FLIP_LEFT_RIGHT and FLIP_TOP_BOTTOM implemented")
        W, H = self.size
        x1, y1, x2, y2 = self._split_into_xyxy()
        if method == FLIP_LEFT_RIGHT:
            box = torch.cat((W - x2 - _ONE, y1, W - x1 - _ONE, y2), dim=-1)
        else:  # the reference flips rows WITHOUT the -1 (bounding_box.py:153-157)
            box = torch.cat((x1, H - y2, x2, H - y1), dim=-1)
        return self._derive(box, None, "xyxy", lambda v: v.transpose(method)).convert(self.mode)

    def crop(self, box):
        x1, y1, x2, y2 = self._split_into_xyxy()
        w, h = box[2] - box[0], box[3] - box[1]
        cropped = torch.cat(((x1 - box[0]).clamp(min=0, max=w), (y1 - box[1]).clamp(min=0, max=h),
                             (x2 - box[0]).clamp(min=0, max=w), (y2 - box[1]).clamp(min=0, max=h)), dim=-1)
        return self._derive(cropped, (w, h), "xyxy", lambda v: v.crop(box)).convert(self.mode)

    # ---- tensor-like
    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def clip_to_image(self, remove_empty=True):
        W, H = self.size
        self.bbox[:, 0].clamp_(min=0, max=W - _ONE)
        self.bbox[:, 1].clamp_(min=0, max=H - _ONE)
        self.bbox[:, 2].clamp_(min=0, max=W - _ONE)
        self.bbox[:, 3].clamp_(min=0, max=H - _ONE)
        if remove_empty:
            b = self.bbox
            return self[(b[:, 3] > b[:, 1]) & (b[:, 2] > b[:, 0])]
        return self

    def area(self):
        b = self.bbox
        if self.mode == "xyxy":
            return (b[:, 2] - b[:, 0] + _ONE) * (b[:, 3] - b[:, 1] + _ONE)
        return b[:, 2] * b[:, 3]

    def copy_with_fields(self, fields):
        out = BoxList(self.bbox, self.size, self.mode)
        for f in (fields if isinstance(fields, (list, tuple)) else [fields]):
            out.add_field(f, self.get_field(f))
        return out

    def __repr__(self):
        return "BoxList(num_boxes=%d, image_width=%s, image_height=%s, mode=%s)" % (
            len(self), self.size[0], self.size[1], self.mode)
