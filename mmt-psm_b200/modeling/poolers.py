"""FPN ROI pooler (reference modeling/poolers.py:11-121): LevelMapper (FPN eq. 1) + per-level ROIAlign.
The per-level loop (nonzero / ROIAlign / index_put per level) is one fused launch over NHWC maps."""
import torch
from torch import nn

from .. import functional as Fk
from .. import kernels as K
from .utils import cat


class LevelMapper(object):
    def __init__(self, k_min, k_max, canonical_scale=224, canonical_level=4, eps=1e-6):
        self.k_min, self.k_max = k_min, k_max
        self.s0, self.lvl0, self.eps = canonical_scale, canonical_level, eps

    def __call__(self, boxlists):
        s = torch.sqrt(cat([b.area() for b in boxlists]))
        lv = torch.floor(self.lvl0 + torch.log2(s / self.s0 + self.eps))
        lv = torch.clamp(lv, min=self.k_min, max=self.k_max)
        return lv.to(torch.int64) - self.k_min


class Pooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio):
        super().__init__()
        self.output_size = output_size
        self.scales = tuple(float(s) for s in scales)
        self.sampling_ratio = sampling_ratio
        lvl_min = -torch.log2(torch.tensor(scales[0], dtype=torch.float32)).item()
        lvl_max = -torch.log2(torch.tensor(scales[-1], dtype=torch.float32)).item()
        self.map_levels = LevelMapper(lvl_min, lvl_max)

    def convert_to_roi_format(self, boxes):
        concat = cat([b.bbox for b in boxes], dim=0)
        ids = cat([torch.full((len(b), 1), i, dtype=concat.dtype, device=concat.device) for i, b in enumerate(boxes)], dim=0)
        return torch.cat([ids, concat], dim=1)

    def forward(self, x, boxes):
        """x: list of NHWC level maps (only the first len(scales) are used); boxes: list[BoxList].
        Returns NHWC [K, R, R, C]."""
        n = len(self.scales)
        if n == 1:
            rois = self.convert_to_roi_format(boxes)
            levels = torch.zeros(rois.shape[0], dtype=torch.int32, device=rois.device)
        else:
            # ROI table (convert_to_roi_format) and LevelMapper in ONE launch instead of ~15 elementwise ones
            m = self.map_levels
            rois, levels = K.fpn_rois(cat([b.bbox for b in boxes], dim=0), [len(b) for b in boxes], m.k_min, m.k_max,
                                      m.s0, m.lvl0, m.eps)
        return Fk.fpn_roi_align(list(x[:n]), self.scales, rois, levels, self.output_size[0], self.sampling_ratio)
