"""Box <-> delta transform (reference modeling/box_coder.py:22-95): legacy +1 extents, weights (wx,wy,ww,wh),
dw/dh clamped at log(1000/16), decoded x2/y2 carry the -1."""
import math

import torch


class BoxCoder(object):
    def __init__(self, weights, bbox_xform_clip=math.log(1000. / 16)):
        self.weights = weights
        self.bbox_xform_clip = bbox_xform_clip

    def encode(self, reference_boxes, proposals):
        one = 1
        ex_w = proposals[:, 2] - proposals[:, 0] + one
        ex_h = proposals[:, 3] - proposals[:, 1] + one
        ex_cx = proposals[:, 0] + 0.5 * ex_w
        ex_cy = proposals[:, 1] + 0.5 * ex_h
        gt_w = reference_boxes[:, 2] - reference_boxes[:, 0] + one
        gt_h = reference_boxes[:, 3] - reference_boxes[:, 1] + one
        gt_cx = reference_boxes[:, 0] + 0.5 * gt_w
        gt_cy = reference_boxes[:, 1] + 0.5 * gt_h
        wx, wy, ww, wh = self.weights
        return torch.stack((wx * (gt_cx - ex_cx) / ex_w, wy * (gt_cy - ex_cy) / ex_h,
                            ww * torch.log(gt_w / ex_w), wh * torch.log(gt_h / ex_h)), dim=1)

    def decode(self, rel_codes, boxes):
        boxes = boxes.to(rel_codes.dtype)
        one = 1
        w = boxes[:, 2] - boxes[:, 0] + one
        h = boxes[:, 3] - boxes[:, 1] + one
        cx = boxes[:, 0] + 0.5 * w
        cy = boxes[:, 1] + 0.5 * h
        wx, wy, ww, wh = self.weights
        dx = rel_codes[:, 0::4] / wx
        dy = rel_codes[:, 1::4] / wy
        dw = torch.clamp(rel_codes[:, 2::4] / ww, max=self.bbox_xform_clip)
        dh = torch.clamp(rel_codes[:, 3::4] / wh, max=self.bbox_xform_clip)
        pcx = dx * w[:, None] + cx[:, None]
        pcy = dy * h[:, None] + cy[:, None]
        pw = torch.exp(dw) * w[:, None]
        ph = torch.exp(dh) * h[:, None]
        out = torch.zeros_like(rel_codes)
        out[:, 0::4] = pcx - 0.5 * pw
        out[:, 1::4] = pcy - 0.5 * ph
        out[:, 2::4] = pcx + 0.5 * pw - 1
        out[:, 3::4] = pcy + 0.5 * ph - 1
        return out
