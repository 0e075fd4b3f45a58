"""BoxList helpers — same functions as the reference's `structures/boxlist_ops.py:9-135`; `boxlist_nms` calls
the on-device NMS of libmmtpsm_b200 instead of `_C.nms` + forced `.to('cuda:0')` (boxlist_ops.py:27-30)."""
import torch

from .bounding_box import BoxList


def boxlist_nms(boxlist, nms_thresh, max_proposals=-1, score_field="score"):
    if nms_thresh <= 0:
        return boxlist
    from ..layers import nms as _nms
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    keep = _nms(boxlist.bbox, boxlist.get_field(score_field), nms_thresh)
    if max_proposals > 0:
        keep = keep[:max_proposals]
    return boxlist[keep.to(boxlist.bbox.device)].convert(mode)


def remove_small_boxes(boxlist, min_size):
    wh = boxlist.convert("xywh").bbox
    keep = ((wh[:, 2] >= min_size) & (wh[:, 3] >= min_size)).nonzero().squeeze(1)
    return boxlist[keep]


def boxlist_iou(boxlist1, boxlist2):
    """IoU matrix [N,M] with the +1 pixel convention (boxlist_ops.py:57-92)."""
    if boxlist1.size != boxlist2.size:
        raise RuntimeError("boxlists should have same image size, got %s, %s" % (boxlist1, boxlist2))
    a1, a2 = boxlist1.area(), boxlist2.area()
    b1, b2 = boxlist1.bbox, boxlist2.bbox
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    return inter / (a1[:, None] + a2 - inter)


def _cat(tensors, dim=0):
    assert isinstance(tensors, (list, tuple))
    return tensors[0] if len(tensors) == 1 else torch.cat(tensors, dim)


def cat_boxlist(bboxes):
    assert isinstance(bboxes, (list, tuple)) and all(isinstance(b, BoxList) for b in bboxes)
    size, mode, fields = bboxes[0].size, bboxes[0].mode, set(bboxes[0].fields())
    assert all(b.size == size and b.mode == mode and set(b.fields()) == fields for b in bboxes)
    out = BoxList(_cat([b.bbox for b in bboxes], dim=0), size, mode)
    for f in bboxes[0].fields():
        out.add_field(f, _cat([b.get_field(f) for b in bboxes], dim=0))
    return out
