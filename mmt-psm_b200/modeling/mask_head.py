"""Mask head: FPN feature extractor (4x conv3x3), C4 predictor (deconv 2x2 + 1x1), loss, post-processing.

Mirrors reference modeling/roi_heads/mask_head/{mask_head.py:16-156, roi_mask_feature_extractors.py:91-146,
roi_mask_predictors.py:9-36, loss.py:37-180, inference.py:14-266}; with MODEL.RELATION_MASK.USE_RELATION the IR-Net
refinement branch (modeling/relation.py) adds its deeply-supervised second set of logits.  B200
differences: all convs are tcgen05 launches on NHWC ROI tiles; mask TARGETS are rasterised on the device
(`mmt_rasterize_polygons_f32`) instead of the per-ROI pycocotools loop on the host; the teacher's
Masker paste loop (one 800x800 CPU canvas per detection) is one paste-and-sum kernel producing the integral
foreground map MGD consumes."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import functional as Fk
from .. import kernels as K
from ..layers import Conv2d, ConvTranspose2x2
from ..layers.conv import kaiming_normal_fanout_ohwi_
from ..structures.bounding_box import BoxList
from ..structures.boxlist_ops import boxlist_iou
from .matcher import Matcher
from .poolers import Pooler
from .utils import cat


def keep_only_positive_boxes(boxes):
    assert isinstance(boxes, (list, tuple)) and isinstance(boxes[0], BoxList)
    from .box_head import split_nonzero
    positive_inds = [b.get_field("labels") > 0 for b in boxes]
    positive_boxes = [b[idx] for b, idx in zip(boxes, split_nonzero(positive_inds))]   # one device drain for the batch
    return positive_boxes, positive_inds


class MaskRCNNFPNFeatureExtractor(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        h = cfg.MODEL.ROI_MASK_HEAD
        res = h.POOLER_RESOLUTION
        self.pooler = Pooler((res, res), h.POOLER_SCALES, h.POOLER_SAMPLING_RATIO)
        nxt = cfg.MODEL.BACKBONE.OUT_CHANNELS
        self.blocks = []
        for i, feats in enumerate(h.CONV_LAYERS, 1):
            name = "mask_fcn%d" % i
            m = Conv2d(nxt, feats, 3, 1, 1)
            kaiming_normal_fanout_ohwi_(m.weight)
            nn.init.constant_(m.bias, 0)
            self.add_module(name, m)
            self.blocks.append(name)
            nxt = feats

    def forward(self, x, proposals):
        x = self.pooler(x, proposals)
        pre = x
        for name in self.blocks:
            x = getattr(self, name)(x, relu=True)
        return x, pre


_ROI_MASK_FEATURE_EXTRACTORS = {"MaskRCNNFPNFeatureExtractor": MaskRCNNFPNFeatureExtractor}


def make_roi_mask_feature_extractor(cfg):
    return _ROI_MASK_FEATURE_EXTRACTORS[cfg.MODEL.ROI_MASK_HEAD.FEATURE_EXTRACTOR](cfg)


class MaskRCNNC4Predictor(nn.Module):
    _PAD = 32

    def __init__(self, cfg):
        super().__init__()
        self.num_classes = cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES
        dim = cfg.MODEL.ROI_MASK_HEAD.CONV_LAYERS[-1]
        assert cfg.MODEL.ROI_HEADS.USE_FPN, "only the FPN configuration is on the MMT-PSM path"
        self.conv5_mask = ConvTranspose2x2(dim, dim)
        self.mask_fcn_logits = Conv2d(dim, self.num_classes, 1, 1, 0)
        kaiming_normal_fanout_ohwi_(self.conv5_mask.weight, fan_out=dim * 4)  # torch's fan_out of the IOHW weight
        kaiming_normal_fanout_ohwi_(self.mask_fcn_logits.weight)
        nn.init.constant_(self.mask_fcn_logits.bias, 0)

    def forward(self, x):
        """x NHWC [P,14,14,C] -> logits NHWC [P,28,28,num_classes]."""
        x = self.conv5_mask(x, relu=True)
        n, pad = self.num_classes, self._PAD - self.num_classes
        w = torch.cat([self.mask_fcn_logits.weight, x.new_zeros((pad, 1, 1, x.shape[3]))], 0)
        b = torch.cat([self.mask_fcn_logits.bias, x.new_zeros(pad)], 0)
        return Fk.conv_bn_act(x, w, None, b, None, False, 1, 0)[..., :n]


_ROI_MASK_PREDICTOR = {"MaskRCNNC4Predictor": MaskRCNNC4Predictor}


def make_roi_mask_predictor(cfg):
    return _ROI_MASK_PREDICTOR[cfg.MODEL.ROI_MASK_HEAD.PREDICTOR](cfg)


def project_masks_on_boxes(segmentation_masks, proposals, instance_idx, discretization_size):
    """Device-side mask targets: ROI k gets the polygons of instance `instance_idx[k]` of `segmentation_masks`
    cropped to proposal k and rasterised at M x M (reference mask_head/loss.py:37-75, bit-exact)."""
    assert segmentation_masks.size == proposals.size, "%s, %s" % (segmentation_masks, proposals)
    proposals = proposals.convert("xyxy")
    if len(proposals) == 0:
        return torch.empty(0, dtype=torch.float32, device=proposals.bbox.device)
    packed = segmentation_masks.packed(proposals.bbox.device)
    return K.rasterize_polygons(packed, proposals.bbox, instance_idx, discretization_size)


class MaskRCNNLossComputation(object):
    def __init__(self, proposal_matcher, discretization_size, cfg=None):
        self.cfg, self.proposal_matcher, self.discretization_size = cfg, proposal_matcher, discretization_size

    def prepare_targets(self, proposals, targets):
        """Reference mask_head/loss.py:101-149.  The reference rasterises only the rows whose re-matched label is
        positive (a `nonzero` per image = a host sync); here every row gets a target and `__call__` weights the rows:
        the mask head is only ever handed the box head's positive ROIs (keep_only_positive_boxes) and both heads match
        with the same thresholds, so the weights are all ones in practice and the sums are the reference's."""
        labels, masks = [], []
        for p, t in zip(proposals, targets):
            matched_idxs = self.proposal_matcher.match(t, p)
            lab = t.get_field("labels")[matched_idxs.clamp(min=0)].to(dtype=torch.int64)
            lab = torch.where(matched_idxs == Matcher.BELOW_LOW_THRESHOLD, lab.new_zeros(()), lab)
            masks.append(project_masks_on_boxes(t.get_field("masks"), p, matched_idxs.clamp(min=0),
                                                self.discretization_size))
            labels.append(lab)
        return labels, masks

    def __call__(self, proposals, mask_logits, targets):
        """mask_logits NHWC [P, M, M, num_classes]."""
        if mask_logits.shape[0] == 0:
            return mask_logits.sum() * 0
        labels, mask_targets = self.prepare_targets(proposals, targets)
        labels = cat(labels, dim=0)
        mask_targets = cat(mask_targets, dim=0)
        rows = torch.arange(labels.shape[0], device=labels.device)
        sel = mask_logits[rows, :, :, labels]
        w = (labels > 0).to(sel.dtype)
        bce = F.binary_cross_entropy_with_logits(sel, mask_targets, reduction="none")
        # mean over the positive rows' pixels == the reference's BCE over the gathered positive rows
        return (bce * w[:, None, None]).sum() / (w.sum() * (sel.shape[1] * sel.shape[2])).clamp(min=1)


def make_roi_mask_loss_evaluator(cfg):
    h = cfg.MODEL.ROI_HEADS
    matcher = Matcher(h.FG_IOU_THRESHOLD, h.BG_IOU_THRESHOLD, allow_low_quality_matches=False)
    return MaskRCNNLossComputation(matcher, cfg.MODEL.ROI_MASK_HEAD.RESOLUTION, cfg)


class MaskPostProcessor(nn.Module):
    """Selects the predicted class's mask probability (reference mask_head/inference.py:14-64).  With a masker the
    reference pastes every mask into a full-image uint8 canvas on the CPU; here `mask` holds the 28x28
    probabilities and, for the teacher, `integral_mask()` does the paste-and-sum on the device."""

    def __init__(self, masker=None):
        super().__init__()
        self.masker = masker

    @torch.no_grad()
    def forward(self, x, boxes):
        """x NHWC logits [D, M, M, C]."""
        prob = x.sigmoid()
        labels = torch.cat([b.get_field("labels") for b in boxes])
        idx = torch.arange(x.shape[0], device=labels.device)
        prob = prob[idx, :, :, labels][:, None]  # [D,1,M,M]
        prob = prob.split([len(b) for b in boxes], dim=0)
        results = []
        for p, box in zip(prob, boxes):
            out = BoxList(box.bbox, box.size, mode="xyxy")
            for f in box.fields():
                out.add_field(f, box.get_field(f))
            out.add_field("mask_prob", p)
            if self.masker is None:
                out.add_field("mask", p)
            results.append(out)
        return results


class Masker(object):
    """Paste parameters of the reference's Masker (mask_head/inference.py:209-246): threshold 0.5, padding 1."""

    def __init__(self, threshold=0.5, padding=1):
        self.threshold, self.padding = threshold, padding

    def integral_mask(self, boxlists, height, width):
        """sum over detections of the pasted binary masks, [B, height, width] float — the only thing the
        mean-teacher path consumes from the pasted masks (reference generalized_rcnn.py:129-132)."""
        assert self.padding == 1
        probs = torch.cat([b.get_field("mask_prob")[:, 0] for b in boxlists], 0)
        boxes = torch.cat([b.bbox for b in boxlists], 0)
        ids = torch.cat([torch.full((len(b),), i, dtype=torch.int32, device=boxes.device) for i, b in enumerate(boxlists)])
        return K.paste_masks_sum(probs, boxes, ids, len(boxlists), height, width, self.threshold)


def make_roi_mask_post_processor(cfg):
    h = cfg.MODEL.ROI_MASK_HEAD
    return MaskPostProcessor(Masker(h.POSTPROCESS_MASKS_THRESHOLD, 1) if h.POSTPROCESS_MASKS else None)


def make_roi_mask_generator(cfg):
    return MaskPostProcessor(Masker(cfg.MODEL.ROI_MASK_HEAD.POSTPROCESS_MASKS_THRESHOLD, 1))


class ROIMaskHead(nn.Module):
    def __init__(self, cfg, is_student=False):
        super().__init__()
        self.cfg = cfg.clone()
        self.feature_extractor = make_roi_mask_feature_extractor(cfg)
        self.predictor = make_roi_mask_predictor(cfg)
        # the reference builds the refinement net whenever the RELATION_MASK node exists (mask_head.py:48, always);
        # its parameters only matter with USE_RELATION, so it is only built (and held in HBM / the optimizer) then
        self.mask_relation_module = None
        if cfg.MODEL.RELATION_MASK.USE_RELATION:
            from .relation import MaskRelationRefineNet
            self.mask_relation_module = MaskRelationRefineNet(cfg, self.predictor)
        self.post_processor = make_roi_mask_post_processor(cfg)
        self.mask_generator = make_roi_mask_generator(cfg)
        self.loss_evaluator = make_roi_mask_loss_evaluator(cfg)
        self.mode = None

    def set_teacher_mode(self, mode):
        self.mode = mode

    def forward(self, features, proposals, targets=None, images=None):
        if self.training:
            all_proposals = proposals
            proposals, positive_inds = keep_only_positive_boxes(proposals)
        x, _ = self.feature_extractor(features, proposals)
        mask_logits = self.predictor(x)
        loss_seg = self.loss_evaluator(proposals, mask_logits, targets) if self.training else None
        if self.mask_relation_module is not None:
            # IR-Net: second set of logits from instance-to-instance attention, rows re-ordered per image by
            # (class, objectness) — all images in one pass (reference mask_head.py:96-124 loops over images)
            mask_logits, proposals = self.mask_relation_module.refine(x, mask_logits, proposals)
        if not self.training:
            if self.mode is None or self.mode == "train":
                result = self.post_processor(mask_logits, proposals)
            else:
                result = self.mask_generator(mask_logits, proposals)
            return x, result, {}
        if self.mask_relation_module is not None:
            loss_2 = self.loss_evaluator(proposals, mask_logits, targets)
            loss_seg = 0.5 * (loss_seg + loss_2) if self.cfg.MODEL.RELATION_MASK.DEEP_SUPER else loss_2
        return x, all_proposals, dict(loss_seg=loss_seg)


def build_roi_mask_head(cfg, is_student=False):
    return ROIMaskHead(cfg, is_student)
