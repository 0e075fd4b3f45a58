"""Box head: FPN2MLP feature extractor, FPN predictor, Fast R-CNN loss + PSM consistency loss, post-processor.

Mirrors reference modeling/roi_heads/box_head/{box_head.py:12-122, roi_box_feature_extractors.py:45-131,
roi_box_predictors.py:33-51, loss.py:12-315, inference.py:11-161}.  fc6/fc7/cls/bbox run on the conv engine
(1x1 conv over ROIs); the pooled features stay NHWC so fc6's columns are stored in (h, w, c) order internally
(converted at the state-dict boundary)."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import functional as Fk
from .. import kernels as K
from ..layers import Conv2d, smooth_l1_loss
from ..layers.conv import kaiming_uniform_ohwi_
from ..structures.bounding_box import BoxList
from ..structures.boxlist_ops import boxlist_iou, cat_boxlist
from . import registry
from .balanced_positive_negative_sampler import BalancedPositiveNegativeSampler
from .box_coder import BoxCoder
from .matcher import Matcher
from .poolers import Pooler
from .utils import cat


class LinearParam(nn.Module):
    """nn.Linear-compatible parameter holder (weight [out, in], bias [out])."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features))

    def forward(self, x, relu=False):
        return Fk.linear(x, self.weight, self.bias, relu)


def split_nonzero(masks):
    """[mask.nonzero().squeeze(1) for mask in masks] with one `nonzero` over the concatenation (the only point where the host
    waits for the device) and one read-back of the per-mask counts, which finds the queue already drained."""
    if len(masks) == 1:
        return [masks[0].nonzero().squeeze(1)]
    sizes = [int(m.numel()) for m in masks]
    idx = torch.cat(list(masks)).nonzero().squeeze(1)
    starts = [0]
    for n in sizes:
        starts.append(starts[-1] + n)
    cuts = torch.searchsorted(idx, torch.tensor(starts, dtype=idx.dtype, device=idx.device)).tolist()
    return [idx[cuts[i]:cuts[i + 1]] - starts[i] for i in range(len(masks))]


class MaskRCNNFPNAdaptor(nn.Module):
    """MGD 1x1 adaptors 256 -> 128 per level (reference roi_box_feature_extractors.py:45-75)."""

    def __init__(self, cfg):
        super().__init__()
        out = 128 if cfg.MT.T_ADAPT is True else 256
        for i in range(1, 6):
            m = Conv2d(256, out, 1, 1, 0)
            kaiming_uniform_ohwi_(m.weight, a=1)
            nn.init.constant_(m.bias, 0)
            self.add_module("adapter_%d" % i, m)

    def forward(self, feats):
        from ..utils.streams import run_parallel
        level = lambda i: (lambda: getattr(self, "adapter_%d" % (i + 1))(feats[i]))
        return run_parallel([level(i) for i in range(5)], feats[0].device)


@registry.ROI_BOX_FEATURE_EXTRACTORS.register("FPN2MLPFeatureExtractor")
class FPN2MLPFeatureExtractor(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        h = cfg.MODEL.ROI_BOX_HEAD
        res = h.POOLER_RESOLUTION
        self.pooler = Pooler((res, res), h.POOLER_SCALES, h.POOLER_SAMPLING_RATIO)
        self.resolution, self.channels = res, cfg.MODEL.BACKBONE.OUT_CHANNELS
        self.fc6 = LinearParam(self.channels * res ** 2, h.MLP_HEAD_DIM)   # columns in (h, w, c) order
        self.fc7 = LinearParam(h.MLP_HEAD_DIM, h.MLP_HEAD_DIM)
        self.dropout = h.DO
        for l in (self.fc6, self.fc7):
            nn.init.kaiming_uniform_(l.weight, a=1)
            nn.init.constant_(l.bias, 0)

    def forward(self, x, proposals, filp=False, istrain=False):
        x = self.pooler(x, proposals)
        x = x.reshape(x.shape[0], -1)
        x = self.fc6(x, relu=True)
        x = self.fc7(x, relu=True)
        if self.dropout > 0:
            x = F.dropout(x, p=self.dropout, training=istrain)
        return x

    # fc6 columns: reference order (c, h, w)  <->  internal (h, w, c)
    def fc6_reference_weight(self):
        r, c = self.resolution, self.channels
        return self.fc6.weight.detach().view(-1, r, r, c).permute(0, 3, 1, 2).reshape(-1, c * r * r).contiguous()

    def load_fc6_reference_weight(self, w):
        r, c = self.resolution, self.channels
        with torch.no_grad():
            self.fc6.weight.copy_(w.view(-1, c, r, r).permute(0, 2, 3, 1).reshape(-1, c * r * r))


def make_roi_box_feature_extractor(cfg):
    return registry.ROI_BOX_FEATURE_EXTRACTORS[cfg.MODEL.ROI_BOX_HEAD.FEATURE_EXTRACTOR](cfg)


class FPNPredictor(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        n = cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES
        d = cfg.MODEL.ROI_BOX_HEAD.MLP_HEAD_DIM
        self.num_classes = n
        self.cls_score = LinearParam(d, n)
        self.bbox_pred = LinearParam(d, n * 4)
        nn.init.normal_(self.cls_score.weight, std=0.01)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)

    def forward(self, x):
        n = self.num_classes
        pad = (-5 * n) % 32
        w = torch.cat([self.cls_score.weight, self.bbox_pred.weight, x.new_zeros((pad, x.shape[1]))], 0)
        b = torch.cat([self.cls_score.bias, self.bbox_pred.bias, x.new_zeros(pad)], 0)
        o = Fk.linear(x, w, b)
        return o[:, :n], o[:, n:5 * n]


_ROI_BOX_PREDICTOR = {"FPNPredictor": FPNPredictor}


def make_roi_box_predictor(cfg):
    return _ROI_BOX_PREDICTOR[cfg.MODEL.ROI_BOX_HEAD.PREDICTOR](cfg)


class PostProcessor(nn.Module):
    """Reference box_head/inference.py:11-145.  The per-class NMS calls of all images (B x (C-1) `_C.nms` calls
    plus a `.cpu()` kthvalue in the reference) are one segmented on-device NMS: score-thresholded entries are
    given score -1 so they sort behind — and therefore can never suppress — any valid detection."""

    def __init__(self, score_thresh=0.05, nms=0.5, detections_per_img=100, box_coder=None, cfg=None):
        super().__init__()
        self.score_thresh, self.nms, self.detections_per_img = score_thresh, nms, detections_per_img
        self.box_coder = box_coder if box_coder is not None else BoxCoder(weights=(10., 10., 5., 5.))
        self.cfg = cfg

    @torch.no_grad()
    def forward(self, x, boxes, head_id=None):
        class_logits, box_regression = x
        class_prob = F.softmax(class_logits, -1)
        image_shapes = [b.size for b in boxes]
        per_image = [len(b) for b in boxes]
        device = class_logits.device
        concat = torch.cat([b.bbox for b in boxes], dim=0)
        C = class_prob.shape[1]
        # BoxCoder.decode of every class's deltas against its ROI, one launch (clipping happens per image below)
        decoded, _, _ = K.decode_clip_boxes(box_regression.reshape(-1, 4), concat.repeat_interleave(C, dim=0),
                                            self.box_coder.weights, self.box_coder.bbox_xform_clip)
        decoded = decoded.reshape(sum(per_image), -1)
        decoded = decoded.split(per_image, dim=0)
        class_prob = class_prob.split(per_image, dim=0)
        # segments in (image, class) order, class-major inside an image: row = class block + roi
        seg_boxes, seg_scores, offs, img_bounds = [], [], [0], [0]
        for prob, bx, (w, h), r in zip(class_prob, decoded, image_shapes, per_image):
            bx = bx.reshape(-1, C, 4)[:, 1:].permute(1, 0, 2).clone()          # [C-1, R, 4]
            bx[..., 0].clamp_(min=0, max=w - 1)
            bx[..., 1].clamp_(min=0, max=h - 1)
            bx[..., 2].clamp_(min=0, max=w - 1)
            bx[..., 3].clamp_(min=0, max=h - 1)
            sc = prob[:, 1:].t()                                                # [C-1, R]
            seg_boxes.append(bx.reshape(-1, 4))
            seg_scores.append(torch.where(sc > self.score_thresh, sc, sc.new_full((), -1.0)).reshape(-1))
            for _ in range(C - 1):
                offs.append(offs[-1] + r)
            img_bounds.append(offs[-1])
        flat_b, flat_s = torch.cat(seg_boxes, 0), torch.cat(seg_scores, 0)
        T = flat_s.shape[0]
        keep, counts = K.nms_segmented(flat_b, flat_s, offs, self.nms)
        # kept flags per row on the device (entry p of segment s is live iff p < counts[s]); thresholded rows carry
        # score -1: they sort behind — and so never suppress — a valid detection, and are dropped here
        seg_id, seg_start, pos = self._segment_rows(tuple(offs), device)
        live = pos < counts.to(torch.int64)[seg_id]
        flags = torch.zeros(T + 1, dtype=torch.bool, device=device)
        flags.index_fill_(0, torch.where(live, keep + seg_start, torch.full_like(keep, T)), True)
        flags = flags[:T] & (flat_s > 0)
        if self.detections_per_img > 0:   # keep the detections_per_img best of an image, ties included (kthvalue rule)
            masked = torch.where(flags, flat_s, flat_s.new_full((), -1.0))
            for i in range(len(boxes)):
                a, b = img_bounds[i], img_bounds[i + 1]
                if b - a > self.detections_per_img:
                    thresh = torch.topk(masked[a:b], self.detections_per_img, sorted=True)[0][-1]
                    flags[a:b] &= masked[a:b] >= thresh
        kept = flags.nonzero().squeeze(1)
        bounds = torch.searchsorted(kept, torch.tensor(img_bounds, dtype=torch.int64, device=device)).tolist()
        seg_label = (seg_id % (C - 1) + 1) if C > 1 else seg_id
        out = []
        for i in range(len(boxes)):
            idx = kept[bounds[i]:bounds[i + 1]]
            bl = BoxList(flat_b[idx], image_shapes[i], mode="xyxy")
            bl.add_field("scores", flat_s[idx])
            bl.add_field("objectness", flat_s[idx])
            bl.add_field("labels", seg_label[idx])
            out.append(bl)
        return out

    def _segment_rows(self, offs, device):
        key = (offs, str(device))
        cache = self.__dict__.setdefault("_seg_cache", {})
        if key not in cache:
            if len(cache) > 16:
                cache.clear()
            seg_id = torch.repeat_interleave(torch.arange(len(offs) - 1), torch.tensor([offs[i + 1] - offs[i] for i in range(len(offs) - 1)]))
            seg_start = torch.tensor(offs[:-1], dtype=torch.int64)[seg_id]
            pos = torch.arange(offs[-1], dtype=torch.int64) - seg_start
            cache[key] = (seg_id.to(device), seg_start.to(device), pos.to(device))
        return cache[key]


def make_roi_box_post_processor(cfg):
    h = cfg.MODEL.ROI_HEADS
    return PostProcessor(h.SCORE_THRESH, h.NMS, h.DETECTIONS_PER_IMG, BoxCoder(weights=h.BBOX_REG_WEIGHTS), cfg=cfg)


class FastRCNNLossComputation(object):
    def __init__(self, proposal_matcher, fg_bg_sampler, box_coder, cfg=None):
        self.proposal_matcher, self.fg_bg_sampler, self.box_coder, self.cfg = proposal_matcher, fg_bg_sampler, box_coder, cfg

    def match_targets_to_proposals(self, proposal, target):
        matched_idxs = self.proposal_matcher.match(target, proposal)
        target = target.copy_with_fields(["labels"])
        matched = target[matched_idxs.clamp(min=0)]
        matched.add_field("matched_idxs", matched_idxs)
        return matched

    def prepare_targets(self, proposals, targets):
        labels, regression_targets = [], []
        for p, t in zip(proposals, targets):
            midx = self.proposal_matcher.match(t, p)
            lab = t.get_field("labels")[midx.clamp(min=0)].to(dtype=torch.int64)
            lab[midx == Matcher.BELOW_LOW_THRESHOLD] = 0
            lab[midx == Matcher.BETWEEN_THRESHOLDS] = -1
            labels.append(lab)
            regression_targets.append(K.encode_boxes(t.bbox, midx, p.bbox, self.box_coder.weights))
        return labels, regression_targets

    def subsample(self, proposals, targets):
        labels, regression_targets = self.prepare_targets(proposals, targets)
        pos, neg = self.fg_bg_sampler(labels)
        proposals = list(proposals)
        for lab, reg, p in zip(labels, regression_targets, proposals):
            p.add_field("labels", lab)
            p.add_field("regression_targets", reg)
        # the sampled rows of every image with ONE device drain (the reference: a nonzero = a host sync per image)
        for i, idx in enumerate(split_nonzero([pi | ni for pi, ni in zip(pos, neg)])):
            proposals[i] = proposals[i][idx]
        self._proposals = proposals
        return proposals

    def __call__(self, class_logits, box_regression):
        class_logits = cat(class_logits, dim=0)
        box_regression = cat(box_regression, dim=0)
        if not hasattr(self, "_proposals"):
            raise RuntimeError("subsample needs to be called before")
        proposals = self._proposals
        labels = cat([p.get_field("labels") for p in proposals], dim=0)
        regression_targets = cat([p.get_field("regression_targets") for p in proposals], dim=0)
        classification_loss = F.cross_entropy(class_logits, labels)
        # class-specific deltas of every ROI, reduced under the foreground mask (reference box_head/loss.py:146-160
        # gathers the foreground rows with nonzero(): same terms, no host sync)
        fg = labels > 0
        map_inds = 4 * labels.clamp(min=0)[:, None] + torch.arange(4, device=class_logits.device)
        n = torch.abs(box_regression.gather(1, map_inds) - regression_targets)
        l1 = torch.where(n < 1, 0.5 * n ** 2, n - 0.5)
        box_loss = torch.where(fg[:, None], l1, l1.new_zeros(())).sum()
        return classification_loss, box_loss / labels.numel()

    def evaluatePSM(self, class_logits, class_logits_t, proposals):
        """PSM consistency loss (reference box_head/loss.py:185-237): one fused CUDA launch per student view."""
        mt = self.cfg.MT
        if not (mt.RANK_FILTER > 0 and mt.CLS_LOSS_TYPE == "bce" and mt.HARD_NEG):
            raise NotImplementedError("only the train_mt.sh PSM configuration (RANK_FILTER>0, 'bce', HARD_NEG) is built")
        labels = cat([p.get_field("labels") for p in proposals], dim=0)
        teacher = torch.stack([t.detach() for t in class_logits_t]).contiguous()
        losses = []
        for s in class_logits:
            loss, _ = Fk.psm_loss(s.contiguous(), teacher, labels, mt.TEMP, bool(mt.SHARPEN), True, mt.CLS_BALANCE_WEIGHT)
            losses.append(loss)
        return torch.mean(torch.stack(losses), dim=0)


def make_roi_box_loss_evaluator(cfg):
    h = cfg.MODEL.ROI_HEADS
    matcher = Matcher(h.FG_IOU_THRESHOLD, h.BG_IOU_THRESHOLD, allow_low_quality_matches=False,
                      top_k=cfg.MODEL.ROI_BOX_HEAD.K_HEAD)
    sampler = BalancedPositiveNegativeSampler(h.BATCH_SIZE_PER_IMAGE, h.POSITIVE_FRACTION)
    return FastRCNNLossComputation(matcher, sampler, BoxCoder(weights=h.BBOX_REG_WEIGHTS), cfg=cfg)


class ROIBoxHead(nn.Module):
    def __init__(self, cfg, relation=True):
        super().__init__()
        self.feature_extractor = make_roi_box_feature_extractor(cfg)
        self.predictor = make_roi_box_predictor(cfg)
        self.post_processor = make_roi_box_post_processor(cfg)
        self.loss_evaluator = make_roi_box_loss_evaluator(cfg)
        self.use_realation_nms = relation and cfg.MODEL.RELATION_NMS.USE_RELATION_NMS
        self.cfg = cfg
        self.mode = None

    def set_teacher_mode(self, mode):
        self.mode = mode

    def forward(self, features, proposals, targets=None):
        if self.training:
            with torch.no_grad():
                proposals = self.loss_evaluator.subsample(proposals, targets)
        x = self.feature_extractor(features, proposals, istrain=self.training)
        class_logits, box_regression = self.predictor(x)
        if not self.training:
            if not self.use_realation_nms:
                proposals = self.post_processor((class_logits, box_regression), proposals)
            return x, proposals, {}, class_logits, box_regression
        loss_classifier, loss_box_reg = self.loss_evaluator([class_logits], [box_regression])
        return x, proposals, dict(loss_classifier=loss_classifier, loss_box_reg=loss_box_reg), class_logits, box_regression

    def _forward_single(self, proposals, targets, feats_list, istrain=False, batched_feats=None):
        if targets is not None:
            proposals = self.loss_evaluator.subsample(proposals, targets)
        proposals_B = [p.transpose(0) for p in proposals]  # h-flipped boxes for the flipped views
        n_views = len(feats_list)
        if batched_feats is not None and not istrain and n_views > 1:
            # the views are slices of ONE stacked trunk output [n_views*B, ...]: pool / fc6 / fc7 / predict all views in
            # one launch each (image index = view * B + b); rows of a GEMM are independent, so every view's logits
            # are the numbers the per-view loop below produces
            boxes = []
            for i in range(n_views):
                boxes += proposals if i % 2 == 0 else proposals_B
            x = self.feature_extractor(batched_feats, boxes, istrain=False)
            c, r = self.predictor(x)
            per_view = x.shape[0] // n_views
            return (list(x.split(per_view)), list(c.split(per_view)), list(r.split(per_view)), proposals)
        feats, logits, regs = [], [], []
        for i, feat in enumerate(feats_list):
            x = self.feature_extractor(feat, proposals if i % 2 == 0 else proposals_B, istrain=istrain)
            c, r = self.predictor(x)
            feats.append(x)
            logits.append(c)
            regs.append(r)
        return feats, logits, regs, proposals

    def forward_teacher(self, feature_tuple, proposals, targets, batched_feats=None):
        feats, logits, regs, proposals = self._forward_single(proposals, targets, feature_tuple, istrain=False,
                                                              batched_feats=batched_feats)
        return feats, proposals, {}, logits, regs

    def forward_student(self, features, proposals, class_logits_t):
        _, logits, _, _ = self._forward_single(proposals, None, features, istrain=True)
        if self.cfg.MT.ODKD:
            raise NotImplementedError("MT.ODKD (comparison method) is out of scope")
        if self.cfg.MT.CLS_LOSS > 0:
            cls_loss = self.loss_evaluator.evaluatePSM(logits, class_logits_t, proposals)
        else:
            cls_loss = torch.zeros((1,), device=logits[0].device)
        return dict(mt_classifier=cls_loss)


def build_roi_box_head(cfg, relation=True):
    return ROIBoxHead(cfg, relation=relation)
