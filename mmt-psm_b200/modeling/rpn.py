"""RPN: head, proposal selection and loss.

Mirrors reference modeling/rpn/rpn.py:14-221 (RPNHead / RPNModule incl. `forward_teacher`),
rpn/inference.py:15-272 (RPNPostProcessor) and rpn/loss.py:21-214 (RPNLossComputation) — same cfg keys, same
selection rules, same loss formulae.  B200 differences:
  * the head's 3x3 conv (+bias+ReLU) and the fused cls(3)+bbox(12) 1x1 conv are tcgen05 launches on NHWC maps,
    so the reference's permute(0,2,3,1).reshape flattening is a free view;
  * the 5 levels x B images NMS problems of one pass (10 `_C.nms` calls with 10 blocking D2H copies in the
    reference, csrc/cuda/nms.cu:100) are ONE segmented on-device NMS returning per-row keep flags; candidate gather,
    box decoding, target matching / encoding, sampling and the loss itself are single kernels of libmmtpsm_b200
    (csrc/proposals.cu, csrc/losses.cu) instead of chains of elementwise torch launches.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .. import kernels as K
from ..layers import Conv2d, smooth_l1_loss
from ..structures.bounding_box import BoxList
from ..structures.boxlist_ops import boxlist_iou, cat_boxlist
from . import registry
from .anchor_generator import make_anchor_generator
from .balanced_positive_negative_sampler import BalancedPositiveNegativeSampler
from .box_coder import BoxCoder
from .matcher import Matcher
from .utils import cat

_HEAD_PAD = 32  # fused cls+bbox output channels are padded to one 32-wide MMA tile


@registry.RPN_HEADS.register("SingleConvRPNHead")
class RPNHead(nn.Module):
    def __init__(self, cfg, in_channels, num_anchors):
        super().__init__()
        self.conv = Conv2d(in_channels, in_channels, 3, 1, 1)
        self.cls_logits = Conv2d(in_channels, num_anchors, 1, 1, 0)
        self.bbox_pred = Conv2d(in_channels, num_anchors * 4, 1, 1, 0)
        self.num_anchors = num_anchors
        for l in (self.conv, self.cls_logits, self.bbox_pred):
            nn.init.normal_(l.weight, std=0.01)
            nn.init.constant_(l.bias, 0)

    def forward_fused(self, x):
        """x: list of NHWC maps -> per level ONE tensor [B,h,w,32]: channels [0,A) objectness, [A,5A) box deltas."""
        A = self.num_anchors
        pad = _HEAD_PAD - 5 * A
        w = torch.cat([self.cls_logits.weight, self.bbox_pred.weight,
                       self.conv.weight.new_zeros((pad, 1, 1, self.conv.in_channels))], 0)
        b = torch.cat([self.cls_logits.bias, self.bbox_pred.bias, self.conv.bias.new_zeros(pad)], 0)
        from .. import functional as Fk
        from ..utils.streams import run_parallel
        level = lambda feat: (lambda: Fk.conv_bn_act(self.conv(feat, relu=True), w, None, b, None, False, 1, 0))
        return run_parallel([level(feat) for feat in x], x[0].device)   # the levels are independent: parallel streams

    @torch.no_grad()
    def proposal_scores(self, fused):
        """sigmoid(objectness) of every anchor as dense [B, h*w*A] rows (h, w, anchor order) — what the proposal
        selection ranks (reference rpn/inference.py:88-92); no gradient flows through proposals."""
        A = self.num_anchors
        return [torch.sigmoid(o[..., :A]).reshape(o.shape[0], -1) for o in fused]

    def split(self, fused):
        A = self.num_anchors
        return [o[..., :A] for o in fused], [o[..., A:5 * A] for o in fused]

    def forward(self, x):
        """x: list of NHWC maps.  Returns (objectness [B,h,w,A], bbox_reg [B,h,w,4A]) per level."""
        return self.split(self.forward_fused(x))


class RPNPostProcessor(nn.Module):
    def __init__(self, pre_nms_top_n, post_nms_top_n, nms_thresh, min_size, box_coder=None, fpn_post_nms_top_n=None,
                 is_teacher=False):
        super().__init__()
        self.pre_nms_top_n, self.post_nms_top_n = pre_nms_top_n, post_nms_top_n
        self.nms_thresh, self.min_size = nms_thresh, min_size
        self.box_coder = box_coder if box_coder is not None else BoxCoder(weights=(1.0, 1.0, 1.0, 1.0))
        self.fpn_post_nms_top_n = post_nms_top_n if fpn_post_nms_top_n is None else fpn_post_nms_top_n
        self.is_teacher = is_teacher

    def add_gt_proposals(self, proposals, targets):
        device = proposals[0].bbox.device
        out = []
        for p, t in zip(proposals, targets):
            gt = t.copy_with_fields([])
            gt.add_field("objectness", torch.ones(len(gt), device=device))
            out.append(cat_boxlist((p, gt)))
        return out

    def _candidates(self, anchors, scores_l, fused, A_, sizes, image_shapes, device):
        """Per (image, level) the `sizes[l]` best-scoring anchors decoded to boxes, their scores, and the keep flags of the
        per-(image, level) NMS (min-size filter folded in): (boxes [N*K_, 4], flat_scores [N*K_], flags [N*K_]).

        `self.shared` (set by RPNModule to one dict for its train- and test-mode selectors): the teacher post-processes the
        SAME head output twice — with the test-mode numbers for its coarse inference, then with the train-mode numbers for
        the proposals the student is trained on (reference generalized_rcnn.py:127-139).  The candidates of a level are
        sorted by score and greedy NMS decides every box from the higher-scored ones only, so the smaller selection's boxes,
        scores and keep flags are exactly the leading rows of the larger one's: the first call computes the larger set,
        the second reads its answer out of it — one top-k, one decode and one NMS chain per teacher pass instead of two."""
        N, num_levels = scores_l[0].shape[0], len(scores_l)
        shared = getattr(self, "shared", None)
        if shared is not None and not shared.get("active"):
            shared = None     # sharing is scoped to ONE teacher pass (RPNModule.candidate_sharing): the head maps of a CUDA-graph
                              # trunk keep their addresses from one iteration to the next, so a key alone cannot tell passes apart
        key = tuple(t.data_ptr() for t in scores_l) + (N,)
        if shared is not None and shared.get("key") == key:
            big = shared["sizes"]
            if all(a <= b for a, b in zip(sizes, big)):
                if list(sizes) == list(big):
                    return shared["boxes"], shared["scores"], shared["flags"].clone()   # callers edit their flags in place
                cols = shared["cols"].get(tuple(sizes))
                if cols is None:
                    off, parts = 0, []
                    for a, b in zip(sizes, big):
                        parts.append(torch.arange(off, off + a, device=device))
                        off += b
                    cols = shared["cols"][tuple(sizes)] = torch.cat(parts)
                Kb = sum(big)
                rows = (torch.arange(N, device=device)[:, None] * Kb + cols[None, :]).reshape(-1)
                return shared["boxes"][rows], shared["scores"][rows], shared["flags"][rows]
        want = sizes
        if shared is not None and shared.get("max_sizes") is not None:
            want = [max(a, min(b, s.shape[1])) for a, b, s in zip(sizes, shared["max_sizes"], scores_l)]
        K_ = sum(want)
        offs, image_wh = self._segments(N, tuple(want), image_shapes, device)
        deltas = torch.empty((N, K_, 4), dtype=torch.float32, device=device)
        anch = torch.empty((N, K_, 4), dtype=torch.float32, device=device)
        cand_scores = torch.empty((N, K_), dtype=torch.float32, device=device)
        level_anchors = [anchors[0][l].bbox for l in range(num_levels)]
        if not (device.type == "cuda" and K.rpn_topk_gather(scores_l, fused, A_, level_anchors, want, deltas, anch, cand_scores)):
            col = 0   # pre_nms_top_n beyond the fused kernel's capacity: torch's top-k + the gather kernel, per level
            for l, (s_l, k) in enumerate(zip(scores_l, want)):
                vals, idx = s_l.topk(k, dim=1, sorted=True)
                K.rpn_gather(fused[l], A_, idx, vals, level_anchors[l], deltas, anch, cand_scores, col)
                col += k
        flat_scores = cand_scores.reshape(-1)                 # image-major: row = j * K_ + (level offset + rank)
        boxes, scores_nms, valid = K.decode_clip_boxes(deltas.reshape(-1, 4), anch.reshape(-1, 4), self.box_coder.weights,
                                                       self.box_coder.bbox_xform_clip, image_wh, K_, flat_scores,
                                                       self.min_size)
        # filtered (too small) boxes carry score -1: they sort behind, and so can never suppress, a valid box
        flags = K.nms_segmented_flags(boxes, scores_nms, offs, self.nms_thresh)
        if valid is not None:
            flags &= valid
        if shared is not None:
            shared.update(key=key, sizes=list(want), boxes=boxes, scores=flat_scores, flags=flags.clone())
            if list(want) != list(sizes):
                return self._candidates(anchors, scores_l, fused, A_, sizes, image_shapes, device)   # the smaller view of it
        return boxes, flat_scores, flags

    def _segments(self, N, sizes, image_shapes, device):
        """Static description of the (image, level) NMS problems of a pass, cached per shape: segment offsets (a
        ctypes int array, image-major) and the per-image (width, height) table on the device."""
        import ctypes
        key = (N, sizes, image_shapes, str(device))
        cache = self.__dict__.setdefault("_seg_cache", {})
        if key not in cache:
            if len(cache) > 16:
                cache.clear()
            offs = [0]
            for j in range(N):
                for k in sizes:
                    offs.append(offs[-1] + k)
            wh = torch.tensor([[float(w), float(h)] for (w, h) in image_shapes], dtype=torch.float32, device=device)
            cache[key] = ((ctypes.c_int * len(offs))(*offs), wh)
        return cache[key]

    @torch.no_grad()
    def forward(self, anchors, objectness, box_regression, targets=None, head_out=None):
        """Reference rpn/inference.py:78-137 (per level and image: sigmoid, top-k, decode, clip, small-box filter, NMS,
        keep post_nms_top_n) + :216-244 (selection over levels), restructured so that every (image, level) problem of
        the pass goes through the SAME few launches: per level one top-k and one gather kernel (deltas from the fused
        head map, anchors, scores -> pass-wide candidate tables), then one fused decode/clip/filter kernel, one
        segmented NMS that returns per-row keep flags, one selection.  One or two host read-backs per pass."""
        num_levels = len(objectness)
        N = objectness[0].shape[0]
        device = objectness[0].device
        image_shapes = tuple(tuple(a[0].size) for a in anchors)
        if head_out is None:   # called with bare maps: rebuild the fused view of the head output
            A_ = objectness[0].shape[-1]
            fused = [torch.cat([o, b], -1).contiguous() for o, b in zip(objectness, box_regression)]
            scores_l = [o.reshape(N, -1).sigmoid() for o in objectness]
        else:
            fused, scores_l = head_out
            A_ = objectness[0].shape[-1]
        sizes = [min(self.pre_nms_top_n, s.shape[1]) for s in scores_l]
        K_ = sum(sizes)
        boxes, flat_scores, flags = self._candidates(anchors, scores_l, fused, A_, sizes, image_shapes, device)
        if self.post_nms_top_n > 0 and max(sizes) > self.post_nms_top_n:   # keep[:post_nms_top_n] per (image, level)
            f2, a = flags.view(N, K_), 0
            for k in sizes:
                if k > self.post_nms_top_n:
                    f2[:, a:a + k] &= f2[:, a:a + k].cumsum(1) <= self.post_nms_top_n
                a += k
        if num_levels > 1 and not self.training:
            # per-image top-k in score order (reference :236-243) straight on the masked score rows: dropped rows
            # carry -1 (< any sigmoid score), so the first n_valid entries of the sorted top-k are the reference's
            # selection; ONE read-back (the per-image counts) for the whole pass
            f2, s2 = flags.view(N, K_), flat_scores.view(N, K_)
            k = min(self.fpn_post_nms_top_n, K_)
            vals, inds = torch.where(f2, s2, s2.new_full((), -1.0)).topk(k, dim=1, sorted=True)
            n_valid = (vals >= 0).sum(1).tolist()
            boxlists = []
            for j in range(N):
                idx = inds[j, :n_valid[j]] + j * K_
                bl = BoxList(boxes[idx], image_shapes[j], mode="xyxy")
                bl.add_field("objectness", flat_scores[idx])
                boxlists.append(bl)
            return boxlists
        # ascending row index = image-major, level-major, descending score inside a level: the reference's cat order
        kept = flags.nonzero().squeeze(1)
        if num_levels > 1 and self.training:
            # batch-wide top-k over the concatenated kept boxes (reference :222-234), original order preserved
            k = min(self.fpn_post_nms_top_n, kept.numel())
            _, inds = torch.topk(flat_scores[kept], k, dim=0, sorted=True)
            kept = kept[inds.sort()[0]]
        per_image = torch.bincount(torch.div(kept, K_, rounding_mode="floor"), minlength=N).tolist()
        boxlists, a = [], 0
        for j in range(N):
            idx = kept[a:a + per_image[j]]
            a += per_image[j]
            bl = BoxList(boxes[idx], image_shapes[j], mode="xyxy")
            bl.add_field("objectness", flat_scores[idx])
            boxlists.append(bl)
        if self.training and targets is not None:
            boxlists = self.add_gt_proposals(boxlists, targets)
        return boxlists

    def select_over_all_levels(self, boxlists):
        n = len(boxlists)
        if self.training:  # batch-wide top-k (reference rpn/inference.py:222-234)
            objectness = torch.cat([b.get_field("objectness") for b in boxlists], dim=0)
            box_sizes = [len(b) for b in boxlists]
            k = min(self.fpn_post_nms_top_n, len(objectness))
            _, inds = torch.topk(objectness, k, dim=0, sorted=True)
            mask = torch.zeros_like(objectness, dtype=torch.bool)
            mask[inds] = True
            mask = mask.split(box_sizes)
            for i in range(n):
                boxlists[i] = boxlists[i][mask[i]]
        else:
            for i in range(n):
                objectness = boxlists[i].get_field("objectness")
                k = min(self.fpn_post_nms_top_n, len(objectness))
                _, inds = torch.topk(objectness, k, dim=0, sorted=True)
                boxlists[i] = boxlists[i][inds]
        return boxlists


def make_rpn_postprocessor(config, rpn_box_coder, is_train, is_teacher=False):
    r = config.MODEL.RPN
    if is_train:
        pre, post, fpn_post = r.PRE_NMS_TOP_N_TRAIN, r.POST_NMS_TOP_N_TRAIN, r.FPN_POST_NMS_TOP_N_TRAIN
    else:
        pre, post, fpn_post = r.PRE_NMS_TOP_N_TEST, r.POST_NMS_TOP_N_TEST, r.FPN_POST_NMS_TOP_N_TEST
    return RPNPostProcessor(pre, post, r.NMS_THRESH, r.MIN_SIZE, rpn_box_coder, fpn_post, is_teacher)


class RPNLossComputation(object):
    def __init__(self, proposal_matcher, fg_bg_sampler, box_coder, cfg=None):
        self.proposal_matcher = proposal_matcher
        self.fg_bg_sampler = fg_bg_sampler
        self.box_coder = box_coder

    def match_targets_to_anchors(self, anchor, target):
        matched_idxs = self.proposal_matcher.match(target, anchor)
        target = target.copy_with_fields([])
        matched = target[matched_idxs.clamp(min=0)]
        matched.add_field("matched_idxs", matched_idxs)
        return matched

    def prepare_targets(self, anchors, targets):
        """Reference rpn/loss.py:47-83; the masked writes are selects (no index lists -> no host sync)."""
        labels, regression_targets = [], []
        for a, t in zip(anchors, targets):
            midx = self.proposal_matcher.match(t, a)
            lab = (midx >= 0).to(dtype=torch.float32)
            ignore = ~a.get_field("visibility").bool() | (midx == Matcher.BETWEEN_THRESHOLDS)
            labels.append(torch.where(ignore, lab.new_full((), -1.0), lab))
            # BoxCoder.encode(target[midx.clamp(0)], anchors) — gather + ~25 elementwise ops — as one launch
            regression_targets.append(K.encode_boxes(t.bbox, midx, a.bbox, self.box_coder.weights))
        return labels, regression_targets

    def _cat_anchors(self, anchors):
        """cat_boxlist over the levels of every image, cached on the (cached) anchor grids."""
        key = tuple(id(a) for per_image in anchors for a in per_image)
        if getattr(self, "_anchor_cat", None) is None or self._anchor_cat[0] != key:
            self._anchor_cat = (key, [cat_boxlist(a) for a in anchors], anchors)  # keeps the keyed objects alive
        return self._anchor_cat[1]

    def _sample(self, anchors, targets):
        """labels / regression targets per anchor and the sampled positive / negative MASKS over the whole batch."""
        anchors = self._cat_anchors(anchors)
        labels, regression_targets = self.prepare_targets(anchors, targets)
        pos, neg = self.fg_bg_sampler(labels)
        return labels, regression_targets, torch.cat(pos, dim=0), torch.cat(neg, dim=0)

    @staticmethod
    def _targets_key(targets):
        return tuple((t.bbox.data_ptr(), len(t)) for t in targets)

    @torch.no_grad()
    def precompute(self, anchors, targets):
        """Everything of the RPN loss that does not depend on the network's output — anchor / ground-truth matching, box
        encoding, the sampler's draw — for the pass that is about to run; `__call__` picks it up if it is then handed the
        same targets.  The trainer calls this right after it has enqueued the trunks, while the device is still busy with
        the previous iteration: ~60 small launches leave the window between the trunk's completion and the heads."""
        self._pre = (self._targets_key(targets), self._sample(anchors, targets))

    def teacher_sample_selection(self, anchors, objectness, box_regression, targets):
        """Reference rpn/loss.py:85-136.  Its outputs are never used downstream; the only observable
        effect is the sampler's draw from the random stream.  When the sampler draws on the host (parity mode,
        `BalancedPositiveNegativeSampler.rng_device = "cpu"`) the draw is reproduced so the iteration consumes the
        reference's random sequence; with the device generator there is no sequence to preserve and the dead
        matching + sampling work is skipped."""
        if self.fg_bg_sampler.rng_device is None:
            return None, None, None, None
        labels, _, pos, neg = self._sample(anchors, targets)
        pos, neg = torch.nonzero(pos).squeeze(1), torch.nonzero(neg).squeeze(1)
        sampled = torch.cat([pos, neg], dim=0)
        obj = cat([o.reshape(o.shape[0], -1) for o in objectness], dim=1).reshape(-1)
        reg = cat([b.reshape(b.shape[0], -1, 4) for b in box_regression], dim=1).reshape(-1, 4)
        labels = torch.cat(labels, dim=0)
        return obj[sampled], reg[pos], (sampled, pos.shape), labels[sampled]

    def __call__(self, anchors, objectness, box_regression, targets, fused=None):
        """Reference rpn/loss.py:138-196.  The reference gathers the sampled anchors into index lists (two
        `nonzero`s = two host syncs) and reduces over them; here the same sums run over all anchors under the
        sampling masks — with the fused head maps at hand, in ONE kernel (functional.rpn_loss) that reads the five
        per-level maps in place, instead of flattening / concatenating them and ~35 elementwise launches."""
        pre, self._pre = getattr(self, "_pre", None), None
        if pre is not None and pre[0] == self._targets_key(targets):
            labels, regression_targets, pos, neg = pre[1]
        else:
            labels, regression_targets, pos, neg = self._sample(anchors, targets)
        labels = torch.cat(labels, dim=0)
        regression_targets = torch.cat(regression_targets, dim=0)
        if fused is not None and all(f.is_cuda and f.is_contiguous() for f in fused):
            from .. import functional as Fk
            objectness_loss, box_loss = Fk.rpn_loss(list(fused), objectness[0].shape[-1], labels, regression_targets,
                                                    pos, neg, 1.0 / 9)
            return objectness_loss, box_loss
        sampled = pos | neg
        n_sampled = sampled.sum().to(torch.float32)
        # NHWC maps flatten directly to the (h, w, anchor) order of the labels
        obj = cat([o.reshape(o.shape[0], -1) for o in objectness], dim=1).reshape(-1)
        reg = cat([b.reshape(b.shape[0], -1, 4) for b in box_regression], dim=1).reshape(-1, 4)
        beta = 1.0 / 9
        n = torch.abs(reg - regression_targets)
        l1 = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
        box_loss = torch.where(pos[:, None], l1, l1.new_zeros(())).sum() / n_sampled
        bce = F.binary_cross_entropy_with_logits(obj, labels.clamp(min=0), reduction="none")
        objectness_loss = torch.where(sampled, bce, bce.new_zeros(())).sum() / n_sampled
        return objectness_loss, box_loss


def make_rpn_loss_evaluator(cfg, box_coder):
    r = cfg.MODEL.RPN
    matcher = Matcher(r.FG_IOU_THRESHOLD, r.BG_IOU_THRESHOLD, allow_low_quality_matches=True)
    sampler = BalancedPositiveNegativeSampler(r.BATCH_SIZE_PER_IMAGE, r.POSITIVE_FRACTION)
    return RPNLossComputation(matcher, sampler, box_coder, cfg)


class RPNModule(nn.Module):
    def __init__(self, cfg, is_teacher=False):
        super().__init__()
        self.mode = None
        self.cfg = cfg.clone()
        self.anchor_generator = make_anchor_generator(cfg)
        in_channels = cfg.MODEL.BACKBONE.OUT_CHANNELS
        self.head = registry.RPN_HEADS[cfg.MODEL.RPN.RPN_HEAD](
            cfg, in_channels, self.anchor_generator.num_anchors_per_location()[0])
        coder = BoxCoder(weights=(1.0, 1.0, 1.0, 1.0))
        self.box_selector_train = make_rpn_postprocessor(cfg, coder, is_train=True, is_teacher=is_teacher)
        self.box_selector_test = make_rpn_postprocessor(cfg, coder, is_train=False, is_teacher=is_teacher)
        if is_teacher and (self.box_selector_train.nms_thresh, self.box_selector_train.min_size) == (
                self.box_selector_test.nms_thresh, self.box_selector_test.min_size):
            # the teacher runs both selectors on one head output (RPNPostProcessor._candidates): shared candidate tables
            shared = {"cols": {}, "max_sizes": [max(self.box_selector_train.pre_nms_top_n, self.box_selector_test.pre_nms_top_n)] * 8}
            self.box_selector_train.shared = self.box_selector_test.shared = shared
        self.loss_evaluator = make_rpn_loss_evaluator(cfg, coder)

    def set_teacher_mode(self, mode):
        self.mode = mode

    def candidate_sharing(self, on):
        """Scope of the teacher's shared candidate tables: switched on for the duration of ONE `forward_teacher` pass."""
        shared = getattr(self.box_selector_train, "shared", None)
        if shared is not None:
            shared["active"] = bool(on) and shared.get("enabled", True)
            shared["key"] = None
            if not on:
                for k in ("boxes", "scores", "flags"):
                    shared.pop(k, None)

    def _head_and_anchors(self, images, features, head_out=None):
        if isinstance(features, list):
            raise NotImplementedError("multi-view RPN averaging is dead code on this path (SURVEY §2.3): the RPN is "
                                      "always handed ONE feature tuple")
        # `head_out`: (fused head maps, proposal scores) already computed with the trunk (inside its CUDA graph)
        if head_out is None:
            fused = self.head.forward_fused(features)
            head_out = (fused, self.head.proposal_scores(fused))
        self._head_out = head_out
        objectness, reg = self.head.split(head_out[0])
        anchors = self.anchor_generator(images, [f.shape[1:3] for f in features])
        return objectness, reg, anchors

    def forward(self, images, features, targets=None, head_out=None):
        objectness, reg, anchors = self._head_and_anchors(images, features, head_out)
        if self.training or self.mode == "train":
            return self._forward_train(anchors, objectness, reg, targets)
        return self._forward_test(anchors, objectness, reg)

    def forward_teacher(self, images, features, targets=None, head_out=None):
        objectness, reg, anchors = self._head_and_anchors(images, features, head_out)
        with torch.no_grad():
            if self.cfg.MT.FFI:
                raise NotImplementedError("MT.FFI (comparison method) is out of scope")
            boxes = self.box_selector_train(anchors, objectness, reg, targets, head_out=self._head_out)
            obj, box_reg, sample_idx, label = self.loss_evaluator.teacher_sample_selection(
                anchors, objectness, reg, targets)
        return obj, box_reg, sample_idx, label, boxes, {}, None

    def _forward_train(self, anchors, objectness, reg, targets):
        if self.cfg.MODEL.RPN_ONLY:
            boxes = anchors
        else:
            with torch.no_grad():
                boxes = self.box_selector_train(anchors, objectness, reg, targets, head_out=self._head_out)
        loss_objectness, loss_rpn_box_reg = self.loss_evaluator(anchors, objectness, reg, targets, fused=self._head_out[0])
        return boxes, {"loss_objectness": loss_objectness, "loss_rpn_box_reg": loss_rpn_box_reg}

    def _forward_test(self, anchors, objectness, reg):
        boxes = self.box_selector_test(anchors, objectness, reg, head_out=self._head_out)
        if self.cfg.MODEL.RPN_ONLY:
            inds = [b.get_field("objectness").sort(descending=True)[1] for b in boxes]
            boxes = [b[i] for b, i in zip(boxes, inds)]
        return boxes, {}


def build_rpn(cfg, is_teacher=False):
    return RPNModule(cfg, is_teacher=is_teacher)
