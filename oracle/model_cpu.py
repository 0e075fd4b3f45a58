"""TEST INFRASTRUCTURE — CPU restatement (plain torch fp32, NCHW) of the reference's mean-teacher training step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this file.
It restates, with the reference's own tensor layouts and call order (so the CPU random stream is consumed
identically), the path  engine/MTtrainer.py:165-229,241-281 -> modeling/detector/generalized_rcnn.py:42-282 ->
backbone/{resnet,fpn}.py, rpn/{rpn,inference,loss}.py, roi_heads/box_head/*, roi_heads/mask_head/*, poolers.py,
solver/build.py, with MODEL.RELATION_* off and the eval-mode mask-logits bug fixed.
Weights are a dict in the reference's state-dict layout (OIHW convs, [out,in] linears, IOHW deconv).
Pinned by tests/golden/step_small.npz, produced by running the patched reference itself
(tests/golden/make_golden.py step).  ROIAlign = torchvision.ops.roi_align(aligned=False) (bit-identical to the
reference's CPU kernel); NMS = oracle.ops_ref.nms_ref (CUDA flavour, `>`)."""
import math
from bisect import bisect_right

import numpy as np
import torch
import torch.nn.functional as F
import torchvision.ops as tvops

from . import ops_ref, polygon_ref, relation_ref

BELOW, BETWEEN = -1, -2


# ------------------------------------------------------------------------------------------- helpers
def area(b):
    return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)


def box_iou(b1, b2):  # structures/boxlist_ops.py:57-92
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    return inter / (area(b1)[:, None] + area(b2) - inter)


def encode(ref, prop, w):  # modeling/box_coder.py:22-50
    ew, eh = prop[:, 2] - prop[:, 0] + 1, prop[:, 3] - prop[:, 1] + 1
    ex, ey = prop[:, 0] + 0.5 * ew, prop[:, 1] + 0.5 * eh
    gw, gh = ref[:, 2] - ref[:, 0] + 1, ref[:, 3] - ref[:, 1] + 1
    gx, gy = ref[:, 0] + 0.5 * gw, ref[:, 1] + 0.5 * gh
    return torch.stack((w[0] * (gx - ex) / ew, w[1] * (gy - ey) / eh, w[2] * torch.log(gw / ew), w[3] * torch.log(gh / eh)), 1)


def decode(codes, boxes, w, clip=math.log(1000. / 16)):  # modeling/box_coder.py:52-95
    bw, bh = boxes[:, 2] - boxes[:, 0] + 1, boxes[:, 3] - boxes[:, 1] + 1
    cx, cy = boxes[:, 0] + 0.5 * bw, boxes[:, 1] + 0.5 * bh
    dx, dy = codes[:, 0::4] / w[0], codes[:, 1::4] / w[1]
    dw = torch.clamp(codes[:, 2::4] / w[2], max=clip)
    dh = torch.clamp(codes[:, 3::4] / w[3], max=clip)
    px, py = dx * bw[:, None] + cx[:, None], dy * bh[:, None] + cy[:, None]
    pw, ph = torch.exp(dw) * bw[:, None], torch.exp(dh) * bh[:, None]
    out = torch.zeros_like(codes)
    out[:, 0::4], out[:, 1::4] = px - 0.5 * pw, py - 0.5 * ph
    out[:, 2::4], out[:, 3::4] = px + 0.5 * pw - 1, py + 0.5 * ph - 1
    return out


def match(iou, hi, lo, low_quality):  # modeling/matcher.py:44-138
    if iou.numel() == 0:
        raise ValueError("No %s boxes available for one of the images during training" % ("ground-truth" if iou.shape[0] == 0 else "proposal"))
    vals, m = iou.max(dim=0)
    allm = m.clone()
    below = vals < lo
    between = (vals >= lo) & (vals < hi)
    m[below], m[between] = BELOW, BETWEEN
    if low_quality:
        best, _ = iou.max(dim=1)
        upd = torch.nonzero(iou == best[:, None])[:, 1]
        m[upd] = allm[upd]
    return m


def sample(labels_list, per_image, pos_frac):  # modeling/balanced_positive_negative_sampler.py:16-71
    pos_out, neg_out = [], []
    for m in labels_list:
        pos = torch.nonzero(m >= 1).squeeze(1)
        neg = torch.nonzero(m == 0).squeeze(1)
        npos = min(pos.numel(), int(per_image * pos_frac))
        nneg = min(neg.numel(), per_image - npos)
        p1 = torch.randperm(pos.numel())[:npos]
        p2 = torch.randperm(neg.numel())[:nneg]
        pm, nm = torch.zeros_like(m, dtype=torch.bool), torch.zeros_like(m, dtype=torch.bool)
        pm[pos[p1]], nm[neg[p2]] = True, True
        pos_out.append(pm)
        neg_out.append(nm)
    return pos_out, neg_out


def smooth_l1(x, t, beta):  # layers/smooth_l1_loss.py:6-16 (size_average=False)
    n = torch.abs(x - t)
    return torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta).sum()


def cell_anchors(stride, size, ratios=(0.5, 1.0, 2.0)):  # modeling/rpn/anchor_generator.py:196-265
    ratios = np.array(ratios, dtype=float)
    scales = np.array([size], dtype=float) / stride

    def whc(a):
        w, h = a[2] - a[0] + 1, a[3] - a[1] + 1
        return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)

    def mk(ws, hs, cx, cy):
        ws, hs = ws[:, None], hs[:, None]
        return np.hstack((cx - 0.5 * (ws - 1), cy - 0.5 * (hs - 1), cx + 0.5 * (ws - 1), cy + 0.5 * (hs - 1)))
    w, h, cx, cy = whc(np.array([1, 1, stride, stride], dtype=float) - 1)
    ws = np.round(np.sqrt(w * h / ratios))
    per = mk(ws, np.round(ws * ratios), cx, cy)
    out = []
    for a in per:
        w, h, cx, cy = whc(a)
        out.append(mk(w * scales, h * scales, cx, cy))
    return torch.from_numpy(np.vstack(out)).float()


def grid_anchors(grid_sizes, strides, sizes):  # anchor_generator.py:79-103
    out = []
    for (gh, gw), s, sz in zip(grid_sizes, strides, sizes):
        base = cell_anchors(s, sz)
        sx = torch.arange(0, gw * s, step=s, dtype=torch.float32)
        sy = torch.arange(0, gh * s, step=s, dtype=torch.float32)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        xx, yy = xx.reshape(-1), yy.reshape(-1)
        sh = torch.stack((xx, yy, xx, yy), 1)
        out.append((sh.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
    return out


def level_map(boxes):  # modeling/poolers.py:31-42
    s = torch.sqrt(area(boxes))
    lv = torch.floor(4 + torch.log2(s / 224 + 1e-6))
    return torch.clamp(lv, min=2, max=5).to(torch.int64) - 2


class Cfg(object):
    """The cfg values of configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml + scripts/train_mt.sh that the path reads."""
    strides, sizes = (4, 8, 16, 32, 64), (32, 64, 128, 256, 512)
    pre_train, post_train, fpn_post_train = 2000, 2000, 2000
    pre_test, post_test, fpn_post_test = 1000, 1000, 1000
    rpn_nms, rpn_fg, rpn_bg, rpn_batch, rpn_pos = 0.7, 0.7, 0.3, 256, 0.5
    roi_fg, roi_bg, roi_batch, roi_pos = 0.5, 0.5, 512, 0.25
    box_w, score_thresh, det_nms, det_per_img = (10., 10., 5., 5.), 0.05, 0.5, 200
    scales = (0.25, 0.125, 0.0625, 0.03125)
    dropout, num_classes, mask_res = 0.5, 3, 28
    temp, neg_weight = 0.5, 1.5
    lam, cls_w, fg_w, alpha_max = 5.0, 0.2, 1.0, 0.99
    rampup, rampdown, start_mt, max_iter = 250, 250, 1000, 7000
    lr, bias_lr_factor, wd, wd_bias, momentum = 0.005, 2, 0.0001, 0, 0.9
    warmup_factor, warmup_iters, steps, gamma = 1.0 / 3, 500, (5000,), 0.1
    freeze_at = 2
    aug_k, aug_s = 2, 1
    relation = False          # IR-Net modules (BASELINE config 4): MODEL.RELATION_NMS.USE_RELATION_NMS + RELATION_MASK.USE_RELATION
    relation_cfg = relation_ref.RelationCfg
    nms_loss_w = 1.0          # MODEL.RELATION_NMS.LOSS (train_mt.sh)
    tf32_operands = False     # True: every conv / linear sees its operands with 10 mantissa bits (low 13 bits dropped), fp32
                              # accumulate — what the reference computes under torch.backends.cudnn.allow_tf32 (its GPU default)
                              # and what the tcgen05 kind::tf32 path computes; used to hold the SHIPPED conv math to a tight bound


class Net(object):
    """One GeneralizedRCNN replica over a reference-layout weight dict."""

    def __init__(self, sd, cfg, trainable):
        self.cfg = cfg
        self.sd = {k: v.clone().float() for k, v in sd.items()}
        self.param_names = [k for k in self.sd if not any(s in k for s in ("running_mean", "running_var", "cell_anchors"))
                            and not (k.split(".")[-2].startswith("bn") or ".downsample.1." in k)]
        frozen = ("backbone.body.stem.",) + tuple("backbone.body.layer%d." % i for i in range(1, cfg.freeze_at))
        self.trainable = [k for k in self.param_names if trainable and not k.startswith(frozen)]
        for k in self.trainable:
            self.sd[k].requires_grad_(True)

    # ---- layers ---------------------------------------------------------------------------------
    def _op(self, t):
        """Operand as the multiplier sees it (straight-through for autograd)."""
        if not self.cfg.tf32_operands:
            return t
        return t + (ops_ref.tf32_round(t.detach()) - t.detach())

    def conv(self, x, name, stride=1, pad=0):
        return F.conv2d(self._op(x), self._op(self.sd[name + ".weight"]), self.sd.get(name + ".bias"), stride, pad)

    def linear(self, x, name):
        return F.linear(self._op(x), self._op(self.sd[name + ".weight"]), self.sd[name + ".bias"])

    def bn(self, x, name):  # layers/batch_norm.py:19-24
        scale = self.sd[name + ".weight"] * self.sd[name + ".running_var"].rsqrt()
        bias = self.sd[name + ".bias"] - self.sd[name + ".running_mean"] * scale
        return x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)

    def backbone(self, x):  # resnet.py:117-124,254-293 ; fpn.py:43-74
        p = "backbone.body."
        x = F.relu(self.bn(self.conv(x, p + "stem.conv1", 2, 3), p + "stem.bn1"))
        x = F.max_pool2d(x, 3, 2, 1)
        feats = []
        for li, nb in ((1, 3), (2, 4), (3, 6), (4, 3)):
            for b in range(nb):
                q = "%slayer%d.%d." % (p, li, b)
                s = 2 if (b == 0 and li > 1) else 1
                idt = x
                if (q + "downsample.0.weight") in self.sd:
                    idt = self.bn(self.conv(x, q + "downsample.0", s), q + "downsample.1")
                o = F.relu(self.bn(self.conv(x, q + "conv1", s), q + "bn1"))
                o = F.relu(self.bn(self.conv(o, q + "conv2", 1, 1), q + "bn2"))
                x = F.relu(self.bn(self.conv(o, q + "conv3"), q + "bn3") + idt)
            feats.append(x)
        f = "backbone.fpn."
        last = self.conv(feats[3], f + "fpn_inner4")
        res = [self.conv(last, f + "fpn_layer4", 1, 1)]
        for i in (3, 2, 1):
            last = self.conv(feats[i - 1], f + "fpn_inner%d" % i) + F.interpolate(last, scale_factor=2, mode="nearest")
            res.insert(0, self.conv(last, f + "fpn_layer%d" % i, 1, 1))
        res.append(F.max_pool2d(res[-1], 1, 2, 0))
        return tuple(res)

    def rpn_head(self, feats):  # rpn.py:39-46
        obj, reg = [], []
        for f in feats:
            t = F.relu(self.conv(f, "rpn.head.conv", 1, 1))
            obj.append(self.conv(t, "rpn.head.cls_logits"))
            reg.append(self.conv(t, "rpn.head.bbox_pred"))
        return obj, reg

    def pool(self, feats, boxes_per_img, res):  # poolers.py:91-121
        boxes = torch.cat(boxes_per_img, 0)
        ids = torch.cat([torch.full((len(b), 1), float(i)) for i, b in enumerate(boxes_per_img)], 0)
        rois = torch.cat([ids, boxes], 1)
        lv = level_map(boxes)
        out = torch.zeros((len(rois), feats[0].shape[1], res, res))
        for l, sc in enumerate(self.cfg.scales):
            idx = torch.nonzero(lv == l).squeeze(1)
            out[idx] = tvops.roi_align(feats[l], rois[idx], (res, res), sc, 2, aligned=False)
        return out

    def box_feat(self, feats, boxes_per_img, istrain):  # roi_box_feature_extractors.py:107-125
        p = "box_heads.box.feature_extractor."
        x = self.pool(feats, boxes_per_img, 7).flatten(1)
        x = F.relu(self.linear(x, p + "fc6"))
        x = F.relu(self.linear(x, p + "fc7"))
        if self.cfg.dropout > 0:
            x = F.dropout(x, p=self.cfg.dropout, training=istrain)
        return x

    def box_pred(self, x):
        p = "box_heads.box.predictor."
        return self.linear(x, p + "cls_score"), self.linear(x, p + "bbox_pred")

    def mask_logits(self, feats, boxes_per_img, with_features=False):  # roi_mask_feature_extractors.py:131-146, roi_mask_predictors.py:34-36
        x = self.pool(feats, boxes_per_img, 14)
        for i in range(1, 5):
            x = F.relu(self.conv(x, "mask_heads.mask.feature_extractor.mask_fcn%d" % i, 1, 1))
        p = "mask_heads.mask.predictor."
        y = F.relu(F.conv_transpose2d(self._op(x), self._op(self.sd[p + "conv5_mask.weight"]), self.sd[p + "conv5_mask.bias"], 2))
        y = self.conv(y, p + "mask_fcn_logits")
        return (x, y) if with_features else y

    def refine_masks(self, x, ml, boxes, labels, obj):  # mask_head.py:96-124 (IR-Net branch, per image)
        """Per-image CIAM refinement; returns (refined logits, boxes, labels) with every image's rows re-ordered by
        (class, objectness) as the reference's sorted proposals are."""
        n = [len(b) for b in boxes]
        out, nb, nl = [], [], []
        for xi, mi, b, l, o in zip(x.split(n), ml.split(n), boxes, labels, obj):
            if len(b) == 0:
                out.append(mi); nb.append(b); nl.append(l)
                continue
            lg, order = relation_ref.mask_relation(self.sd, relation_ref.MASK_PREFIX, xi, mi, l, o, self.cfg.num_classes - 1)
            out.append(lg); nb.append(b[order]); nl.append(l[order])
        return torch.cat(out), nb, nl

    def seg_loss(self, ml, posb, targets):  # mask_head/loss.py:118-180
        tl, tm = [], []
        for pb, t in zip(posb, targets):
            m = match(box_iou(t["bbox"], pb), self.cfg.roi_fg, self.cfg.roi_bg, False)
            lab = t["labels"][m.clamp(min=0)].to(torch.int64)
            lab[m == BELOW] = 0
            pi = torch.nonzero(lab > 0).squeeze(1)
            inst = [t["polygons"][int(i)] for i in m[pi]]
            tm.append(torch.from_numpy(polygon_ref.project_masks_on_boxes_ref(inst, pb[pi].numpy(), self.cfg.mask_res)))
            tl.append(lab)
        tl, tm = torch.cat(tl), torch.cat(tm)
        pi = torch.nonzero(tl > 0).squeeze(1)
        return F.binary_cross_entropy_with_logits(ml[pi, tl[pi]], tm) if tm.numel() else ml.sum() * 0

    def adaptors(self, feats):
        return [self.conv(feats[i], "hint_adaptor.adapter_%d" % (i + 1)) for i in range(5)]

    # ---- RPN -------------------------------------------------------------------------------------
    def anchors(self, feats, image_sizes):
        per_level = grid_anchors([f.shape[-2:] for f in feats], self.cfg.strides, self.cfg.sizes)
        vis = []
        for (h, w) in image_sizes:
            a = torch.cat(per_level, 0)
            vis.append((a[:, 0] >= 0) & (a[:, 1] >= 0) & (a[:, 2] < w) & (a[:, 3] < h))
        return per_level, vis

    def proposals(self, per_level_anchors, obj, reg, image_sizes, pre, post, fpn_post, batch_topk, gt_boxes=None):
        """rpn/inference.py:78-172,216-244.  Returns list of (boxes, objectness) per image."""
        c = self.cfg
        N = obj[0].shape[0]
        per_img = [[] for _ in range(N)]
        for a, o, r in zip(per_level_anchors, obj, reg):
            _, A, H, W = o.shape
            o = o.permute(0, 2, 3, 1).reshape(N, -1).sigmoid()
            r = r.view(N, -1, 4, H, W).permute(0, 3, 4, 1, 2).reshape(N, -1, 4)
            k = min(pre, A * H * W)
            o, idx = o.topk(k, dim=1, sorted=True)
            bi = torch.arange(N)[:, None]
            r = r[bi, idx]
            an = a[None].expand(N, -1, -1)[bi, idx]
            props = decode(r.reshape(-1, 4), an.reshape(-1, 4), (1., 1., 1., 1.)).view(N, -1, 4)
            for j, (h, w) in enumerate(image_sizes):
                b = props[j].clone()
                b[:, 0].clamp_(min=0, max=w - 1); b[:, 1].clamp_(min=0, max=h - 1)
                b[:, 2].clamp_(min=0, max=w - 1); b[:, 3].clamp_(min=0, max=h - 1)
                keep = torch.from_numpy(ops_ref.nms_ref(b.detach().numpy(), o[j].detach().numpy(), c.rpn_nms))[:post]
                per_img[j].append((b[keep], o[j][keep]))
        out = [(torch.cat([b for b, _ in lv], 0), torch.cat([s for _, s in lv], 0)) for lv in per_img]
        if batch_topk:
            allo = torch.cat([s for _, s in out], 0)
            k = min(fpn_post, len(allo))
            _, inds = torch.topk(allo, k, dim=0, sorted=True)
            mask = torch.zeros_like(allo, dtype=torch.bool)
            mask[inds] = True
            mask = mask.split([len(s) for _, s in out])
            out = [(b[m], s[m]) for (b, s), m in zip(out, mask)]
        else:
            new = []
            for b, s in out:
                _, inds = torch.topk(s, min(fpn_post, len(s)), dim=0, sorted=True)
                new.append((b[inds], s[inds]))
            out = new
        if gt_boxes is not None:
            out = [(torch.cat([b, g], 0), torch.cat([s, torch.ones(len(g))], 0)) for (b, s), g in zip(out, gt_boxes)]
        return out

    def rpn_targets(self, per_level_anchors, vis, gt_boxes):
        a = torch.cat(per_level_anchors, 0)
        labels, regt = [], []
        for v, g in zip(vis, gt_boxes):
            m = match(box_iou(g, a), self.cfg.rpn_fg, self.cfg.rpn_bg, True)
            lab = (m >= 0).float()
            lab[~v] = -1
            lab[m == BETWEEN] = -1
            labels.append(lab)
            regt.append(encode(g[m.clamp(min=0)], a, (1., 1., 1., 1.)))
        return labels, regt

    def rpn_loss(self, per_level_anchors, vis, obj, reg, gt_boxes):  # rpn/loss.py:138-196
        labels, regt = self.rpn_targets(per_level_anchors, vis, gt_boxes)
        pos, neg = sample(labels, self.cfg.rpn_batch, self.cfg.rpn_pos)
        pos = torch.nonzero(torch.cat(pos)).squeeze(1)
        neg = torch.nonzero(torch.cat(neg)).squeeze(1)
        sampled = torch.cat([pos, neg])
        N = obj[0].shape[0]
        of = torch.cat([o.permute(0, 2, 3, 1).reshape(N, -1) for o in obj], 1).reshape(-1)
        rf = torch.cat([r.view(N, -1, 4, r.shape[2], r.shape[3]).permute(0, 3, 4, 1, 2).reshape(N, -1, 4) for r in reg], 1).reshape(-1, 4)
        labels, regt = torch.cat(labels), torch.cat(regt)
        box_loss = smooth_l1(rf[pos], regt[pos], 1.0 / 9) / sampled.numel()
        obj_loss = F.binary_cross_entropy_with_logits(of[sampled], labels[sampled])
        return obj_loss, box_loss

    # ---- box head ----------------------------------------------------------------------------------
    def subsample(self, props, gt_boxes, gt_labels):  # box_head/loss.py:48-116
        labels, regt = [], []
        for (b, _), g, gl in zip(props, gt_boxes, gt_labels):
            m = match(box_iou(g, b), self.cfg.roi_fg, self.cfg.roi_bg, False)
            lab = gl[m.clamp(min=0)].to(torch.int64)
            lab[m == BELOW] = 0
            lab[m == BETWEEN] = -1
            labels.append(lab)
            regt.append(encode(g[m.clamp(min=0)], b, self.cfg.box_w))
        pos, neg = sample(labels, self.cfg.roi_batch, self.cfg.roi_pos)
        out = []
        for (b, _), lab, rt, p, n in zip(props, labels, regt, pos, neg):   # `_` = RPN objectness (1 for appended GT boxes)
            idx = torch.nonzero(p | n).squeeze(1)
            out.append(dict(bbox=b[idx], labels=lab[idx], regt=rt[idx], obj=_[idx]))
        return out

    def detections(self, cls, reg, props, image_sizes):  # box_head/inference.py:36-145
        c = self.cfg
        prob = F.softmax(cls, -1)
        n = [len(b) for b, _ in props]
        dec = decode(reg, torch.cat([b for b, _ in props], 0), c.box_w).split(n)
        prob = prob.split(n)
        out = []
        for pr, bx, (h, w) in zip(prob, dec, image_sizes):
            bx = bx.reshape(-1, 4).clone()
            bx[:, 0].clamp_(min=0, max=w - 1); bx[:, 1].clamp_(min=0, max=h - 1)
            bx[:, 2].clamp_(min=0, max=w - 1); bx[:, 3].clamp_(min=0, max=h - 1)
            bx = bx.reshape(-1, c.num_classes * 4)
            rb, rs, rl = [], [], []
            for j in range(1, c.num_classes):
                inds = (pr[:, j] > c.score_thresh).nonzero().squeeze(1)
                sj, bj = pr[inds, j], bx[inds, j * 4:(j + 1) * 4]
                keep = torch.from_numpy(ops_ref.nms_ref(bj.numpy(), sj.numpy(), c.det_nms))
                rb.append(bj[keep]); rs.append(sj[keep]); rl.append(torch.full((len(keep),), j, dtype=torch.int64))
            rb, rs, rl = torch.cat(rb), torch.cat(rs), torch.cat(rl)
            if len(rs) > c.det_per_img > 0:
                th, _ = torch.kthvalue(rs, len(rs) - c.det_per_img + 1)
                k = torch.nonzero(rs >= th.item()).squeeze(1)
                rb, rs, rl = rb[k], rs[k], rl[k]
            out.append(dict(bbox=rb, scores=rs, labels=rl))
        return out

    # ---- supervised forward (generalized_rcnn.py:42-115, training) ----------------------------------
    def forward_train(self, images, targets):
        image_sizes = [tuple(images.shape[-2:])] * images.shape[0]
        feats = self.backbone(images)
        obj, reg = self.rpn_head(feats)
        pla, vis = self.anchors(feats, image_sizes)
        gtb = [t["bbox"] for t in targets]
        with torch.no_grad():
            props = self.proposals(pla, obj, reg, image_sizes, self.cfg.pre_train, self.cfg.post_train,
                                   self.cfg.fpn_post_train, True, gtb)
        l_obj, l_rpn = self.rpn_loss(pla, vis, obj, reg, gtb)
        with torch.no_grad():
            samp = self.subsample(props, gtb, [t["labels"] for t in targets])
        x = self.box_feat(feats, [s["bbox"] for s in samp], True)
        cls, br = self.box_pred(x)
        labels = torch.cat([s["labels"] for s in samp])
        regt = torch.cat([s["regt"] for s in samp])
        l_cls = F.cross_entropy(cls, labels)
        pos = torch.nonzero(labels > 0).squeeze(1)
        mi = 4 * labels[pos][:, None] + torch.tensor([0, 1, 2, 3])
        l_box = smooth_l1(br[pos[:, None], mi], regt[pos], 1.0) / labels.numel()
        out = {}
        if self.cfg.relation:  # generalized_rcnn.py:63-95: IR-Net duplicate removal, one image per call
            n = [len(s["bbox"]) for s in samp]
            prob = F.softmax(cls, 1)
            per = [relation_ref.relation_nms_loss(self.sd, relation_ref.NMS_PREFIX, self.cfg.relation_cfg, xi, pi_, ri, s["bbox"],
                                                  s["obj"], (isz[1], isz[0]), t["bbox"], t["labels"])[0]
                   for xi, pi_, ri, s, isz, t in zip(x.split(n), prob.split(n), br.split(n), samp, image_sizes, targets)]
            out["nms_loss"] = torch.mean(torch.stack(per))
        # mask head on positives (mask_head.py:74-150, loss.py:118-180)
        posb = [s["bbox"][s["labels"] > 0] for s in samp]
        mx, ml = self.mask_logits(feats, posb, with_features=True)
        l_seg = self.seg_loss(ml, posb, targets)
        if self.cfg.relation:  # mask_head.py:96-150, DEEP_SUPER
            ml2, sb, _ = self.refine_masks(mx, ml, posb, [s["labels"][s["labels"] > 0] for s in samp],
                                           [s["obj"][s["labels"] > 0] for s in samp])
            l_seg = 0.5 * (l_seg + self.seg_loss(ml2, sb, targets))
        out.update(loss_classifier=l_cls, loss_box_reg=l_box, loss_seg=l_seg, loss_objectness=l_obj, loss_rpn_box_reg=l_rpn)
        return out

    # ---- teacher (generalized_rcnn.py:117-167) --------------------------------------------------------
    @torch.no_grad()
    def forward_teacher(self, views):
        c = self.cfg
        image_sizes = [tuple(views[0].shape[-2:])] * views[0].shape[0]
        H, W = image_sizes[0]
        f0 = self.backbone(views[0])
        obj, reg = self.rpn_head(f0)
        pla, vis = self.anchors(f0, image_sizes)
        props = self.proposals(pla, obj, reg, image_sizes, c.pre_test, c.post_test, c.fpn_post_test, False)
        x = self.box_feat(f0, [b for b, _ in props], False)
        cls, br = self.box_pred(x)
        if c.relation:  # the box head hands its raw ROIs to the duplicate-removal network (box_head.py:55-58)
            n = [len(b) for b, _ in props]
            prob = F.softmax(cls, 1)
            dets = [relation_ref.relation_nms_detect(self.sd, relation_ref.NMS_PREFIX, c.relation_cfg, xi, pi_, ri, b, o, (sz[1], sz[0]))
                    for xi, pi_, ri, (b, o), sz in zip(x.split(n), prob.split(n), br.split(n), props, image_sizes)]
            mx, ml = self.mask_logits(f0, [d["bbox"] for d in dets], with_features=True)
            ml, sb, sl = self.refine_masks(mx, ml, [d["bbox"] for d in dets], [d["labels"] for d in dets],
                                           [d["objectness"] for d in dets])
            dets = [dict(bbox=b, labels=l) for b, l in zip(sb, sl)]
            ml = ml.sigmoid()
        else:
            dets = self.detections(cls, br, props, image_sizes)
            ml = self.mask_logits(f0, [d["bbox"] for d in dets]).sigmoid()
        labels = torch.cat([d["labels"] for d in dets])
        prob = ml[torch.arange(len(labels)), labels].split([len(d["bbox"]) for d in dets])
        integral = torch.stack([paste_sum(p, d["bbox"], H, W) for p, d in zip(prob, dets)])
        aug = []
        for v in views:
            aug.append(self.backbone(v))
            aug.append(self.backbone(torch.flip(v, (3,))))
        # rpn.forward_teacher on view 0 (train selector, module in eval mode -> per-image top-k, no GT)
        tprops = self.proposals(pla, obj, reg, image_sizes, c.pre_train, c.post_train, c.fpn_post_train, False)
        labels_a, _ = self.rpn_targets(pla, vis, [d["bbox"] for d in dets])
        sample(labels_a, c.rpn_batch, c.rpn_pos)  # teacher_sample_selection: result unused, RNG consumed
        emb = [self.adaptors(f) for f in aug]
        samp = self.subsample(tprops, [d["bbox"] for d in dets], [d["labels"] for d in dets])
        boxes = [s["bbox"] for s in samp]
        flipped = []
        for b in boxes:
            fb = b.clone()
            fb[:, 0], fb[:, 2] = W - b[:, 2] - 1, W - b[:, 0] - 1
            flipped.append(fb)
        logits = [self.box_pred(self.box_feat(f, boxes if i % 2 == 0 else flipped, False))[0] for i, f in enumerate(aug)]
        return dict(samples=samp, class_logit_t=logits, embedding=emb, seg_mask=integral, dets=dets)

    # ---- student on the unlabeled view (generalized_rcnn.py:170-199) -----------------------------------
    def forward_student(self, image, t):
        feats = self.backbone(image)
        out = {}
        student = self.adaptors(feats)
        out["mt_fg_loss"], _ = ops_ref.mgd_loss_ref(t["embedding"], student, t["seg_mask"])
        x = self.box_feat(feats, [s["bbox"] for s in t["samples"]], True)
        logits, _ = self.box_pred(x)
        labels = torch.cat([s["labels"] for s in t["samples"]])
        out["mt_classifier"], _ = ops_ref.psm_loss_ref(logits, t["class_logit_t"], labels, self.cfg.temp, True, True,
                                                       self.cfg.neg_weight)
        return out


def paste_sum(probs, boxes, im_h, im_w, thresh=0.5, padding=1):
    """mask_head/inference.py:169-231 per detection + `.sum(0)` (generalized_rcnn.py:129-132)."""
    out = torch.zeros((im_h, im_w))
    for mask, box in zip(probs, boxes):
        M = mask.shape[-1]
        scale = float(M + 2 * padding) / M
        padded = mask.new_zeros((M + 2 * padding, M + 2 * padding))
        padded[padding:-padding, padding:-padding] = mask
        wh, hh = (box[2] - box[0]) * .5, (box[3] - box[1]) * .5
        xc, yc = (box[2] + box[0]) * .5, (box[3] + box[1]) * .5
        wh, hh = wh * scale, hh * scale
        b = torch.stack([xc - wh, yc - hh, xc + wh, yc + hh]).to(torch.int32)
        w, h = max(int(b[2] - b[0] + 1), 1), max(int(b[3] - b[1] + 1), 1)
        m = F.interpolate(padded[None, None], size=(h, w), mode="bilinear", align_corners=False)[0, 0] > thresh
        x0, x1 = max(int(b[0]), 0), min(int(b[2]) + 1, im_w)
        y0, y1 = max(int(b[1]), 0), min(int(b[3]) + 1, im_h)
        if x1 > x0 and y1 > y0:
            out[y0:y1, x0:x1] += m[(y0 - int(b[1])):(y1 - int(b[1])), (x0 - int(b[0])):(x1 - int(b[0]))].float()
    return out


class Trainer(object):
    """engine/MTtrainer.py:165-229 + solver/build.py:5-23 + lr_scheduler.py:41-53."""

    def __init__(self, sd, cfg=None):
        self.cfg = cfg or Cfg()
        self.student = Net(sd, self.cfg, True)
        self.teacher = Net(sd, self.cfg, False)
        self.mom = {}
        self.last_epoch = 0

    def lr(self):
        c, e = self.cfg, self.last_epoch
        f = 1
        if e < c.warmup_iters:
            a = e / c.warmup_iters
            f = c.warmup_factor * (1 - a) + a
        return c.lr * f * c.gamma ** bisect_right(list(c.steps), e)

    def mt_weight(self, step):
        c = self.cfg
        if 0 < (step - c.start_mt) < c.rampup:
            return c.lam * float(np.exp(-5.0 * (1 - np.clip(step - c.start_mt, 0.0, c.rampup) / c.rampup) ** 2))
        if (c.max_iter - step) < c.rampdown:
            return c.lam * float(np.exp(-12 * (1 - (c.max_iter - step) / c.rampup) ** 2))
        return c.lam

    def step(self, iteration, images, targets, views=None):
        c = self.cfg
        losses = self.student.forward_train(images, targets)
        if iteration > c.start_mt and c.lam > 0 and views is not None:
            t = self.teacher.forward_teacher(views[:c.aug_k])
            losses.update(self.student.forward_student(views[-1], t))
        self.last_epoch += 1
        w = self.mt_weight(iteration)
        bal = {"mt_classifier": c.cls_w, "mt_fg_loss": c.fg_w, "nms_loss": c.nms_loss_w}
        weighted = {k: (v * w if "mt" in k else v) * bal.get(k, 1.0) for k, v in losses.items()}
        total = sum(weighted.values())
        for k in self.student.trainable:
            self.student.sd[k].grad = None
        total.backward()
        lr = self.lr()
        with torch.no_grad():
            for k in self.student.trainable:
                p = self.student.sd[k]
                if p.grad is None:
                    continue
                plr, wd = (lr * c.bias_lr_factor, c.wd_bias) if "bias" in k else (lr, c.wd)
                d = p.grad + wd * p if wd != 0 else p.grad
                self.mom[k] = d.clone() if k not in self.mom else self.mom[k] * c.momentum + d
                p -= plr * self.mom[k]
            if c.lam > 0 and iteration > (c.start_mt - 10):
                a = min(1 - 1 / (iteration - (c.start_mt - 10) + 1), c.alpha_max)
                for k in self.student.param_names:
                    self.teacher.sd[k].mul_(a).add_(self.student.sd[k].detach(), alpha=1 - a)
        return {k: v.detach() for k, v in weighted.items()}, total.detach()
