"""TEST INFRASTRUCTURE — CPU restatement (plain torch fp32, reference layouts) of the reference's IR-Net modules.

Only tests/ (and oracle/model_cpu.py, itself test infrastructure) import this file; the product never does.
Follows  modeling/relation/relation_module.py  (RelationModule :13-91, DuplicationRemovalNetwork :94-331,
prepare_reg_label :333-413, prepare_label :415-473, prepare_ranking / filter_results :475-593,
extract_rank_embedding :607-627, extract_multi_position_matrix :629-682)  and
modeling/relation/mask_relation_module.py  (MaskRelationRefineNet :16-156, CIAM_Module :190-242) with
relation_mask_feature_extractor.py:10-47, for the configuration configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml sets:
    RELATION_NMS:  FIRST_N 90, THREAD (0.1,), GROUP 16, HID_DIM (1024, 1024, 128), GEO_FEAT_DIM 64, TOPK 40,
                   CLASS_AGNOSTIC False, USE_IOU False, FG_THREAD 0.1, MERGE_METHOD 0, POS_NMS 0.55, REG_IOU (train_mt.sh)
    RELATION_MASK: TYPE CIAM, FEATURE_EXTRACTOR RoiAlignMaskFeatureExtractor, EXTRACTOR_CHANNEL 16, SAME_PREDICTOR False,
                   NORM -1, PRE_NORM False, DEEP_SUPER True
Pinned by tests/golden/relation.npz (the patched reference's modules executed on seeded inputs; make_golden.py relation).
Weights: a dict in the reference's state-dict layout under a key prefix."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ops_ref


class RelationCfg(object):
    first_n, thresholds, group, hid = 90, (0.1,), 16, (1024, 1024, 128)
    geo_dim, topk, fg_thread, pos_nms, nuclei_nms, merge = 64, 40, 0.1, 0.55, 0.5, 0
    roi_feat_dim, feat_dim, det_per_img = 1024, 128, 200
    box_w = (10., 10., 5., 5.)
    reg_iou = True


# ------------------------------------------------------------------------------------------- embeddings
def rank_embedding(n, feat_dim, wave=1000.0):  # relation_module.py:607-627
    r = torch.arange(0, n).float().view(-1, 1)
    k = torch.arange(feat_dim / 2)
    m = r * (1.0 / torch.pow(wave, k / (feat_dim / 2))).view(1, -1)
    return torch.cat((torch.sin(m), torch.cos(m)), -1)


def position_embedding(boxes, dim_g=64, wave=1000.0):  # relation_module.py:629-682 (iou None, clswise False)
    """boxes [N, C, 4] -> [C, N, N, dim_g]: sin/cos code of (log|dx/w|, log|dy/h|, log(w_i/w_j), log(h_i/h_j))."""
    b = boxes.permute(1, 0, 2)
    x0, y0, x1, y1 = torch.chunk(b, 4, dim=2)
    cx, cy = (x0 + x1) * 0.5, (y0 + y1) * 0.5
    w, h = (x1 - x0) + 1.0, (y1 - y0) + 1.0
    T = lambda t: t.permute(0, 2, 1)
    dx = torch.log(torch.clamp(torch.abs((cx - T(cx)) / w), min=1e-3))
    dy = torch.log(torch.clamp(torch.abs((cy - T(cy)) / h), min=1e-3))
    dw = torch.log(w / T(w))
    dh = torch.log(h / T(h))
    pos = torch.stack((dx, dy, dw, dh), -1)                       # [C,N,N,4]
    k = torch.arange(dim_g / 8)
    freq = (1.0 / torch.pow(wave, k / (dim_g / 8))).view(1, 1, 1, 1, -1)
    m = (100.0 * pos.unsqueeze(-1)) * freq                        # [C,N,N,4,dim_g/8]
    m = m.reshape(pos.shape[0], pos.shape[1], pos.shape[2], -1)
    return torch.cat((torch.sin(m), torch.cos(m)), -1)


# ------------------------------------------------------------------------------------------- label preparation
def _iou_plus1(a, b):
    area_a = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
    area_b = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    lt = torch.max(a[:, None, :2], b[:, :2])
    rb = torch.min(a[:, None, 2:], b[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    return inter / (area_a[:, None] + area_b - inter)


def reg_labels(sorted_boxes, sorted_scores, gt_boxes, gt_labels, thresholds, binary=False):
    """relation_module.py:333-413 (IoU regression targets) / :415-473 (binary=True: keep / duplicate labels).
    sorted_boxes [N, C, 4], sorted_scores [N, C] -> float32 [N, C, len(thresholds)].
    Per class and threshold: every ground-truth box picks, among the ranked boxes whose best-overlapping ground truth it
    is and whose IoU with it exceeds the threshold, the one with the highest class score; that box's target is the IoU
    (or 1), all others 0.  A ground truth without any candidate falls on rank 0 (argmax of an all-zero column), and
    np.intersect1d's first-occurrence index decides which ground truth's value rank 0 then receives."""
    N, C = sorted_scores.shape
    out = np.zeros((N, C, len(thresholds)), dtype=np.float32)
    for c in range(C):
        gt = gt_boxes[gt_labels == c + 1]
        M = gt.shape[0]
        if M == 0:
            continue
        iou = _iou_plus1(sorted_boxes[:, c, :].reshape(-1, 4), gt).numpy()
        score = sorted_scores[:, c:c + 1].numpy()
        onehot = np.eye(M)[np.argmax(iou, axis=1)]
        for t, thr in enumerate(thresholds):
            over = iou > thr
            cand = np.tile(score, (1, M)) * over * onehot
            winner = np.argmax(cand, axis=0)                                  # [M] rank index per ground truth
            value = (iou * over * onehot)[winner, np.arange(M)]
            rows, first, _ = np.intersect1d(winner, np.where(over)[0], return_indices=True)
            out[rows, c, t] = 1.0 if binary else value[first]
    return out


# ------------------------------------------------------------------------------------------- duplicate removal
def rank_rois(cfg, prob, reg, rois, obj, image_size, gt=None):  # relation_module.py:475-593
    """prob [n, 1+C] softmax scores, reg [n, 4(1+C)], rois [n, 4], obj [n]; image_size (w, h).
    Per foreground class the first_n ROIs by class score: dict(idx [N,C], scores [N,C], all_scores [N,C,C],
    obj [N,C], boxes [N,C,4] clipped, labels [N,C,T] when gt=(boxes, labels))."""
    C = prob.shape[1] - 1
    dec = ops_ref_decode(reg, rois, cfg.box_w)
    per_cls = torch.stack([dec[:, 4 * j:4 * j + 4] for j in range(1, C + 1)], 2)       # [n,4,C]
    scores = prob[:, 1:]
    N = min(rois.shape[0], cfg.first_n)
    top, idx = torch.topk(scores, N, dim=0, largest=True, sorted=True)
    picked = per_cls[idx]                                                              # [N,C,4,C]
    boxes = torch.stack([picked[:, c, :, c] for c in range(C)], 1)                     # class c's own regression
    out = dict(idx=idx, scores=top, all_scores=scores[idx], obj=obj[idx])
    if gt is not None:
        out["labels_cls"] = torch.from_numpy(reg_labels(boxes, top, gt[0], gt[1], cfg.thresholds, binary=True))
        out["labels_reg"] = torch.from_numpy(reg_labels(boxes, top, gt[0], gt[1], cfg.thresholds))
    w, h = image_size
    boxes = boxes.clone()
    boxes[..., 0].clamp_(min=0, max=w - 1); boxes[..., 1].clamp_(min=0, max=h - 1)
    boxes[..., 2].clamp_(min=0, max=w - 1); boxes[..., 3].clamp_(min=0, max=h - 1)
    out["boxes"] = boxes
    return out


def ops_ref_decode(codes, boxes, w, clip=math.log(1000. / 16)):  # modeling/box_coder.py:52-95
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


def relation_attention(sd, p, cfg, f_a, pos):  # relation_module.py:31-91 (iou None)
    """f_a [N, C, D], pos [C, N, N, geo] -> [N, C, hid[2]]: 16-group attention whose logits are
    log(max(relu(WG pos), 1e-6)) + <WQ f_i, WK f_j>/sqrt(d), softmax over each row's top-k, values = f_a itself,
    output through a grouped 1x1 conv (one [D -> hid[2]/group] map per group)."""
    N, C, D = f_a.shape
    G = cfg.group
    dq, dk = cfg.hid[0] // G, cfg.hid[1] // G
    f = f_a.permute(1, 0, 2)                                                      # [C,N,D]
    flat = f.contiguous().view(C * N, D)
    w_g = F.relu(F.linear(pos.view(-1, pos.shape[-1]), sd[p + "WG.weight"], sd[p + "WG.bias"]))
    k = F.linear(flat, sd[p + "WK.weight"], sd[p + "WK.bias"]).view(C, N, G, dk).permute(0, 2, 3, 1).reshape(C * G, dk, N)
    q = F.linear(flat, sd[p + "WQ.weight"], sd[p + "WQ.bias"]).view(C, N, G, dq).transpose(1, 2).reshape(C * G, N, dq)
    aff = (1.0 / math.sqrt(float(dk))) * torch.bmm(q, k)
    w_g = w_g.view(C, N, N, G).permute(0, 3, 1, 2).reshape(C * G, N, N)
    logit = torch.log(torch.clamp(w_g, min=1e-6)) + aff
    top, ind = torch.topk(logit, min(N, cfg.topk), dim=2, largest=True, sorted=True)
    att = torch.zeros_like(logit).scatter(2, ind, F.softmax(top, dim=2)).view(C, G * N, N)
    o = torch.bmm(att, f).view(C, G, N, D).permute(1, 3, 2, 0).reshape(1, G * D, N, C)
    o = F.conv2d(o, sd[p + "conv1.weight"], sd[p + "conv1.bias"], groups=G)       # [1,hid2,N,C]
    return o.squeeze(0).permute(1, 2, 0)


def relation_scores(sd, p, cfg, x, ranked):  # relation_module.py:145-210
    """x [n, roi_feat_dim] box-head features of one image -> [N, C, T] regressed IoU (REG_IOU) per ranked box."""
    emb = F.linear(x, sd[p + "roi_feat_embedding_fc.weight"], sd[p + "roi_feat_embedding_fc.bias"])
    feat = emb[ranked["idx"]]                                                     # [N,C,D]
    N = feat.shape[0]
    rank = F.linear(rank_embedding(N, cfg.roi_feat_dim), sd[p + "nms_rank_fc.weight"], sd[p + "nms_rank_fc.bias"])
    feat = feat + rank[:, None, :]
    pos = position_embedding(ranked["boxes"], cfg.geo_dim)
    feat = F.relu(feat + relation_attention(sd, p + "relation_module.", cfg, feat, pos))
    out = F.linear(feat.view(-1, cfg.feat_dim), sd[p + "classifier.weight"], sd[p + "classifier.bias"])
    out = out.view(N, -1, len(cfg.thresholds))
    return out if cfg.reg_iou else torch.sigmoid(out)


def relation_nms_loss(sd, p, cfg, x, prob, reg, rois, obj, image_size, gt_boxes, gt_labels):  # :211-236
    with torch.no_grad():
        ranked = rank_rois(cfg, prob, reg, rois, obj, image_size, (gt_boxes, gt_labels))
    out = relation_scores(sd, p, cfg, x, ranked)
    if cfg.reg_iou:
        return F.mse_loss(ranked["labels_reg"], out), ranked
    s = ranked["scores"][:, :, None]
    return F.binary_cross_entropy(s * (s * out), ranked["labels_cls"]), ranked


@torch.no_grad()
def relation_nms_detect(sd, p, cfg, x, prob, reg, rois, obj, image_size):  # :237-331
    """Eval branch: dict(bbox, scores, labels, objectness, all_scores) of one image, nuclei (label 2) first."""
    ranked = rank_rois(cfg, prob, reg, rois, obj, image_size)
    out = relation_scores(sd, p, cfg, x, ranked)
    s_in = ranked["scores"][:, :, None]
    score = out * (s_in > cfg.fg_thread).float() if cfg.reg_iou else out * s_in
    score = score[:, :, min(max(cfg.merge, 0), len(cfg.thresholds))]              # MERGE_METHOD >= 0
    parts = []
    for c, thr in ((1, cfg.nuclei_nms), (0, cfg.pos_nms)):
        sel = (score[:, c] >= cfg.fg_thread).nonzero()[:, 0]
        b, s = ranked["boxes"][sel, c], score[sel, c]
        keep = torch.from_numpy(ops_ref.nms_ref(b.numpy(), s.numpy(), thr)) if thr > 0 else torch.arange(len(sel))
        sel = sel[keep]
        parts.append(dict(bbox=ranked["boxes"][sel, c], scores=score[sel, c], objectness=ranked["obj"][sel, c],
                          all_scores=ranked["all_scores"][sel, c], labels=torch.full((len(sel),), c + 1, dtype=torch.int64)))
    det = {k: torch.cat([q[k] for q in parts]) for k in parts[0]}
    n = det["scores"].shape[0]
    if n > cfg.det_per_img > 0:
        th, _ = torch.kthvalue(det["scores"], n - cfg.det_per_img + 1)
        k = torch.nonzero(det["scores"] >= th.item()).squeeze(1)
        det = {f: v[k] for f, v in det.items()}
    return det


# ------------------------------------------------------------------------------------------- mask refinement
def ciam(x, gamma):  # mask_relation_module.py:190-242 (NORM -1, PRE_NORM False)
    """x [n, C, H, W]: instance-to-instance attention from per-channel Gram matrices, averaged over channels."""
    n, C, H, W = x.shape
    q = x.permute(1, 0, 2, 3).contiguous().view(C, n, -1)
    e = torch.bmm(q, q.permute(0, 2, 1))                                          # [C,n,n]
    e = torch.max(e, -1, keepdim=True)[0].expand_as(e) - e
    att = F.softmax(torch.mean(e, 0), dim=-1)
    out = torch.bmm(att[None], x.view(1, n, -1)).view(n, C, H, W)
    return gamma * out + x


def sort_by_class(labels, obj, num_fg=2):  # mask_relation_module.py:110-156
    """Row order the module works in: class 1 rows by objectness (descending), then class 2 rows, ...; also the
    per-class lengths."""
    order, lengths = [], []
    for c in range(num_fg):
        idx = torch.nonzero(labels == c + 1)[:, 0]
        _, s = torch.sort(obj[idx], descending=True)
        order.append(idx[s])
        lengths.append(len(idx))
    return torch.cat(order), lengths


def mask_relation(sd, p, feats, mask_logits, labels, obj, num_fg=2):  # mask_relation_module.py:55-108
    """feats [P,256,14,14] mask-head features, mask_logits [P,1+C,28,28] of one image's ROIs ->
    (refined logits [P,1+C,28,28] in the sorted row order, that order)."""
    order, lengths = sort_by_class(labels, obj, num_fg)
    own = torch.sigmoid(mask_logits[order, labels[order]])                        # each ROI's own-class probability map
    x = torch.cat((feats[order], F.max_pool2d(own[:, None], kernel_size=2, stride=2)), 1)
    e = p + "appearance_feature_extractor."
    for name in ("mask_fcn1", "mask_fcn2", "mask_fcn3", "conv5_mask"):
        x = F.relu(F.conv2d(x, sd[e + name + ".weight"], sd[e + name + ".bias"], 1, 1))
    groups = [ciam(g, sd[p + "relation_module.gamma"]) for g in torch.split(x, lengths) if g.shape[0]]
    x = torch.cat(groups)
    x = F.relu(F.conv_transpose2d(x, sd[p + "deconv_1.weight"], sd[p + "deconv_1.bias"], 2))
    return F.conv2d(x, sd[p + "classifier.weight"], sd[p + "classifier.bias"]), order


# ------------------------------------------------------------------------------------------- weight tables (tests)
NMS_PREFIX = "relation_nms."
MASK_PREFIX = "mask_heads.mask.mask_relation_module."


def nms_weight_shapes(cfg=RelationCfg):
    """State-dict names / shapes of DuplicationRemovalNetwork under `relation_nms.` (relation_module.py:17-31,103-131)."""
    d, r, g = cfg.feat_dim, cfg.roi_feat_dim, cfg.group
    t = {"nms_rank_fc": (d, r), "roi_feat_embedding_fc": (d, r), "relation_module.WG": (g, cfg.geo_dim),
         "relation_module.WK": (cfg.hid[1], d), "relation_module.WQ": (cfg.hid[0], d),
         "relation_module.conv1": (cfg.hid[2], g * d // g, 1, 1), "classifier": (len(cfg.thresholds), d)}
    out = {}
    for k, s in t.items():
        out[NMS_PREFIX + k + ".weight"] = s
        out[NMS_PREFIX + k + ".bias"] = (s[0],)
    return out


def mask_weight_shapes(ch=16, num_classes=3):
    """State-dict names / shapes of MaskRelationRefineNet (mask_relation_module.py:19-53, relation_mask_feature_extractor.py:16-33)."""
    e = "appearance_feature_extractor."
    t = {e + "mask_fcn1": (256, 257, 3, 3), e + "mask_fcn2": (256, 256, 3, 3), e + "mask_fcn3": (256, 256, 3, 3),
         e + "conv5_mask": (ch, 256, 3, 3), "classifier": (num_classes, ch, 1, 1)}
    out = {MASK_PREFIX + "deconv_1.weight": (ch, ch, 2, 2), MASK_PREFIX + "deconv_1.bias": (ch,),
           MASK_PREFIX + "relation_module.gamma": (1,)}
    for k, s in t.items():
        out[MASK_PREFIX + k + ".weight"] = s
        out[MASK_PREFIX + k + ".bias"] = (s[0],)
    return out
