"""IR-Net relation modules of the MMT-PSM detector (BASELINE config 4).

* `DuplicationRemovalNetwork` / `RelationModule` — learned duplicate removal on the box head's ranked detections
  (reference modeling/relation/relation_module.py:13-91, 94-331, label preparation :333-473, ranking :475-593,
  embeddings :607-682): trains an IoU regressor (`nms_loss`, REG_IOU) on the supervised pass and replaces the
  score-threshold + greedy-NMS post-processing of the teacher's coarse inference.
* `MaskRelationRefineNet` / `CIAM_Module` / `RoiAlignMaskFeatureExtractor` — instance-to-instance refinement of the mask
  logits (reference modeling/relation/mask_relation_module.py:16-156, 190-242, relation_mask_feature_extractor.py:10-47).

Same constructor arguments, parameter names and state-dict keys as the reference (`relation_nms.relation_module.WK.weight`,
`mask_heads.mask.mask_relation_module.appearance_feature_extractor.mask_fcn1.weight`, ...).  What is done differently:

* all images of the batch go through ONE pass (the reference calls the module once per image, generalized_rcnn.py:72-84,
  mask_head.py:108-115): the attention carries a leading image dimension; the instance attention of the mask branch is
  one masked [P, P] softmax over all positive ROIs (instances only attend within their own image and class);
* the IoU-regression targets (numpy argmax / intersect1d on the host after a device->host copy of boxes and scores,
  relation_module.py:333-413) are computed on the device without a read-back, ties resolved towards the first index like
  numpy's argmax and intersect1d's first occurrence;
* eval mode selects survivors with the segmented on-device NMS (keep flags) and ONE read-back (the final row list) instead
  of nonzero + NMS (+ D2H) per class and image;
* the 3x3 convs of the appearance extractor (257 -> 256 -> 256 -> 256 -> 16 @14x14) and the 1024 / 128-wide linears run on
  the tcgen05 conv engine over NHWC ROI tiles (the 257-channel input and the 16-channel output are zero-padded to the
  engine's 32-channel granularity at call time; the parameters keep the reference's shapes).  The layers narrower than
  that granularity (WG 64 -> 16, the 128 -> 1 IoU classifier, the 16-channel deconv / 1x1 classifier of the mask branch,
  < 0.1 GFLOP together) and the [N, N] attention products are plain fp32 matmuls.
Only the configuration the repo's yaml + scripts/train_mt.sh select is built: USE_IOU False, CLASS_AGNOSTIC False,
CLS_WISE_RELATION False, MERGE_METHOD >= 0; RELATION_MASK TYPE CIAM, FEATURE_EXTRACTOR RoiAlignMaskFeatureExtractor,
EXTRACTOR_CHANNEL != 1, SAME_PREDICTOR False, NORM -1, PRE_NORM False.  Anything else raises at construction."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import functional as Fk
from .. import kernels as K
from ..layers import Conv2d, ConvTranspose2x2
from ..layers.conv import kaiming_normal_fanout_ohwi_
from ..structures.bounding_box import BoxList
from .box_coder import BoxCoder
from .box_head import LinearParam


def _linear(n_in, n_out):
    """LinearParam with nn.Linear's default initialisation (uniform +-1/sqrt(fan_in) for weight and bias)."""
    m = LinearParam(n_in, n_out)
    bound = 1.0 / math.sqrt(n_in)
    with torch.no_grad():
        m.weight.uniform_(-bound, bound)
        m.bias.uniform_(-bound, bound)
    return m


# ----------------------------------------------------------------------------------------------- embeddings
def extract_rank_embedding(rank_dim, feat_dim, wave_length=1000, device="cpu"):
    """[rank_dim, feat_dim] sin / cos code of the rank (reference relation_module.py:607-627)."""
    r = torch.arange(0, rank_dim, device=device).float().view(-1, 1)
    k = torch.arange(feat_dim / 2, device=device)
    m = r * (1.0 / torch.pow(wave_length, k / (feat_dim / 2))).view(1, -1)
    return torch.cat((torch.sin(m), torch.cos(m)), -1)


def extract_multi_position_matrix(boxes, iou, dim_g, wave_len, clswise=False):
    """boxes [..., N, C, 4] -> [..., C, N, N, dim_g]: sin / cos code of (log|dx/w_i|, log|dy/h_i|, log(w_i/w_j),
    log(h_i/h_j)) between the ranked boxes of one class (reference relation_module.py:629-682, iou None)."""
    if iou is not None or clswise:
        raise NotImplementedError("RELATION_NMS.USE_IOU / CLS_WISE_RELATION are off in the MMT-PSM configuration")
    b = boxes.transpose(-2, -3)                                   # [..., C, N, 4]
    x0, y0, x1, y1 = b.unbind(-1)
    cx, cy = (x0 + x1) * 0.5, (y0 + y1) * 0.5
    w, h = (x1 - x0) + 1.0, (y1 - y0) + 1.0
    col, row = (lambda t: t.unsqueeze(-1)), (lambda t: t.unsqueeze(-2))
    dx = torch.log(torch.clamp(torch.abs((col(cx) - row(cx)) / col(w)), min=1e-3))
    dy = torch.log(torch.clamp(torch.abs((col(cy) - row(cy)) / col(h)), min=1e-3))
    dw = torch.log(col(w) / row(w))
    dh = torch.log(col(h) / row(h))
    pos = torch.stack((dx, dy, dw, dh), -1)                       # [..., C, N, N, 4]
    k = torch.arange(dim_g / 8, device=boxes.device)
    freq = 1.0 / torch.pow(wave_len, k / (dim_g / 8))
    m = ((100.0 * pos).unsqueeze(-1) * freq).flatten(-2)          # [..., C, N, N, dim_g / 2]
    return torch.cat((torch.sin(m), torch.cos(m)), -1)


# ----------------------------------------------------------------------------------------------- relation attention
class RelationModule(nn.Module):
    """Multi-head relation attention (reference relation_module.py:13-91): per group g the logits are
    log(max(relu(WG pos), 1e-6))_g + <WQ f_i, WK f_j>_g / sqrt(d_g), softmax over each row's top-k, the values are the
    input features themselves and every group's result goes through its own [D -> dim[2]/group] map (a grouped 1x1 conv)."""

    def __init__(self, appearance_feature_dim=1024, geo_feature_dim=64, fc_dim=(64, 16), group=16, dim=(1024, 1024, 1024),
                 topk=10, iou_method="b"):
        super().__init__()
        assert fc_dim[1] == group, "Check the dimensions in attention!"
        self.fc_dim, self.dim, self.group, self.topk, self.iou_method = fc_dim, dim, group, topk, iou_method
        self.dim_group = (dim[0] // group, dim[1] // group, dim[2] // group)
        self.WG = _linear(geo_feature_dim, fc_dim[1])
        self.WK = _linear(appearance_feature_dim, dim[1])
        self.WQ = _linear(appearance_feature_dim, dim[0])
        # grouped 1x1 conv (groups = group) over fc_dim[1] * D input channels: weight [dim[2], D, 1, 1] in the reference's
        # layout, i.e. a Conv2d(D -> dim[2]) parameter whose output rows [g * dim[2]/group, ...) belong to group g
        self.conv1 = Conv2d(appearance_feature_dim, dim[2], 1, 1, 0)

    def forward(self, f_a, position_embedding, iou=None):
        """f_a [..., N, C, D], position_embedding [..., C, N, N, geo] -> [..., N, C, dim[2]] (any leading image dims)."""
        if iou is not None:
            raise NotImplementedError("RELATION_NMS.USE_IOU is off in the MMT-PSM configuration")
        lead = f_a.shape[:-3]
        N, C, D = f_a.shape[-3:]
        G, (dq, dk, dr) = self.group, self.dim_group
        f = f_a.reshape(-1, N, C, D).transpose(1, 2).contiguous()                   # [B,C,N,D]
        B = f.shape[0]
        flat = f.view(B * C * N, D)
        w_g = F.relu(F.linear(position_embedding.reshape(-1, position_embedding.shape[-1]), self.WG.weight, self.WG.bias))
        w_g = w_g.view(B, C, N, N, G).permute(0, 1, 4, 2, 3)                        # [B,C,G,N,N]
        k = self.WK(flat).view(B, C, N, G, dk).permute(0, 1, 3, 4, 2)               # [B,C,G,dk,N]
        q = self.WQ(flat).view(B, C, N, G, dq).permute(0, 1, 3, 2, 4)               # [B,C,G,N,dq]
        logit = torch.log(torch.clamp(w_g, min=1e-6)) + (1.0 / math.sqrt(float(dk))) * torch.matmul(q, k)
        top, ind = torch.topk(logit, min(N, self.topk), dim=-1, largest=True, sorted=True)
        att = torch.zeros_like(logit).scatter(-1, ind, F.softmax(top, dim=-1))
        o = torch.matmul(att, f.unsqueeze(2))                                       # [B,C,G,N,D]
        w = self.conv1.weight.view(G, dr, D)
        out = torch.einsum("bcgnd,grd->bncgr", o, w).reshape(B, N, C, G * dr) + self.conv1.bias
        return out.reshape(*lead, N, C, G * dr)


def iou_regression_targets(boxes, scores, gt_boxes, gt_labels, thresholds, binary=False):
    """Targets of the duplicate-removal head, on the device (reference relation_module.py:333-413; :415-473 for binary).
    boxes [N, C, 4] ranked per-class boxes, scores [N, C], gt_boxes [M, 4], gt_labels [M] -> float32 [N, C, T].
    Per class and threshold every ground-truth box of that class picks, among the ranked boxes whose best-overlapping
    ground truth it is and whose IoU with it exceeds the threshold, the one with the highest class score; that box's
    target is its IoU (binary: 1), every other box's 0.  A ground truth without a candidate falls on rank 0 exactly as
    the reference's argmax over an all-zero column does, and rank 0 then takes the value of the FIRST ground truth that
    points at it (np.intersect1d's first-occurrence index)."""
    N, C = scores.shape
    M, dev = gt_boxes.shape[0], boxes.device
    out = boxes.new_zeros((N, C, len(thresholds)))
    if M == 0 or N == 0:
        return out
    b = boxes.permute(1, 0, 2)                                                      # [C,N,4]
    area_b = (b[..., 2] - b[..., 0] + 1) * (b[..., 3] - b[..., 1] + 1)
    area_g = (gt_boxes[:, 2] - gt_boxes[:, 0] + 1) * (gt_boxes[:, 3] - gt_boxes[:, 1] + 1)
    lt = torch.max(b[:, :, None, :2], gt_boxes[:, :2])
    rb = torch.min(b[:, :, None, 2:], gt_boxes[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    iou = inter / (area_b[:, :, None] + area_g - inter)                             # [C,N,M]
    valid = gt_labels[None, :] == torch.arange(1, C + 1, device=dev)[:, None]       # [C,M] ground truth of this class
    iou = torch.where(valid[:, None, :], iou, iou.new_full((), -1.0))
    m_ids = torch.arange(M, device=dev)
    n_ids = torch.arange(N, device=dev)
    best = torch.where(iou == iou.max(2, keepdim=True)[0], m_ids, M).min(2)[0]      # first argmax over ground truth
    mine = m_ids == best[:, :, None]                                                # [C,N,M]
    s = scores.t()[:, :, None]                                                      # [C,N,1]
    for t, thr in enumerate(thresholds):
        over = iou > thr
        cand = torch.where(over & mine, s, s.new_zeros(()))                         # [C,N,M]
        winner = torch.where(cand == cand.max(1, keepdim=True)[0], n_ids[:, None], N).min(1)[0]     # [C,M] first argmax
        value = torch.where(over & mine, iou, iou.new_zeros(())).gather(1, winner[:, None, :])[:, 0]  # [C,M]
        points = (winner[:, None, :] == n_ids[None, :, None]) & valid[:, None, :]   # [C,N,M] ground truth m points at n
        first = torch.where(points, m_ids, M).min(2)[0]                             # [C,N]
        hit = (first < M) & over.any(2)
        val = value.gather(1, first.clamp(max=M - 1))
        tgt = torch.where(hit, torch.ones_like(val) if binary else val, val.new_zeros(()))
        out[:, :, t] = tgt.t()
    return out


class DuplicationRemovalNetwork(nn.Module):
    def __init__(self, cfg, is_teacher=False):
        super().__init__()
        self.cfg = cfg.clone()
        r = cfg.MODEL.RELATION_NMS
        if r.USE_IOU or r.CLASS_AGNOSTIC or r.CLS_WISE_RELATION or r.MERGE_METHOD < 0:
            raise NotImplementedError("RELATION_NMS: only USE_IOU False, CLASS_AGNOSTIC False, CLS_WISE_RELATION False, "
                                      "MERGE_METHOD >= 0 (the MMT-PSM yaml) are built")
        self.reg_iou = r.REG_IOU
        self.first_n = r.FIRST_N
        self.NMS_thread = self.target_thresh = tuple(r.THREAD)
        self.geo_feature_dim = r.GEO_FEAT_DIM
        self.nms_rank_fc = _linear(r.ROI_FEAT_DIM, r.APPEARANCE_FEAT_DIM)
        self.roi_feat_embedding_fc = _linear(r.ROI_FEAT_DIM, r.APPEARANCE_FEAT_DIM)
        self.relation_module = RelationModule(r.APPEARANCE_FEAT_DIM, geo_feature_dim=self.geo_feature_dim,
                                              fc_dim=(self.geo_feature_dim, 16), group=r.GROUP, dim=tuple(r.HID_DIM),
                                              topk=r.TOPK, iou_method=r.IOU_METHOD)
        self.boxcoder = BoxCoder(weights=(10., 10., 5., 5.))
        self.class_agnostic = r.CLASS_AGNOSTIC
        self.fg_class = cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES - 1
        self.classifier = _linear(r.APPEARANCE_FEAT_DIM, len(self.target_thresh))
        self.fg_thread = r.FG_THREAD
        self.detections_per_img = cfg.MODEL.ROI_HEADS.DETECTIONS_PER_IMG
        self.nms = r.POS_NMS
        self.nuclei_nms = 0.5                      # greedy NMS threshold of class 2, fixed in the reference (:264-266)
        self.mode = None
        self._rank_cache = {}

    def set_teacher_mode(self, mode):
        self.mode = mode

    # ---- ranking (reference prepare_ranking / filter_results :475-593), one image
    @torch.no_grad()
    def _rank(self, prob, reg, proposals, target):
        """prob [n, 1+C], reg [n, 4(1+C)] of one image -> dict(idx [N,C], scores [N,C], all_scores [N,C,C], obj [N,C],
        boxes [N,C,4] clipped to the image, labels [N,C,T] in training)."""
        C, n = self.fg_class, prob.shape[0]
        scores = prob[:, 1:]
        N = min(n, self.first_n)
        top, idx = torch.topk(scores, N, dim=0, largest=True, sorted=True)
        cls = torch.arange(C, device=prob.device)
        deltas = reg.view(n, C + 1, 4)[idx, cls + 1]                                # [N,C,4]: class c's regression of its rank-r ROI
        boxes = self.boxcoder.decode(deltas.reshape(-1, 4), proposals.bbox[idx].reshape(-1, 4)).view(N, C, 4)
        out = dict(idx=idx, scores=top, all_scores=scores[idx], obj=proposals.get_field("objectness")[idx])
        if self.training:
            gt, gl = target.bbox, target.get_field("labels")
            out["labels"] = iou_regression_targets(boxes, top, gt, gl, self.target_thresh, binary=not self.reg_iou)
        w, h = proposals.size
        hi = boxes.new_tensor([w - 1, h - 1, w - 1, h - 1])
        out["boxes"] = torch.min(boxes.clamp(min=0), hi)
        return out

    def _rank_code(self, N, device):
        key = (N, str(device))
        if key not in self._rank_cache:
            self._rank_cache[key] = extract_rank_embedding(N, self.cfg.MODEL.RELATION_NMS.ROI_FEAT_DIM, device=device)
        return self._rank_cache[key]

    def _scores(self, x, ranked):
        """x [n_total, ROI_FEAT_DIM] features of all images (rows of image i at offset ranked[i]['offset']) ->
        [B, N, C, T] regressed IoU (sigmoid keep-probability without REG_IOU); all images share N."""
        r = self.cfg.MODEL.RELATION_NMS
        emb = self.roi_feat_embedding_fc(x)
        feat = torch.stack([emb[q["idx"] + q["offset"]] for q in ranked])           # [B,N,C,D]
        B, N = feat.shape[:2]
        rank = self.nms_rank_fc(self._rank_code(N, x.device))                       # [N,D]
        feat = feat + rank[None, :, None, :]
        pos = extract_multi_position_matrix(torch.stack([q["boxes"] for q in ranked]), None, self.geo_feature_dim, 1000)
        feat = F.relu(feat + self.relation_module(feat, pos, None))
        out = F.linear(feat.reshape(-1, r.APPEARANCE_FEAT_DIM), self.classifier.weight, self.classifier.bias)
        out = out.view(B, N, self.fg_class, len(self.target_thresh))
        return out if self.reg_iou else torch.sigmoid(out)

    def forward(self, x):
        """x = (appearance_feature [sum n_i, ROI_FEAT_DIM], proposals list[BoxList] (field `objectness`), cls_score
        [sum n_i, 1+C] softmax scores, box_reg [sum n_i, 4(1+C)], targets list[BoxList] | list[None]).
        Training: (None, {'nms_loss': mean over images}).  Eval: (list[BoxList] with scores / objectness / all_scores /
        labels, {}).  The reference takes ONE image per call (its multi-image branch mixes images in one attention,
        relation_module.py:181-188); here every image is its own attention problem, so calling with the whole batch
        equals the reference's per-image loop (generalized_rcnn.py:72-95)."""
        appearance_feature, proposals, cls_score, box_reg, targets = x
        counts = [len(p) for p in proposals]
        ranked, off = [], 0
        for p, prob, reg, t in zip(proposals, cls_score.split(counts), box_reg.split(counts), targets):
            q = self._rank(prob, reg, p, t)
            q["offset"] = off
            off += len(p)
            ranked.append(q)
        groups = {}                                                                 # images with equal N share one pass
        for i, q in enumerate(ranked):
            groups.setdefault(q["idx"].shape[0], []).append(i)
        out = [None] * len(ranked)
        for _, members in groups.items():
            s = self._scores(appearance_feature, [ranked[i] for i in members])
            for j, i in enumerate(members):
                out[i] = s[j]
        if self.training:
            losses = []
            for q, o in zip(ranked, out):
                if self.reg_iou:
                    losses.append(F.mse_loss(q["labels"], o))
                else:
                    s = q["scores"][:, :, None]
                    losses.append(F.binary_cross_entropy(s * (s * o), q["labels"]))
            return None, dict(nms_loss=torch.mean(torch.stack(losses)))
        return self._detect(ranked, out, [p.size for p in proposals]), {}

    # ---- eval: survivors (reference :237-331)
    @torch.no_grad()
    def _detect(self, ranked, out, sizes):
        r = self.cfg.MODEL.RELATION_NMS
        C = self.fg_class
        merge = min(max(r.MERGE_METHOD, 0), len(self.target_thresh))
        finals = []
        for q, o in zip(ranked, out):
            s_in = q["scores"][:, :, None]
            s = o * (s_in > self.fg_thread).float() if self.reg_iou else o * s_in
            finals.append(s[:, :, merge])                                           # [N,C]
        # greedy NMS per (image, class): rows below FG_THREAD come last in score order, so they cannot suppress a row
        # above it — NMS over all N rows followed by the threshold mask equals the reference's filter-then-NMS
        keep = [None] * len(ranked)
        thr = [self.nms if c == 0 else self.nuclei_nms for c in range(C)]
        for c in range(C):
            boxes = torch.cat([q["boxes"][:, c] for q in ranked])
            sc = torch.cat([f[:, c] for f in finals])
            offs = [0]
            for q in ranked:
                offs.append(offs[-1] + q["boxes"].shape[0])
            flags = K.nms_segmented_flags(boxes, sc, offs, thr[c]) if thr[c] > 0 else torch.ones_like(sc, dtype=torch.bool)
            flags = flags & (sc >= self.fg_thread)
            for i in range(len(ranked)):
                part = flags[offs[i]:offs[i + 1]]
                keep[i] = part[:, None] if keep[i] is None else torch.cat([keep[i], part[:, None]], 1)
        results = []
        for q, f, k, size in zip(ranked, finals, keep, sizes):
            # reference order: class C (nuclei) first, then class C-1, ..., inside a class by rank
            sel = k.t().flip(0).nonzero()                                           # rows (C-1-c, n), lexicographic
            c, n = C - 1 - sel[:, 0], sel[:, 1]
            det = BoxList(q["boxes"][n, c], size, mode="xyxy")
            det.add_field("scores", f[n, c])
            det.add_field("objectness", q["obj"][n, c])
            det.add_field("all_scores", q["all_scores"][n, c])
            det.add_field("labels", (c + 1).to(torch.int64))
            if len(det) > self.detections_per_img > 0:
                sc = det.get_field("scores")
                kth, _ = torch.kthvalue(sc, len(det) - self.detections_per_img + 1)
                det = det[torch.nonzero(sc >= kth).squeeze(1)]
            results.append(det)
        return results


# ----------------------------------------------------------------------------------------------- mask refinement
class CIAM_Module(nn.Module):
    """Class-wise instance attention (reference mask_relation_module.py:190-242, NORM -1, PRE_NORM False): per channel
    the Gram matrix of the instances' maps, e = rowmax(e) - e, mean over channels, softmax -> every instance becomes
    itself + gamma * attention-weighted sum of the instances of its group."""

    def __init__(self, cfg):
        super().__init__()
        m = cfg.MODEL.RELATION_MASK
        if m.NORM in (1, 2) or m.PRE_NORM:
            raise NotImplementedError("RELATION_MASK.NORM 1/2 and PRE_NORM are off in the MMT-PSM configuration")
        self.norm, self.topk, self.prenorm = m.NORM, m.TOPK, m.PRE_NORM
        self.gamma = nn.Parameter(torch.zeros(1))

    def forward(self, x, group=None):
        """x NHWC [n, H, W, C]; group [n] int ids (instances attend only within their id; None = one group)."""
        n, H, W, C = x.shape
        q = x.permute(3, 0, 1, 2).reshape(C, n, H * W)
        e = torch.bmm(q, q.transpose(1, 2))                                         # [C,n,n]
        if group is not None:
            same = group[:, None] == group[None, :]
            e = torch.where(same, e, e.new_full((), float("-inf")))
        e = e.max(-1, keepdim=True)[0] - e                                          # +inf outside the group
        e = torch.mean(e, 0)
        if group is not None:
            e = torch.where(same, e, e.new_full((), float("-inf")))
        att = F.softmax(e, dim=-1)
        out = torch.mm(att, x.reshape(n, -1)).view(n, H, W, C)
        return self.gamma * out + x


class RoiAlignMaskFeatureExtractor(nn.Module):
    """Appearance features of the refinement branch (reference relation_mask_feature_extractor.py:10-47): the mask-head
    ROI features concatenated with the 2x2-max-pooled own-class mask probability, 3x conv3x3(256) + conv3x3(16), ReLU."""

    def __init__(self, cfg):
        super().__init__()
        ch = cfg.MODEL.RELATION_MASK.EXTRACTOR_CHANNEL
        if ch == 1:
            raise NotImplementedError("RELATION_MASK.EXTRACTOR_CHANNEL = 1 is not the MMT-PSM configuration")
        self.mask_fcn1 = Conv2d(257, 256, 3, 1, 1)
        self.mask_fcn2 = Conv2d(256, 256, 3, 1, 1)
        self.mask_fcn3 = Conv2d(256, 256, 3, 1, 1)
        self.conv5_mask = Conv2d(256, ch, 3, 1, 1)
        self.mask_fcn_logits = None
        for m in (self.mask_fcn1, self.mask_fcn2, self.mask_fcn3, self.conv5_mask):
            kaiming_normal_fanout_ohwi_(m.weight)
            nn.init.constant_(m.bias, 0)

    def forward(self, x):
        """x = (ROI features NHWC [P,14,14,256], mask probability [P,1,28,28]) -> NHWC [P,14,14,ch]."""
        feats, mask = x
        pooled = F.max_pool2d(mask, kernel_size=2, stride=2)[:, 0]
        c_in = feats.shape[-1] + 1
        pad_in = (-c_in) % 32
        x = torch.cat([feats, pooled.unsqueeze(-1), feats.new_zeros(feats.shape[:-1] + (pad_in,))], -1)
        x = Fk.conv_bn_act(x, F.pad(self.mask_fcn1.weight, (0, pad_in)), None, self.mask_fcn1.bias, None, True, 1, 1)
        x = self.mask_fcn2(x, relu=True)
        x = self.mask_fcn3(x, relu=True)
        ch = self.conv5_mask.out_channels
        pad_out = (-ch) % 32
        w = torch.cat([self.conv5_mask.weight, x.new_zeros((pad_out,) + tuple(self.conv5_mask.weight.shape[1:]))], 0)
        b = torch.cat([self.conv5_mask.bias, x.new_zeros(pad_out)], 0)
        return Fk.conv_bn_act(x, w, None, b, None, True, 1, 1)[..., :ch].contiguous()


_RELATION_MASK_FEATURE_EXTRACTORS = {"RoiAlignMaskFeatureExtractor": RoiAlignMaskFeatureExtractor}


def make_relation_mask_feature_extractor(cfg):
    name = cfg.MODEL.RELATION_MASK.FEATURE_EXTRACTOR
    if name not in _RELATION_MASK_FEATURE_EXTRACTORS:
        raise NotImplementedError("RELATION_MASK.FEATURE_EXTRACTOR %s: only RoiAlignMaskFeatureExtractor (the MMT-PSM yaml)" % name)
    return _RELATION_MASK_FEATURE_EXTRACTORS[name](cfg)


class MaskRelationRefineNet(nn.Module):
    def __init__(self, cfg, predictor=None):
        super().__init__()
        self.cfg = cfg.clone()
        m = cfg.MODEL.RELATION_MASK
        if m.TYPE != "CIAM" or m.SAME_PREDICTOR:
            raise NotImplementedError("RELATION_MASK: only TYPE CIAM with its own predictor (the MMT-PSM yaml) is built")
        ch = m.EXTRACTOR_CHANNEL
        self.fg_class = cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES - 1
        self.appearance_feature_extractor = make_relation_mask_feature_extractor(cfg)
        self.relation_module = CIAM_Module(cfg)
        self.deconv_1 = ConvTranspose2x2(ch, ch)
        self.classifier = Conv2d(ch, cfg.MODEL.ROI_BOX_HEAD.NUM_CLASSES, 1, 1, 0)
        bound = 1.0 / math.sqrt(ch * 4)                            # nn.ConvTranspose2d default: fan_in = out_channels * k * k
        with torch.no_grad():
            self.deconv_1.weight.uniform_(-bound, bound)
            self.deconv_1.bias.uniform_(-bound, bound)

    @staticmethod
    def sort_rows(proposals, num_fg):
        """Row order the refinement works in (reference prepare_msk_relation :110-156): per image the class-1 ROIs by
        objectness (descending), then class 2, ...  Returns (order [P] into the concatenated rows, group id [P] =
        image * num_fg + class - 1 of the SORTED rows, sorted labels [P]); no read-back."""
        labels = torch.cat([p.get_field("labels") for p in proposals])
        obj = torch.cat([p.get_field("objectness") for p in proposals])
        img = torch.cat([torch.full((len(p),), i, dtype=torch.int64, device=labels.device) for i, p in enumerate(proposals)])
        by_obj = torch.sort(obj, descending=True, stable=True)[1]
        gid = img * num_fg + (labels - 1)
        order = by_obj[torch.sort(gid[by_obj], stable=True)[1]]
        return order, gid[order], labels[order]

    def refine(self, features_roi, mask_logits, proposals):
        """All images at once.  features_roi NHWC [P,14,14,256] (mask feature extractor output), mask_logits NHWC
        [P,28,28,1+C], proposals list[BoxList] with `labels` (1..C) and `objectness`.
        Returns (refined logits NHWC [P,28,28,1+C] in the sorted row order, list[BoxList] sorted the same way)."""
        counts = [len(p) for p in proposals]
        if sum(counts) == 0:
            return mask_logits, list(proposals)
        order, gid, labels = self.sort_rows(proposals, self.fg_class)
        rows = torch.arange(order.shape[0], device=order.device)
        own = torch.sigmoid(mask_logits[order][rows, :, :, labels])                 # [P,28,28] own-class probability
        x = self.appearance_feature_extractor((features_roi[order], own[:, None]))
        x = self.relation_module(x, gid)
        # ConvTranspose2d(ch, ch, 2, 2) + ReLU + 1x1 classifier on 16 channels: two small matmuls and a pixel shuffle
        P, H, W, ch = x.shape
        y = F.relu(torch.matmul(x, self.deconv_1.weight.view(4 * ch, ch).t()) + self.deconv_1.bias.repeat(4))
        y = y.view(P, H, W, 2, 2, ch).permute(0, 1, 3, 2, 4, 5).reshape(P, 2 * H, 2 * W, ch)
        logits = torch.matmul(y, self.classifier.weight.view(-1, ch).t()) + self.classifier.bias
        out, off = [], 0
        for p, n in zip(proposals, counts):
            out.append(p[order[off:off + n] - off])
            off += n
        return logits, out

    def forward(self, x):
        """Reference call convention, one image: x = (features_RoI, mask_logits, proposals BoxList, targets) ->
        (refined logits, [sorted BoxList], targets, None)."""
        features_roi, mask_logits, proposals, targets = x
        logits, boxes = self.refine(features_roi, mask_logits, [proposals])
        return logits, boxes, targets, None
