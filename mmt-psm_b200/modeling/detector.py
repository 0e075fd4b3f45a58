"""GeneralizedRCNN meta-architecture with the mean-teacher entry points.

Mirrors reference modeling/detector/generalized_rcnn.py:17-282 (`forward`, `forward_teacher`,
`forward_student`, `extract_aug_feat`, `get_emb_feature`, `fg_hint_loss`) and detectors.py:5-7
(`build_detection_model(cfg, is_teacher, is_student)`): same signatures, same returned dict keys
(`result_t`, `class_logit_t`, `embedding`, `seg_mask`, `ffi_boxes`), same loss-dict keys.
Internally every feature map is NHWC.  Two result-preserving restructurings of the teacher pass:
  * the K colour views x {orig, h-flip} are ONE batched backbone pass ([2K*B,3,H,W]) instead of 2K passes;
  * the coarse-inference pass re-uses the features of view 0 (the reference runs the backbone on the very
    same tensor twice: generalized_rcnn.py:127 and :204; frozen BN makes both passes identical).
"""
import torch
import torch.nn.functional as F
from torch import nn

from .. import functional as Fk
from ..structures.image_list import ImageList, to_image_list
from .backbone import build_backbone
from .box_head import MaskRCNNFPNAdaptor, build_roi_box_head
from .mask_head import build_roi_mask_head
from .relation import DuplicationRemovalNetwork
from .rpn import build_rpn


class BoxROIHeads(nn.ModuleDict):
    def __init__(self, cfg, heads):
        super().__init__(heads)
        self.cfg = cfg.clone()

    def forward(self, features, proposals, targets=None):
        x, detections, loss_box, class_logits, box_regression = self.box(features, proposals, targets)
        return x, detections, dict(loss_box), class_logits, box_regression

    def forward_student(self, features, proposals, class_logits_t):
        return self.box.forward_student(features, proposals, class_logits_t)

    def forward_teacher(self, feature_tuple, proposals, teacher_infer, batched_feats=None):
        x, detections, loss_box, class_logits, box_regression = self.box.forward_teacher(
            feature_tuple, proposals, targets=teacher_infer, batched_feats=batched_feats)
        return x, detections, dict(loss_box), class_logits, box_regression


class MaskROIHeads(nn.ModuleDict):
    def __init__(self, cfg, heads):
        super().__init__(heads)
        self.cfg = cfg

    def forward(self, losses, features, detections, targets=None, images=None):
        _, detections, loss_mask = self.mask(features, detections, targets, images)
        losses.update(loss_mask)
        return detections, losses


def box_roi_heads(cfg, relation=True):
    return BoxROIHeads(cfg, [("box", build_roi_box_head(cfg, relation=relation))])


def mask_roi_heads(cfg, is_student=False):
    return MaskROIHeads(cfg, [("mask", build_roi_mask_head(cfg, is_student))])


class GeneralizedRCNN(nn.Module):
    def __init__(self, cfg, is_teacher=False, is_student=False):
        super().__init__()
        self.cfg = cfg
        self.backbone = build_backbone(cfg)
        self.rpn = build_rpn(cfg, is_teacher)
        self.box_heads = box_roi_heads(cfg)
        self.mask_heads = mask_roi_heads(cfg, is_student)
        self.relation_nms = DuplicationRemovalNetwork(cfg, is_teacher) if cfg.MODEL.RELATION_NMS.USE_RELATION_NMS else None
        self.mt_fg_hint = cfg.MT.FG_HINT
        self.mt_cls = cfg.MT.CLS_LOSS
        self.hint_adaptor = MaskRCNNFPNAdaptor(cfg)
        self._graphs, self._graphs_on, self._graph_on_backward = {}, False, None
        self._pending, self._side = {}, None
        self.side_priority = 0     # CUDA priority of the trunk-prefetch stream (-1 = high; the trainer raises the teacher's)
        self.trunk_max_ctas = None  # persistent-grid size the trunk graphs are captured with (None = every SM)
        self._stu, self._stu_ptrs, self._deferred = None, {}, []

    # ------------------------------------------------------------------------------ CUDA graphs
    def enable_cuda_graphs(self, on=True, on_backward=None):
        """Replay the static-shape conv trunk (backbone + FPN) from CUDA graphs, one captured instance per call site
        and input shape (utils/cuda_graphs.py).  `on_backward(params)` tells the optimizer which parameters the
        replayed backward accumulated into (their post-accumulate hooks only fire at capture time)."""
        self._graphs_on, self._graph_on_backward = bool(on), on_backward
        self._pending = {}
        if not on:
            self._graphs = {}

    def _trunk_fn(self, site, n_views=1):
        """The fixed-shape part of a pass as one function of the image batch: backbone + FPN, then
        'sup': the RPN head;  'aug' (teacher, n_views stacked views): MGD adaptors on every view and the RPN head on
        view 0;  'unl' (student, unlabeled view): MGD adaptors.  Returns a flat tuple (5 features [+5 emb] [+5 rpn])."""
        want_emb = site != "sup" and bool(self.cfg.MT.FG_HINT or self.cfg.MT.HINT or self.cfg.MT.ODKD)

        if site == "stu":
            # student, labeled batch and unlabeled view stacked ([2B,...], same weights): ONE backbone pass serves the
            # supervised and the consistency branch; RPN head on the labeled half, MGD adaptors on the unlabeled half
            def fn_stu(x):
                feats = tuple(self.backbone(x))
                B = x.shape[0] // 2
                fused = self.rpn.head.forward_fused([f[:B] for f in feats])
                out = feats + tuple(fused) + tuple(self.rpn.head.proposal_scores(fused))
                if want_emb:
                    out = out + tuple(self.hint_adaptor([f[B:] for f in feats]))
                return out
            return fn_stu, want_emb

        def fn(x):
            feats = tuple(self.backbone(x))
            out = feats
            if want_emb:
                out = out + tuple(self.hint_adaptor(feats))
            if site != "unl":
                B = x.shape[0] // n_views
                fused = self.rpn.head.forward_fused([f[:B] for f in feats] if n_views > 1 else list(feats))
                out = out + tuple(fused) + tuple(self.rpn.head.proposal_scores(fused))
            return out
        return fn, want_emb

    def _graph_for(self, site, x, n_views):
        """The captured trunk for (call site, input shape, grad mode); captured on first use.  Returns (graph, fresh)."""
        from ..utils.cuda_graphs import GraphedCallable
        fn, _ = self._trunk_fn(site, n_views)
        need_grad = torch.is_grad_enabled() and self.training
        key = (site, tuple(x.shape), n_views, need_grad)
        g = self._graphs.get(key)
        if g is not None:
            return g, False
        if len(self._graphs) >= 8:
            raise RuntimeError("more than 8 distinct (call site, input shape) trunks captured; CUDA graphs are meant "
                               "for fixed-size crops — call enable_cuda_graphs(False) for variable-size input")
        from .. import kernels as K
        prev = K.set_conv_max_ctas(self.trunk_max_ctas) if self.trunk_max_ctas else None
        try:
            g = GraphedCallable(fn, [x], backward=need_grad, on_backward=self._graph_on_backward)
        finally:
            if self.trunk_max_ctas:
                K.set_conv_max_ctas(prev)
        self._graphs[key] = g
        return g, True

    @staticmethod
    def _aug_batch(tensors, teacher):
        """Stacked trunk input of extract_aug_feat: teacher = every view followed by its h-flip; student = odd views flipped."""
        out = []
        for i, t in enumerate(tensors):
            if teacher:
                out += [t, torch.flip(t, (3,))]
            else:
                out.append(torch.flip(t, (3,)) if i % 2 == 1 else t)
        return (out[0] if len(out) == 1 else torch.cat(out, 0)), len(out)

    def prefetch_trunk(self, site, tensors):
        """Enqueue the trunk of `site` ('sup': the labeled batch; 'aug' / 'unl': list of colour views) on the side
        stream NOW; the matching call inside forward / forward_teacher / forward_student then only waits for its
        completion event.  The three trunks of an iteration are independent, so the host-side glue of one pass (RPN
        selection, samplers, target assignment: dozens of tiny launches and a few read-backs) overlaps the conv work
        of the next instead of serialising with it.  No-op without CUDA graphs."""
        if not self._graphs_on:
            return
        head = tensors if site == "sup" else tensors[0]
        pend = self._pending.get(site)
        if pend is not None and pend[1] == head.data_ptr():
            return      # already enqueued for exactly these tensors (MTtrainer.stage launched it ahead of the iteration)
        if site == "sup":
            x, n_views, first = tensors, 1, tensors
        else:
            x, n_views = self._aug_batch(list(tensors), teacher=(site == "aug"))
            first = tensors[0]
        if not x.is_cuda:
            return
        g, fresh = self._graph_for(site, x, n_views)
        if fresh:
            return  # captured just now (with warm-up runs); the regular call replays it
        cur = torch.cuda.current_stream(x.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=x.device, priority=self.side_priority)
        self._side.wait_stream(cur)
        g.launch_forward([x], self._side)
        self._pending[site] = (g, first.data_ptr(), tuple(x.shape), x)

    def prefetch_student(self, images, views):
        """Stacked student trunk: the labeled batch and the (single, AUG_S = 1) unlabeled view go through the backbone
        as ONE [2B,3,H,W] pass on the side stream.  `forward` then picks up the labeled half — as detached leaves, so
        that the supervised backward issued right after it stops at the trunk boundary — and `forward_student` the
        unlabeled half; `backward_deferred` finally runs the trunk's backward ONCE with both halves' gradients."""
        if not (self._graphs_on and self.training and images.is_cuda and len(views) == 1
                and tuple(views[0].shape) == tuple(images.shape)):
            return False
        pend = self._pending.get("stu")
        if pend is not None and pend[1] == images.data_ptr() and pend[2] == views[0].data_ptr():
            return True   # already enqueued for exactly these tensors (MTtrainer.stage launched it ahead of the iteration)
        x = torch.cat([images, views[0]], 0)
        g, _ = self._graph_for("stu", x, 1)
        cur = torch.cuda.current_stream(x.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=x.device, priority=self.side_priority)
        self._side.wait_stream(cur)
        g.launch_forward([x], self._side)
        self._pending["stu"] = (g, images.data_ptr(), views[0].data_ptr(), x)
        self._stu, self._deferred = None, []
        return True

    def precompute_rpn_targets(self, images, targets):
        """RPN loss targets of the supervised pass whose trunk has just been enqueued (rpn.RPNLossComputation.precompute):
        the grid sizes come from the trunk graph's static outputs, nothing waits for the device."""
        pend = self._pending.get("stu") or self._pending.get("sup")
        if pend is None or not self.training:
            return
        feats = pend[0].static_out[:len(self.rpn.anchor_generator.strides)]
        anchors = self.rpn.anchor_generator(to_image_list(images), [f.shape[1:3] for f in feats])
        self.rpn.loss_evaluator.precompute(anchors, targets)

    def _stu_outputs(self):
        if self._stu is None:
            g, _, _, x = self._pending.pop("stu")
            # plain static buffers: the trunk boundary is handled by hand (leaves in, grad buffers out), which spares
            # the zero-padded slice gradients autograd would build for the two halves of every pyramid level
            flat = tuple(g.forward_detached([x], prefetched=True))
            want_emb = self._trunk_fn("stu")[1]
            n = len(flat) // (3 + int(want_emb))      # feature levels | fused RPN maps | proposal scores | [embeddings]
            self._stu = (flat[:n], (flat[n:2 * n], flat[2 * n:3 * n]), flat[3 * n:] if want_emb else None, x.shape[0] // 2)
            self._stu_graph, self._stu_n = g, n
        return self._stu

    def _leaf(self, tensor, out_index, lo, hi):
        leaf = tensor.detach().requires_grad_(True)
        self._deferred.append((out_index, lo, hi, leaf))
        return leaf

    def backward_deferred(self, extra_losses=None):
        """Backward of the stacked student trunk: `extra_losses` (the consistency terms) are back-propagated through
        their heads down to the boundary leaves; then every boundary gradient — the supervised branch's, produced by
        the early supervised backward, and the consistency branch's — is written into its half of the trunk graph's
        gradient buffers and the captured trunk backward is replayed ONCE."""
        self.backward_deferred_heads(extra_losses)
        self.backward_deferred_trunk()

    def backward_deferred_heads(self, extra_losses=None):
        """First half of `backward_deferred`: the consistency terms' backward through their heads.  After it every
        gradient that does NOT come out of the trunk graph (box / mask / relation heads) is final — a data-parallel
        trainer starts all-reducing that bucket here, under the trunk's backward (MTtrainer.train_step)."""
        if extra_losses is not None:
            extra_losses.backward()

    def backward_deferred_trunk(self):
        g = self._stu_graph
        covered = set()
        for out_index, lo, hi, leaf in self._deferred:
            buf = g.grad_buffer(out_index)
            if buf is None:
                continue
            region = buf if lo is None else buf[lo:hi]
            if leaf.grad is None:
                region.zero_()
            else:
                region.copy_(leaf.grad)
            covered.add((out_index, lo))
        B = self._stu[3]
        for i in g.grad_idx:   # halves nobody consumed this iteration (e.g. the teacher produced no boxes)
            whole = g.static_gout[g.grad_idx.index(i)]
            if i < self._stu_n:
                for lo in (0, B):
                    if (i, lo) not in covered:
                        whole[lo:lo + B].zero_()
            elif (i, None) not in covered:
                whole.zero_()
        self._deferred, self._stu = [], None
        g.backward_direct()

    def _trunk(self, site, x, n_views=1, first_ptr=None):
        """Runs (or replays from a CUDA graph / picks up the prefetched replay of) the trunk of call site `site`;
        returns dict(features, emb, rpn).  `x` may be a callable building the input (skipped when prefetched)."""
        fn, want_emb = self._trunk_fn(site, n_views)
        stu = self._pending.get("stu") if site in ("sup", "unl") else None
        if (stu is not None or self._stu is not None) and first_ptr is not None and n_views == 1:
            key_ptr = (self._pending["stu"][1 if site == "sup" else 2]) if stu is not None else self._stu_ptrs[site]
            if key_ptr == first_ptr:
                if stu is not None:
                    self._stu_ptrs = {"sup": stu[1], "unl": stu[2]}
                feats, rpn, emb, B = self._stu_outputs()
                n = len(feats)
                if site == "sup":
                    fused, scores = rpn
                    f_leaves = tuple(self._leaf(f[:B], i, 0, B) for i, f in enumerate(feats))
                    r_leaves = tuple(self._leaf(t, n + i, None, None) for i, t in enumerate(fused))
                    return {"features": f_leaves, "emb": None, "rpn": (r_leaves, tuple(scores))}
                f_leaves = tuple(self._leaf(f[B:], i, B, 2 * B) for i, f in enumerate(feats))
                e_leaves = tuple(self._leaf(t, 3 * n + i, None, None) for i, t in enumerate(emb)) if emb is not None else None
                return {"features": f_leaves, "emb": e_leaves, "rpn": None}
        pend = self._pending.pop(site, None)
        if pend is not None and first_ptr is not None and pend[1] == first_ptr:
            g, _, shape, keep_alive = pend
            flat = tuple(g(keep_alive, prefetched=True))
        else:
            x = x() if callable(x) else x
            if not (self._graphs_on and x.is_cuda):
                flat = fn(x)
            else:
                g, _ = self._graph_for(site, x, n_views)
                flat = tuple(g(x))
        n = len(flat) // (1 + int(want_emb) + 2 * int(site != "unl"))
        out = {"features": flat[:n], "emb": None, "rpn": None}
        rest = flat[n:]
        if want_emb:
            out["emb"], rest = rest[:n], rest[n:]
        if site != "unl":
            out["rpn"] = (rest[:n], rest[n:2 * n])     # (fused head maps, proposal scores)
        return out

    def set_module_mode(self, mode):
        self.rpn.set_teacher_mode(mode)
        self.box_heads.box.set_teacher_mode(mode)
        if self.relation_nms is not None:
            self.relation_nms.set_teacher_mode(mode)
        self.mask_heads.mask.set_teacher_mode(mode)

    # ------------------------------------------------------------------------------ supervised
    def forward(self, images, targets=None, tta=None, _features=None, _rpn_head=None):
        if self.training and targets is None:
            raise ValueError("In training mode, targets should be passed")
        images = to_image_list(images)
        if _features is None:
            t = self._trunk("sup", images.tensors, first_ptr=images.tensors.data_ptr())
            _features, _rpn_head = t["features"], t["rpn"]
        features = _features
        proposals, proposal_losses = self.rpn(images, features, targets, head_out=_rpn_head)
        x, result, losses, class_logits, box_regression = self.box_heads(features, proposals, targets)
        nms_loss = None
        if self.relation_nms is not None:
            # IR-Net duplicate removal on the box head's ROIs (reference generalized_rcnn.py:63-95, one image per call
            # there; every image is its own attention problem, so the whole batch goes through in one call here):
            # training adds `nms_loss`, eval replaces the greedy-NMS detections the mask head and the teacher consume
            nms_result, nms_loss = self.relation_nms((x, result, F.softmax(class_logits, dim=1), box_regression,
                                                      targets if targets is not None else [None] * len(result)))
            if not self.training:
                result = nms_result
        result, detector_losses = self.mask_heads(losses, features, result, targets, images)
        if self.training:
            out = {}
            out.update(detector_losses)
            out.update(proposal_losses)
            if nms_loss is not None:
                out.update(nms_loss)
            return out
        return result

    # --------------------------------------------------------------------------------- teacher
    def extract_aug_feat(self, imglist, teacher=True):
        """Teacher: every view and its h-flip (the reference flips the ImageList in place and leaves it flipped,
        generalized_rcnn.py:201-208 — reproduced).  One batched backbone pass, split back per view."""
        first = imglist[0].tensors
        B, n = first.shape[0], (2 * len(imglist) if teacher else len(imglist))
        plain = [img.tensors for img in imglist]
        for i, img in enumerate(imglist):   # the reference flips the ImageLists in place and leaves them flipped
            if teacher or i % 2 == 1:
                img.hflip()
        site = "aug" if teacher else "unl"
        t = self._trunk(site, lambda: self._aug_batch(plain, teacher)[0], n_views=n, first_ptr=first.data_ptr())
        split = lambda maps: [tuple(f[i * B:(i + 1) * B] for f in maps) for i in range(n)]
        self._last_emb = split(t["emb"]) if t["emb"] is not None else None   # adaptor outputs computed with the trunk
        self._last_rpn = t["rpn"]
        self._last_stacked = t["features"]   # the un-split [n_views*B, ...] maps
        return split(t["features"])

    def get_emb_feature(self, feature_list):
        return [self.hint_adaptor(f) for f in feature_list]

    def forward_teacher(self, images, targets=None):
        if targets is not None:
            raise NotImplementedError("forward_teacher with ground-truth targets (PLTRAIN) is out of scope")
        images = [to_image_list(im) for im in images]
        view0 = ImageList(images[0].tensors, images[0].image_sizes)  # un-flipped handle on colour view 0
        self.rpn.candidate_sharing(True)     # the coarse inference and the proposal pass post-process the same RPN output
        try:
            return self._forward_teacher(images, view0)
        finally:
            self.rpn.candidate_sharing(False)

    def _forward_teacher(self, images, view0):
        aug_features = self.extract_aug_feat(images)
        # coarse inference on view 0 (eval-mode heads) re-using its backbone features
        self.set_module_mode("test")
        rpn_head = self._last_rpn   # RPN head maps of view 0: used by the coarse inference AND the proposal pass below
        teacher_infer = self.forward(view0, _features=aug_features[0], _rpn_head=rpn_head)
        integral = None
        if self.mt_fg_hint > 0:
            h, w = view0.tensors.shape[-2:]
            integral = self.mask_heads.mask.mask_generator.masker.integral_mask(teacher_infer, h, w)
        self.set_module_mode("train")
        rpn_feat = aug_features[0]
        _, _, _, _, proposals, _, ffi_boxes = self.rpn.forward_teacher(view0, rpn_feat, teacher_infer, head_out=rpn_head)
        embeddings = None
        if self.cfg.MT.FG_HINT or self.cfg.MT.HINT or self.cfg.MT.ODKD:
            embeddings = [list(e) for e in self._last_emb]
        result = class_logits = None
        if self.cfg.MT.CLS_LOSS or self.cfg.MT.ODKD:
            _, result, _, class_logits, _ = self.box_heads.forward_teacher(aug_features, proposals, teacher_infer,
                                                                          batched_feats=self._last_stacked)
        return {"result_t": result, "class_logit_t": class_logits, "embedding": embeddings,
                "seg_mask": integral, "ffi_boxes": ffi_boxes}

    # --------------------------------------------------------------------------------- student
    def forward_student(self, images, result_t):
        loss_dict = {}
        images = [to_image_list(im) for im in (images if isinstance(images, list) else [images])]
        feat_list = self.extract_aug_feat(images, teacher=False)
        if self.cfg.MT.FG_HINT:
            loss_dict.update(mt_fg_loss=self.get_fg_feature_loss(feat_list, result_t["seg_mask"], result_t["embedding"]))
        if self.cfg.MT.FFI or self.cfg.MT.ODKD:
            raise NotImplementedError("MT.FFI / MT.ODKD (comparison methods) are out of scope")
        if self.cfg.MT.CLS_LOSS:
            loss_dict.update(self.box_heads.forward_student(feat_list, result_t["result_t"], result_t["class_logit_t"]))
        return loss_dict

    def get_fg_feature_loss(self, feature_list, seg_mask, teacher_feat):
        if len(feature_list) != 1:
            raise NotImplementedError("MGD is built for MT.AUG_S = 1 (one student view), as scripts/train_mt.sh sets")
        student = list(self._last_emb[0]) if self._last_emb is not None else self.hint_adaptor(feature_list[0])
        return fg_hint_loss(teacher_feat, [student], seg_mask)


def fg_hint_loss(teachers, students, masks):
    """MGD (reference generalized_rcnn.py:243-282).  teachers: list (views) of lists (levels) of NHWC maps, odd
    views are h-flipped views (un-flipped inside the kernel); students: [levels of the single student view];
    masks: integral foreground map [B,H,W] (tensor, or list of [H,W] tensors like the reference)."""
    if isinstance(masks, (list, tuple)):
        masks = torch.stack(list(masks))
    flip = sum(1 << i for i in range(len(teachers)) if i % 2 == 1)
    return Fk.mgd_loss(masks.float(), list(students[0]), [list(t) for t in teachers], flip)


_META = {"GeneralizedRCNN": GeneralizedRCNN}


def build_detection_model(cfg, is_teacher=False, is_student=False):
    return _META[cfg.MODEL.META_ARCHITECTURE](cfg, is_teacher=is_teacher, is_student=is_student)


# ------------------------------------------------------------------------------------------------
# state-dict boundary: reference layout (OIHW convs, IOHW deconv, fc6 columns in (c,h,w) order)
def _special(model):
    from ..layers.conv import Conv2d, ConvTranspose2x2
    conv, deconv = {}, {}
    for name, m in model.named_modules():
        if isinstance(m, Conv2d):
            conv[name + ".weight"] = m
        elif isinstance(m, ConvTranspose2x2):
            deconv[name + ".weight"] = m
    return conv, deconv


def reference_state_dict(model):
    """state_dict with every tensor in the REFERENCE's layout and key names, loadable by the reference's
    checkpointer (utils/checkpoint.py) and model_serialization.py suffix matcher."""
    conv, deconv = _special(model)
    fe = model.box_heads.box.feature_extractor
    out = {}
    for k, v in model.state_dict().items():
        if k in conv:
            out[k] = conv[k].reference_weight()
        elif k in deconv:
            out[k] = deconv[k].reference_weight()
        elif k == "box_heads.box.feature_extractor.fc6.weight":
            out[k] = fe.fc6_reference_weight()
        else:
            out[k] = v.detach().clone()
    return out


def load_reference_state_dict(model, sd, strict=False):
    """Load a reference-layout state dict (e.g. e2e_mask_rcnn_R_50_FPN_1x.pth contents).  Keys this model does
    not build (e.g. the IR-Net modules when they are switched off in cfg) are ignored unless strict."""
    conv, deconv = _special(model)
    fe = model.box_heads.box.feature_extractor
    own = model.state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    if strict and (missing or unexpected):
        raise KeyError("missing %s unexpected %s" % (missing, unexpected))
    with torch.no_grad():
        for k, v in sd.items():
            if k not in own:
                continue
            v = v.to(own[k].device, torch.float32)
            if k in conv:
                conv[k].load_reference_weight(v)
            elif k in deconv:
                deconv[k].load_reference_weight(v)
            elif k == "box_heads.box.feature_extractor.fc6.weight":
                fe.load_fc6_reference_weight(v)
            else:
                own[k].copy_(v)
    return missing, unexpected
