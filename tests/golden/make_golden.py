"""Generates tests/golden/*.npz by EXECUTING THE (PATCHED) REFERENCE in this container.

    python oracle/build_ref.py            # builds /tmp/mmtpsm_ref (patched copy, see that file)
    python tests/golden/make_golden.py    # writes the fixtures next to this script

The fixtures are small seeded input/output pairs of the reference's own functions on the hot path;
they travel to the GPU box (where /root/reference does not exist) and pin both the CPU oracle
(`oracle/ops_ref.py`, test_oracle_cpu.py) and the CUDA path (`-m gpu` tests).
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

TREE = os.environ.get("MMTPSM_REF_TREE", "/tmp/mmtpsm_ref")
sys.path.insert(0, TREE)
HERE = os.path.dirname(os.path.abspath(__file__))

from maskrcnn_benchmark import _C  # noqa: E402  (reference csrc, CPU build)
from maskrcnn_benchmark.config import cfg  # noqa: E402

MT_OPTS = ["MT.ALPHA", 0.99, "MT.LAMBDA", 5., "MT.RAMPUP_STEP", 250, "MT.RAMPDOWN_STEP", 250, "MT.TEMP", 0.5,
           "MT.SHARPEN", True, "MT.START_MT", 1000, "MT.AUG_S", 1, "MT.CLS_LOSS_TYPE", "bce", "MT.CLS_LOSS", 0.2,
           "MT.HARD_NEG", True, "MT.CLS_BALANCE_WEIGHT", 1.5, "MT.RANK_FILTER", 0.2, "MT.FG_HINT", 1.,
           "MT.T_ADAPT", True, "MODEL.ROI_BOX_HEAD.DO", 0.5, "SOLVER.BASE_LR", 0.005]


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


def clustered_boxes(g, n, size=800.0):
    centers = torch.rand(max(n // 6, 1), 2, generator=g) * size
    idx = torch.randint(0, centers.shape[0], (n,), generator=g)
    c = centers[idx] + torch.randn(n, 2, generator=g) * 6
    wh = torch.rand(n, 2, generator=g) * 60 + 8
    b = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, size - 1)
    return b.float()


def gen_nms():
    g = torch.Generator().manual_seed(11)
    out = {}
    for i, (n, thr) in enumerate([(300, 0.7), (257, 0.5), (64, 0.3), (1, 0.5), (900, 0.7)]):
        b = clustered_boxes(g, n)
        s = torch.rand(n, generator=g)
        keep = _C.nms(b, s, thr)  # CPU flavour (>=); inputs have no IoU == thr (checked in the test)
        out.update({"boxes%d" % i: b.numpy(), "scores%d" % i: s.numpy(), "thr%d" % i: thr, "keep%d" % i: keep.numpy()})
    # integer-coordinate boxes with many exactly-equal scores (sigmoid saturation case): the reference
    # CPU sort order for ties is recorded through its result
    save("nms", n_cases=5, **out)


def gen_roi_align():
    g = torch.Generator().manual_seed(12)
    out = {}
    cases = [(2, 8, 40, 50, 30, 0.25, 7, 2), (1, 4, 25, 25, 17, 1 / 16., 14, 2), (2, 12, 13, 13, 9, 1 / 64., 7, 0),
             (3, 8, 31, 47, 25, 0.125, 7, 3)]
    for i, (B, C, H, W, K, scale, P, sr) in enumerate(cases):
        x = torch.randn(B, C, H, W, generator=g)
        img_w, img_h = W / scale, H / scale
        ctr = torch.rand(K, 2, generator=g) * torch.tensor([img_w, img_h])
        wh = torch.rand(K, 2, generator=g) * torch.tensor([img_w, img_h]) * 0.5 + 1
        boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        boxes[0] = torch.tensor([-20., -30., 10., 12.])           # partly outside
        boxes[1] = torch.tensor([img_w - 5, img_h - 4, img_w + 40, img_h + 30])
        boxes[2] = torch.tensor([5., 5., 5.2, 5.1])                # degenerate (< 1 px on the map)
        ids = torch.randint(0, B, (K, 1), generator=g).float()
        rois = torch.cat([ids, boxes], 1)
        y = _C.roi_align_forward(x, rois, scale, P, P, sr)
        out.update({"x%d" % i: x.numpy(), "rois%d" % i: rois.numpy(), "cfg%d" % i: np.array([scale, P, sr]), "y%d" % i: y.numpy()})
    save("roi_align", n_cases=len(cases), **out)


def gen_psm():
    from maskrcnn_benchmark.modeling.roi_heads.box_head.loss import make_roi_box_loss_evaluator
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    c = cfg.clone()
    c.merge_from_list(MT_OPTS)
    ev = make_roi_box_loss_evaluator(c)
    g = torch.Generator().manual_seed(13)
    out = {}
    cases = [(1024, 4, 0.25), (512, 4, 0.05), (64, 2, 0.5), (100, 4, 0.9), (40, 4, 0.0)]
    for i, (n, views, pos_frac) in enumerate(cases):
        base = torch.randn(n, 3, generator=g) * 2
        teacher = [base + torch.randn(n, 3, generator=g) * (0.3 + torch.rand(n, 1, generator=g)) for _ in range(views)]
        student = (base + torch.randn(n, 3, generator=g)).requires_grad_(True)
        labels = torch.zeros(n, dtype=torch.int64)
        npos = int(n * pos_frac)
        labels[torch.randperm(n, generator=g)[:npos]] = torch.randint(1, 3, (npos,), generator=g)
        halves = [n // 2, n - n // 2]
        props, o = [], 0
        for h in halves:
            bl = BoxList(torch.zeros(h, 4), (800, 800))
            bl.add_field("labels", labels[o:o + h])
            props.append(bl)
            o += h
        loss = ev.evaluatePSM([student], teacher, props)
        if torch.isfinite(loss):
            loss.backward()
            grad = student.grad.numpy()
        else:
            grad = np.zeros((n, 3), np.float32)
        out.update({"student%d" % i: student.detach().numpy(), "teacher%d" % i: torch.stack(teacher).numpy(),
                    "labels%d" % i: labels.numpy(), "loss%d" % i: loss.detach().numpy(), "grad%d" % i: grad})
    save("psm", n_cases=len(cases), temperature=c.MT.TEMP, neg_weight=c.MT.CLS_BALANCE_WEIGHT, **out)


def gen_mgd():
    from maskrcnn_benchmark.modeling.detector.generalized_rcnn import fg_hint_loss
    g = torch.Generator().manual_seed(14)
    out = {}
    cases = [(2, 16, 64, 4), (1, 128, 96, 4), (2, 8, 64, 2)]
    for i, (B, C, S, V) in enumerate(cases):
        sizes = [S // 4, S // 8, S // 16, S // 32, (S // 32 + 1) // 2]
        student = tuple(torch.randn(B, C, s, s, generator=g).requires_grad_(True) for s in sizes)
        teachers = [tuple(torch.randn(B, C, s, s, generator=g) for s in sizes) for _ in range(V)]
        mask = torch.zeros(B, S, S, dtype=torch.uint8)
        for b in range(B):
            for _ in range(5):
                cy, cx = torch.randint(0, S, (2,), generator=g).tolist()
                r = int(torch.randint(3, S // 4, (1,), generator=g))
                mask[b, max(cy - r, 0):cy + r, max(cx - r, 0):cx + r] += 1
        loss = fg_hint_loss(teachers, [student], [m for m in mask])
        loss.backward()
        out.update({"mask%d" % i: mask.numpy(), "loss%d" % i: loss.detach().numpy(), "dims%d" % i: np.array([B, C, S, V])})
        for l, s in enumerate(student):
            out["student%d_%d" % (i, l)] = s.detach().numpy()
            out["grad%d_%d" % (i, l)] = s.grad.numpy()
            for v in range(V):
                out["teacher%d_%d_%d" % (i, v, l)] = teachers[v][l].numpy()
    save("mgd", n_cases=len(cases), **out)


def gen_ema_and_weights():
    from maskrcnn_benchmark.engine.MTtrainer import MTtrainer, weight_sum_losses
    g = torch.Generator().manual_seed(15)
    student = torch.nn.Linear(37, 53)
    teacher = torch.nn.Linear(37, 53)
    with torch.no_grad():
        for p in list(student.parameters()) + list(teacher.parameters()):
            p.copy_(torch.randn(p.shape, generator=g))
    out = {"s_w": student.weight.detach().numpy().copy(), "s_b": student.bias.detach().numpy().copy(),
           "t_w": teacher.weight.detach().numpy().copy(), "t_b": teacher.bias.detach().numpy().copy()}
    me = SimpleNamespace(teacher=teacher, student=student, alpha=0.99)
    steps = [0, 3, 99, 5000]
    for k in steps:
        MTtrainer.update_teacher(me, k)
        out["t_w_after%d" % k] = teacher.weight.detach().numpy().copy()
        out["t_b_after%d" % k] = teacher.bias.detach().numpy().copy()
    out["steps"] = np.array(steps)
    # loss weighting schedule (engine/MTtrainer.py:67-109)
    balanced = {"mt_classifier": 0.2, "nms_loss": 1.0, "mt_fg_loss": 1.0}
    ld = {"loss_classifier": 0.7, "loss_box_reg": 0.3, "loss_seg": 0.9, "mt_classifier": 0.11, "mt_fg_loss": 0.05, "nms_loss": 0.4}
    iters = [900, 1001, 1100, 1249, 1250, 5000, 9800, 9999]
    table = []
    for it in iters:
        w = weight_sum_losses(dict(ld), it, 250, 250, 10000, l=5.0, balanced=balanced, start_mt=1000)
        table.append([float(w[k]) for k in sorted(ld)])
    out["ws_iters"] = np.array(iters)
    out["ws_keys"] = np.array(sorted(ld))
    out["ws_in"] = np.array([ld[k] for k in sorted(ld)])
    out["ws_out"] = np.array(table)
    save("ema_weights", **out)


def gen_glue():
    """Host-side glue the CUDA ops sit between: anchors, box coder, matcher, level mapper, boxlist ops."""
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.matcher import Matcher
    from maskrcnn_benchmark.modeling.poolers import LevelMapper
    from maskrcnn_benchmark.modeling.rpn.anchor_generator import make_anchor_generator
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.boxlist_ops import boxlist_iou
    from maskrcnn_benchmark.structures.image_list import to_image_list
    g = torch.Generator().manual_seed(16)
    out = {}
    c = cfg.clone()
    ag = make_anchor_generator(c)
    feats = [torch.zeros(1, 1, s, s) for s in (16, 8, 4, 2, 1)]
    il = to_image_list(torch.zeros(1, 3, 64, 64))
    anchors = ag(il, feats)[0]
    for l, a in enumerate(anchors):
        out["anchors%d" % l] = a.bbox.numpy()
        out["anchors_vis%d" % l] = a.get_field("visibility").numpy()
    ref = clustered_boxes(g, 200)
    prop = ref + torch.randn(200, 4, generator=g) * 4
    prop[:, 2:] = torch.max(prop[:, 2:], prop[:, :2] + 1)
    for name, w in (("rpn", (1., 1., 1., 1.)), ("box", (10., 10., 5., 5.))):
        bc = BoxCoder(weights=w)
        enc = bc.encode(ref, prop)
        dec = bc.decode(enc * 1.3 + 0.01, prop)
        out["coder_%s_enc" % name] = enc.numpy()
        out["coder_%s_dec" % name] = dec.numpy()
    out["coder_ref"] = ref.numpy()
    out["coder_prop"] = prop.numpy()
    gt = BoxList(clustered_boxes(g, 12), (800, 800))
    pr = BoxList(torch.cat([clustered_boxes(g, 300), gt.bbox + 2]), (800, 800))
    iou = boxlist_iou(gt, pr)
    out["iou_gt"] = gt.bbox.numpy()
    out["iou_pr"] = pr.bbox.numpy()
    out["iou"] = iou.numpy()
    out["match_rpn"] = Matcher(0.7, 0.3, allow_low_quality_matches=True)(iou).numpy()
    out["match_box"] = Matcher(0.5, 0.5, allow_low_quality_matches=False)(iou).numpy()
    lm = LevelMapper(2, 5)
    bl = BoxList(torch.cat([clustered_boxes(g, 300), torch.tensor([[0, 0, 111, 111.], [0, 0, 223, 223.], [0, 0, 447, 447.], [0, 0, 799, 799.], [3, 3, 4, 4.]])]), (800, 800))
    out["lvl_boxes"] = bl.bbox.numpy()
    out["lvl"] = lm([bl]).numpy()
    save("glue", **out)


def synthetic_instances(g, n_obj, size=800.0):
    """n_obj instances: 16-gons (sometimes two polygons per instance), like the bench's synthetic targets."""
    inst = []
    for i in range(n_obj):
        polys = []
        for _ in range(2 if i % 5 == 4 else 1):
            c = torch.rand(2, generator=g) * size
            r = float(torch.rand(1, generator=g)) * 60 + 8
            ang = torch.arange(16, dtype=torch.float32) / 16 * 6.2831853 + float(torch.rand(1, generator=g))
            rad = r * (0.7 + 0.6 * torch.rand(16, generator=g))
            pts = torch.stack([c[0] + rad * torch.cos(ang), c[1] + rad * torch.sin(ang)], 1).clamp(0, size - 1)
            polys.append(pts.reshape(-1).tolist())
        inst.append(polys)
    return inst


def gen_mask_targets():
    from maskrcnn_benchmark.modeling.roi_heads.mask_head.loss import project_masks_on_boxes
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.segmentation_mask import SegmentationMask
    g = torch.Generator().manual_seed(17)
    inst = synthetic_instances(g, 24)
    sm = SegmentationMask(inst, (800, 800))
    K = 96
    which = torch.randint(0, len(inst), (K,), generator=g)
    boxes = []
    for i in which.tolist():
        pts = torch.tensor([v for p in inst[i] for v in p]).view(-1, 2)
        lo, hi = pts.min(0)[0], pts.max(0)[0]
        jit = (torch.rand(4, generator=g) - 0.5) * 0.4 * torch.cat([hi - lo, hi - lo])
        boxes.append(torch.cat([lo, hi]) + jit)
    boxes = torch.stack(boxes)
    boxes[3] = torch.tensor([100.2, 100.7, 100.9, 101.1])  # sub-pixel box (w,h < 1 -> clamped to 1)
    boxes[5] = boxes[5].round()                              # integer box with w == h possible
    boxes[5, 3] = boxes[5, 1] + (boxes[5, 2] - boxes[5, 0])
    masks = project_masks_on_boxes(sm[which], BoxList(boxes, (800, 800)), 28)
    flat = [np.asarray(p, np.float32) for polys in inst for p in polys]
    save("mask_targets", xy=np.concatenate(flat), poly_len=np.array([len(p) // 2 for p in flat]),
         inst_npoly=np.array([len(polys) for polys in inst]), boxes=boxes.numpy(), which=which.numpy(),
         masks=masks.numpy().astype(np.uint8))


def gen_step(size=256, n_cyto=6, n_nuc=6, name="step_small", **init_kw):
    """One supervised step + one mean-teacher forward (teacher -> student-unlabeled) of the REFERENCE model on CPU,
    from the deterministic synthetic weights / inputs of mmt_psm_b200 (DO = 0 so no dropout stream is involved)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.data.synthetic import make_batch
    from mmt_psm_b200.utils.synthetic_init import synthetic_state_dict
    from maskrcnn_benchmark.modeling.detector import build_detection_model
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.segmentation_mask import SegmentationMask
    c = cfg.clone()
    c.merge_from_file(os.path.join(TREE, "configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml"))
    c.merge_from_list(MT_OPTS + ["MODEL.RELATION_NMS.USE_RELATION_NMS", False, "MODEL.RELATION_MASK.USE_RELATION", False,
                                 "MODEL.ROI_BOX_HEAD.DO", 0.0])
    student = build_detection_model(c, is_student=True)
    teacher = build_detection_model(c, is_teacher=True)
    shapes = {k: tuple(v.shape) for k, v in student.state_dict().items() if "mask_relation_module" not in k}
    sd = synthetic_state_dict(shapes, 0, **init_kw)
    for m in (student, teacher):
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all("mask_relation_module" in k or "cell_anchors" in k for k in missing), missing
    images, targets, views = make_batch(1234, 2, size, 3, n_cyto, n_nuc)
    rt = []
    for t in targets:
        b = BoxList(t.bbox, t.size, "xyxy")
        b.add_field("labels", t.get_field("labels"))
        b.add_field("masks", SegmentationMask([[p.tolist() for p in inst.polygons] for inst in t.get_field("masks").polygons], t.size))
        rt.append(b)
    student.train()
    teacher.eval()
    torch.manual_seed(7)
    sup = student(images, rt)
    out = {"sup_" + k: v.detach().numpy() for k, v in sup.items()}
    with torch.no_grad():
        tr = teacher.forward_teacher([views[0].clone(), views[1].clone()])
    mt = student.forward_student([views[2].clone()], tr)
    out.update({"mt_" + k: v.detach().numpy() for k, v in mt.items()})
    out["t_n_rois"] = np.array([len(b) for b in tr["result_t"]])
    out["t_labels"] = torch.cat([b.get_field("labels") for b in tr["result_t"]]).numpy()
    out["t_boxes"] = torch.cat([b.bbox for b in tr["result_t"]]).numpy()
    out["t_logits"] = torch.stack(tr["class_logit_t"]).numpy()
    out["t_seg_mask_sum"] = np.array([float(m.sum()) for m in tr["seg_mask"]])
    out["t_seg_mask"] = np.packbits(torch.stack(tr["seg_mask"]).numpy() > 0)
    save(name, size=size, n_cyto=n_cyto, n_nuc=n_nuc, adaptor_scale=init_kw.get("adaptor_scale", 0.1),
         box_reg_gain=init_kw.get("box_reg_gain", 1.0), **out)
    print({k: float(v) for k, v in out.items() if v.ndim == 0})


def gen_color_views():
    """K + S colour views of one unlabeled crop through the REFERENCE's own transform classes
    (data/transforms/transforms.py, `transpart2` of data/transforms/build.py:27-35), seeded python / numpy streams."""
    import importlib.util
    import random
    from PIL import Image
    spec = importlib.util.spec_from_file_location("ref_transforms", os.path.join(TREE, "maskrcnn_benchmark/data/transforms/transforms.py"))
    T = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(T)          # by path: the package __init__ pulls openslide / slidingwindow
    g = np.random.default_rng(42)
    img = g.integers(0, 256, (72, 96, 3), dtype=np.uint8)
    img[10:30, 20:60] = (img[10:30, 20:60] // 8) + 200          # a bright, low-saturation patch
    img[40:60, 5:40, 1] = 0                                      # a saturated patch
    part2 = T.Compose([T.AdjustBrightness(0.15), T.AdjustContrast(0.15), T.AdjustHue(0.05), T.RandomErasing(0.9),
                       T.ToTensor(), T.Normalize(mean=[102.9801, 115.9465, 122.7717], std=[1., 1., 1.], to_bgr255=True)])
    random.seed(5)
    np.random.seed(5)
    views = []
    for _ in range(3):
        out, _ = part2(Image.fromarray(img.copy(), "RGB"), None)
        views.append(out.numpy())
    save("color_views", image=img, views=np.stack(views), seed=5)


def relation_cfg():
    """cfg of configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml + scripts/train_mt.sh with both IR-Net modules on (BASELINE config 4)."""
    c = cfg.clone()
    c.merge_from_file(os.path.join(TREE, "configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml"))
    c.merge_from_list(MT_OPTS + ["MODEL.RELATION_NMS.USE_RELATION_NMS", True, "MODEL.RELATION_MASK.USE_RELATION", True,
                                 "MODEL.RELATION_NMS.REG_IOU", True, "MODEL.RELATION_NMS.REG_IOU_MSK", False,
                                 "MODEL.RELATION_NMS.LOSS", 1., "MODEL.ROI_BOX_HEAD.DO", 0.0])
    return c


def _relation_inputs(g, n, n_gt, size):
    """ROIs scattered around n_gt ground-truth boxes (several near-duplicates per object), head outputs and RPN objectness."""
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    ctr = torch.rand(n_gt, 2, generator=g) * (size - 80) + 40
    wh = torch.rand(n_gt, 2, generator=g) * 50 + 14
    gt = torch.cat([ctr - wh / 2, ctr + wh / 2], 1).clamp(0, size - 1)
    gl = torch.randint(1, 3, (n_gt,), generator=g)
    own = torch.randint(0, n_gt, (n,), generator=g)
    rois = (gt[own] + torch.randn(n, 4, generator=g) * 5).clamp(0, size - 1)
    rois[:, 2:] = torch.max(rois[:, 2:], rois[:, :2] + 4)
    far = torch.rand(n, generator=g) < 0.25                       # a quarter of the ROIs are background
    rois[far] = clustered_boxes(g, int(far.sum()), float(size))
    logits = torch.randn(n, 3, generator=g)
    logits[~far, 0] -= 1.5
    logits[torch.arange(n)[~far], gl[own][~far]] += 2.0
    reg = torch.randn(n, 12, generator=g) * 0.4
    x = torch.relu(torch.randn(n, 1024, generator=g)).half().float()   # stored as fp16
    obj = torch.rand(n, generator=g)
    props = BoxList(rois, (size, size))
    props.add_field("objectness", obj)
    target = BoxList(gt, (size, size))
    target.add_field("labels", gl)
    return x, props, torch.softmax(logits, 1), reg, target


def gen_relation():
    """The reference's IR-Net modules on their own (modeling/relation/relation_module.py DuplicationRemovalNetwork,
    mask_relation_module.py MaskRelationRefineNet), train and eval mode, synthetic weights keyed like the detector's."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.utils.synthetic_init import synthetic_state_dict
    from maskrcnn_benchmark.modeling.relation.relation_module import DuplicationRemovalNetwork
    from maskrcnn_benchmark.modeling.relation.mask_relation_module import MaskRelationRefineNet
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    c = relation_cfg()
    out = {}
    g = torch.Generator().manual_seed(77)

    net = DuplicationRemovalNetwork(c)
    pre = "relation_nms."
    sd = synthetic_state_dict({pre + k: tuple(v.shape) for k, v in net.state_dict().items()}, 0)
    net.load_state_dict({k[len(pre):]: v for k, v in sd.items()})
    for tag, n, n_gt in (("a", 120, 14), ("b", 50, 9)):          # more / fewer ROIs than FIRST_N = 90
        x, props, prob, reg, target = _relation_inputs(g, n, n_gt, 256)
        if tag == "b":
            target = target[target.get_field("labels") == 1]     # one class without any ground truth
        for k, v in (("x", x.half()), ("rois", props.bbox), ("obj", props.get_field("objectness")), ("prob", prob), ("reg", reg),
                     ("gt", target.bbox), ("gt_labels", target.get_field("labels"))):
            out["nms_%s_%s" % (tag, k)] = v.numpy()
        net.train()
        xr = x.clone().requires_grad_(True)
        _, loss = net((xr, [props], prob, reg, [target]))
        loss["nms_loss"].backward()
        out["nms_%s_loss" % tag] = loss["nms_loss"].detach().numpy()
        out["nms_%s_dx" % tag] = xr.grad.numpy()
        out["nms_%s_dclassifier" % tag] = net.classifier.weight.grad.clone().numpy()
        out["nms_%s_dWG" % tag] = net.relation_module.WG.weight.grad.clone().numpy()
        net.zero_grad()
        with torch.no_grad():
            ranked = net.prepare_ranking(prob, reg, [props], [target], reg_iou=True)[0]
        out["nms_%s_sorted_idx" % tag] = ranked.get_field("sorted_idx").numpy()
        out["nms_%s_reg_label" % tag] = ranked.get_field("labels_iou_reg").numpy()
        net.eval()
        with torch.no_grad():
            res, _ = net((x, [props], prob, reg, [None]))
        r = res[0]
        for f in ("scores", "objectness", "all_scores", "labels"):
            out["nms_%s_det_%s" % (tag, f)] = r.get_field(f).numpy()
        out["nms_%s_det_bbox" % tag] = r.bbox.numpy()
        print("relation nms", tag, float(loss["nms_loss"]), "detections", len(r))

    mnet = MaskRelationRefineNet(c, None)
    pre = "mask_heads.mask.mask_relation_module."
    sd = synthetic_state_dict({pre + k: tuple(v.shape) for k, v in mnet.state_dict().items()}, 0)
    mnet.load_state_dict({k[len(pre):]: v for k, v in sd.items()})
    mnet.train()
    for tag, P in (("a", 14), ("b", 4)):
        f = torch.relu(torch.randn(P, 256, 14, 14, generator=g)).half().float()   # stored as fp16
        m = torch.randn(P, 3, 28, 28, generator=g) * 2
        boxes = clustered_boxes(g, P, 256.0)
        labels = torch.randint(1, 3, (P,), generator=g) if tag == "a" else torch.full((P,), 2, dtype=torch.int64)
        obj = torch.rand(P, generator=g)
        obj[::4] = 1.0                                            # ground-truth boxes join the proposals with objectness 1
        p = BoxList(boxes, (256, 256))
        p.add_field("labels", labels)
        p.add_field("objectness", obj)
        fr = f.clone().requires_grad_(True)
        mr = m.clone().requires_grad_(True)
        logits2, sp, _, _ = mnet((fr, mr, p, None))
        wsum = torch.randn(logits2.shape, generator=g)
        (logits2 * wsum).sum().backward()
        for k, v in (("f", f.half()), ("m", m), ("boxes", boxes), ("labels", labels), ("obj", obj), ("logits2", logits2.detach()),
                     ("sorted_boxes", sp[0].bbox), ("sorted_labels", sp[0].get_field("labels")), ("w", wsum),
                     ("df_sum", fr.grad.sum((2, 3))), ("dm", mr.grad), ("dgamma", mnet.relation_module.gamma.grad.clone())):
            out["msk_%s_%s" % (tag, k)] = v.numpy()
        mnet.zero_grad()
        print("relation mask", tag, tuple(logits2.shape), float(logits2.abs().mean()))
    save("relation", **out)


def gen_step_relation():
    """As gen_step with both IR-Net modules on (BASELINE config 4 = what the reference's yaml + train_mt.sh select):
    supervised step with nms_loss and the deeply-supervised refined mask loss, teacher whose coarse inference goes through
    the duplicate-removal network and the mask refinement, student-unlabeled losses."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.data.synthetic import make_batch
    from mmt_psm_b200.utils.synthetic_init import synthetic_state_dict
    from maskrcnn_benchmark.modeling.detector import build_detection_model
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.segmentation_mask import SegmentationMask
    c = relation_cfg()
    student = build_detection_model(c, is_student=True)
    teacher = build_detection_model(c, is_teacher=True)
    sd = synthetic_state_dict({k: tuple(v.shape) for k, v in student.state_dict().items()}, 0)
    for m in (student, teacher):
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all("cell_anchors" in k for k in missing), missing
    images, targets, views = make_batch(1234, 2, size, 3, n_cyto, n_nuc)
    rt = []
    for t in targets:
        b = BoxList(t.bbox, t.size, "xyxy")
        b.add_field("labels", t.get_field("labels"))
        b.add_field("masks", SegmentationMask([[p.tolist() for p in inst.polygons] for inst in t.get_field("masks").polygons], t.size))
        rt.append(b)
    student.train()
    teacher.eval()
    torch.manual_seed(7)
    sup = student(images, rt)
    out = {"sup_" + k: v.detach().numpy() for k, v in sup.items()}
    with torch.no_grad():
        infer = teacher(views[0].clone())
        out["infer_n"] = np.array([len(b) for b in infer])
        out["infer_labels"] = torch.cat([b.get_field("labels") for b in infer]).numpy()
        out["infer_boxes"] = torch.cat([b.bbox for b in infer]).numpy()
        out["infer_scores"] = torch.cat([b.get_field("scores") for b in infer]).numpy()
        torch.manual_seed(11)
        tr = teacher.forward_teacher([views[0].clone(), views[1].clone()])
    mt = student.forward_student([views[2].clone()], tr)
    out.update({"mt_" + k: v.detach().numpy() for k, v in mt.items()})
    out["t_n_rois"] = np.array([len(b) for b in tr["result_t"]])
    out["t_labels"] = torch.cat([b.get_field("labels") for b in tr["result_t"]]).numpy()
    out["t_boxes"] = torch.cat([b.bbox for b in tr["result_t"]]).numpy()
    out["t_logits"] = torch.stack(tr["class_logit_t"]).numpy()
    out["t_seg_mask_sum"] = np.array([float(m.sum()) for m in tr["seg_mask"]])
    out["t_seg_mask"] = np.packbits(torch.stack(tr["seg_mask"]).numpy() > 0)
    save("step_relation", size=size, **out)
    print({k: float(v) for k, v in out.items() if v.ndim == 0}, out["infer_n"], out["t_n_rois"])


def gen_source_items():
    """Labeled (`source`) items through the REFERENCE's own transform classes and containers: data/transforms/build.py:37-50
    (Resize, RandomHorizontalFlip, AdjustBrightness / Contrast / Hue, RandomErasing(0.7), ToTensor, Normalize) applied to a
    PIL crop and a BoxList target with polygon masks (structures/bounding_box.py, segmentation_mask.py)."""
    import importlib.util
    import random
    from PIL import Image
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.segmentation_mask import SegmentationMask
    spec = importlib.util.spec_from_file_location("ref_transforms", os.path.join(TREE, "maskrcnn_benchmark/data/transforms/transforms.py"))
    T = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(T)
    tr = T.Compose([T.Resize(64, 120), T.RandomHorizontalFlip(0.5), T.AdjustBrightness(0.15), T.AdjustContrast(0.15),
                    T.AdjustHue(0.05), T.RandomErasing(0.7), T.ToTensor(),
                    T.Normalize(mean=[102.9801, 115.9465, 122.7717], std=[1., 1., 1.], to_bgr255=True)])
    g = np.random.default_rng(21)
    tg = torch.Generator().manual_seed(21)
    out = {"seed": 9, "n": 4}
    random.seed(9)
    np.random.seed(9)
    for i, (h, w) in enumerate(((100, 120), (90, 150), (128, 96), (100, 120))):
        img = g.integers(0, 256, (h, w, 3), dtype=np.uint8)
        inst = synthetic_instances(tg, 5, float(min(h, w)))
        boxes = []
        for polys in inst:
            xs = np.concatenate([np.asarray(p[0::2]) for p in polys])
            ys = np.concatenate([np.asarray(p[1::2]) for p in polys])
            boxes.append([xs.min(), ys.min(), xs.max(), ys.max()])
        labels = torch.randint(1, 3, (5,), generator=tg)
        t = BoxList(torch.tensor(boxes, dtype=torch.float32), (w, h), "xyxy")
        t.add_field("labels", labels)
        t.add_field("masks", SegmentationMask(inst, (w, h)))
        image, t2 = tr(Image.fromarray(img, "RGB"), t)
        out["img_%d" % i] = img
        out["boxes_%d" % i] = np.asarray(boxes, dtype=np.float32)
        out["labels_%d" % i] = labels.numpy()
        out["poly_%d" % i] = np.array([len(polys) for polys in inst])
        out["poly_flat_%d" % i] = np.concatenate([np.asarray(p, dtype=np.float64) for polys in inst for p in polys])
        out["poly_len_%d" % i] = np.array([len(p) for polys in inst for p in polys])
        out["out_image_%d" % i] = image.numpy()
        out["out_boxes_%d" % i] = t2.bbox.numpy()
        out["out_size_%d" % i] = np.array(t2.size)
        out["out_poly_flat_%d" % i] = np.concatenate([p.numpy().astype(np.float64) for m in t2.get_field("masks").polygons for p in m.polygons])
    save("source_items", **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    which = sys.argv[1:] or ["nms", "roi_align", "psm", "mgd", "ema", "glue", "mask_targets", "step", "color_views"]
    if "nms" in which: gen_nms()
    if "roi_align" in which: gen_roi_align()
    if "psm" in which: gen_psm()
    if "mgd" in which: gen_mgd()
    if "ema" in which: gen_ema_and_weights()
    if "glue" in which: gen_glue()
    if "mask_targets" in which: gen_mask_targets()
    if "step" in which: gen_step()
    if "step_full" in which: gen_step(800, 20, 20, "step_full", adaptor_scale=0.4, box_reg_gain=40.0)   # the benchmarked size: 800x800, 40 objects / image; init gains so that loss_box_reg and mt_fg_loss are O(0.1)
    if "color_views" in which: gen_color_views()
    if "source_items" in which: gen_source_items()
    if "relation" in which: gen_relation()
    if "step_relation" in which: gen_step_relation()
