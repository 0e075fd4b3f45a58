"""GPU suite, IR-Net relation modules (BASELINE config 4): this package's DuplicationRemovalNetwork / MaskRelationRefineNet
and the detector step with both switched on, against
  (1) tests/golden/relation.npz and step_relation.npz — outputs of the PATCHED REFERENCE's own modules / model on the same
      weights and inputs (tests/golden/make_golden.py relation, step_relation), and
  (2) the oracle's label preparation (oracle/relation_ref.py, numpy restatement of relation_module.py:333-413) on larger
      random cases.
fp32 conv math isolates parity from TF32 operand rounding; the shipped TF32 default is checked at a looser tolerance."""
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden

sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def _t(g, k, cuda):
    a = g[k]
    return torch.from_numpy(a.astype(np.float32) if a.dtype == np.float16 else a).to(cuda)


def _load_reference_layout(module, sd, prefix):
    """Fill a stand-alone module from a reference-layout dict (OIHW convs, IOHW deconv) under `prefix`."""
    from mmt_psm_b200.layers.conv import Conv2d, ConvTranspose2x2
    params, mods = dict(module.named_parameters()), dict(module.named_modules())
    assert set(params) == {k[len(prefix):] for k in sd}, set(params) ^ {k[len(prefix):] for k in sd}
    with torch.no_grad():
        for k, v in sd.items():
            k = k[len(prefix):]
            owner = mods.get(k.rsplit(".", 1)[0])
            v = v.to(params[k].device)
            if k.endswith(".weight") and isinstance(owner, (Conv2d, ConvTranspose2x2)):
                owner.load_reference_weight(v)
            else:
                params[k].copy_(v)


@pytest.fixture
def restore():
    from mmt_psm_b200 import kernels
    from mmt_psm_b200.modeling.balanced_positive_negative_sampler import BalancedPositiveNegativeSampler
    yield
    kernels.set_conv_math("tf32")
    BalancedPositiveNegativeSampler.rng_device = None


def _cfg(**kw):
    from mmt_psm_b200.config.presets import mt_psm_cfg
    return mt_psm_cfg(start_mt=0, relation=True, extra=["MODEL.ROI_BOX_HEAD.DO", 0.0], **kw)


def _nms_case(g, tag, cuda):
    from mmt_psm_b200.structures.bounding_box import BoxList
    a = {k: _t(g, "nms_%s_%s" % (tag, k), cuda) for k in ("x", "prob", "reg", "rois", "obj", "gt", "gt_labels")}
    props = BoxList(a["rois"], (256, 256))
    props.add_field("objectness", a["obj"])
    target = BoxList(a["gt"], (256, 256))
    target.add_field("labels", a["gt_labels"])
    return a, props, target


@pytest.mark.parametrize("math,tol", [("fp32", 2e-4), ("tf32", 2e-2)])
def test_duplicate_removal_network_matches_reference(cuda, restore, math, tol):
    """Ranking, IoU-regression targets, nms_loss + gradients, and the eval-mode detections of the reference's
    DuplicationRemovalNetwork (relation_module.py:94-331) on the golden inputs; more / fewer ROIs than FIRST_N and a class
    without ground truth."""
    from mmt_psm_b200 import kernels, lib
    from mmt_psm_b200.modeling.relation import DuplicationRemovalNetwork
    from mmt_psm_b200.utils.synthetic_init import synthetic_state_dict
    from oracle import relation_ref as RR
    kernels.set_conv_math(math)
    g = golden("relation")
    net = DuplicationRemovalNetwork(_cfg()).to(cuda)
    _load_reference_layout(net, synthetic_state_dict(RR.nms_weight_shapes(), 0), RR.NMS_PREFIX)
    lib.reset_launch_count()
    for tag in "ab":
        a, props, target = _nms_case(g, tag, cuda)
        net.train()
        net.zero_grad()
        ranked = net._rank(a["prob"], a["reg"], props, target)
        assert np.array_equal(ranked["idx"].cpu().numpy(), g["nms_%s_sorted_idx" % tag])
        ref_lab = g["nms_%s_reg_label" % tag]
        lab = ranked["labels"].cpu().numpy()
        assert np.array_equal(lab > 0, ref_lab > 0)                        # the same boxes are the ground truths' winners
        np.testing.assert_allclose(lab, ref_lab, rtol=0, atol=2e-6)        # IoU values: fp32 division on GPU vs CPU
        x = a["x"].clone().requires_grad_(True)
        _, loss = net((x, [props], a["prob"], a["reg"], [target]))
        loss["nms_loss"].backward()
        ref = float(g["nms_%s_loss" % tag])
        assert abs(float(loss["nms_loss"]) - ref) <= tol * ref, (math, tag, float(loss["nms_loss"]), ref)
        for got, name in ((x.grad, "dx"), (net.relation_module.WG.weight.grad, "dWG"), (net.classifier.weight.grad, "dclassifier")):
            r = g["nms_%s_%s" % (tag, name)]
            err = np.abs(got.cpu().numpy() - r).max() / np.abs(r).max()
            assert err <= 5 * tol, (math, tag, name, err)
        net.eval()
        with torch.no_grad():
            res, _ = net((a["x"], [props], a["prob"], a["reg"], [None]))
        det = res[0]
        rb, rs, rl = g["nms_%s_det_bbox" % tag], g["nms_%s_det_scores" % tag], g["nms_%s_det_labels" % tag]
        if math == "fp32":
            assert len(det) == len(rl)
            assert np.array_equal(det.get_field("labels").cpu().numpy(), rl)            # nuclei first, then cytoplasm
            np.testing.assert_allclose(det.bbox.cpu().numpy(), rb, rtol=1e-5, atol=1e-3)
            np.testing.assert_allclose(det.get_field("scores").cpu().numpy(), rs, rtol=1e-3, atol=1e-4)
            np.testing.assert_allclose(det.get_field("objectness").cpu().numpy(), g["nms_%s_det_objectness" % tag], rtol=1e-6)
            np.testing.assert_allclose(det.get_field("all_scores").cpu().numpy(), g["nms_%s_det_all_scores" % tag], rtol=1e-6)
        else:
            assert abs(len(det) - len(rl)) <= 3                                        # scores next to FG_THREAD may flip
    # both images in ONE call (different N -> two attention groups) = the reference's per-image loop
    net.train()
    cases = [_nms_case(g, tag, cuda) for tag in "ab"]
    _, loss = net((torch.cat([c[0]["x"] for c in cases]), [c[1] for c in cases], torch.cat([c[0]["prob"] for c in cases]),
                   torch.cat([c[0]["reg"] for c in cases]), [c[2] for c in cases]))
    ref = 0.5 * (float(g["nms_a_loss"]) + float(g["nms_b_loss"]))
    assert abs(float(loss["nms_loss"]) - ref) <= tol * ref
    assert lib.launch_count() > 10


def test_iou_regression_targets_match_numpy_label_preparation(cuda):
    """Device label preparation against the restated numpy argmax / intersect1d logic (oracle reg_labels) on 40 random
    cases at the path's sizes (90 ranked boxes x 2 classes, up to 40 ground-truth boxes, three thresholds), including
    classes without ground truth, ground truths nobody overlaps (rank-0 fallback) and near-duplicate boxes."""
    from mmt_psm_b200.modeling.relation import iou_regression_targets
    from oracle import relation_ref as RR
    gen = torch.Generator().manual_seed(5)
    hits = 0
    for case in range(40):
        M = int(torch.randint(0, 41, (1,), generator=gen))
        N = 90 if case % 4 else int(torch.randint(1, 90, (1,), generator=gen))
        ctr = torch.rand(M, 2, generator=gen) * 700 + 50
        wh = torch.rand(M, 2, generator=gen) * 80 + 10
        gt = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        gl = torch.randint(1, 3, (M,), generator=gen) if case % 5 else torch.ones(M, dtype=torch.int64)
        if M:
            own = torch.randint(0, M, (N, 2), generator=gen)
            boxes = gt[own] + torch.randn(N, 2, 4, generator=gen) * (3 + case % 7)
            far = torch.rand(N, 2, generator=gen) < 0.3
            boxes[far] = torch.rand(int(far.sum()), 4, generator=gen) * 100
        else:
            boxes = torch.rand(N, 2, 4, generator=gen) * 300
        boxes[..., 2:] = torch.max(boxes[..., 2:], boxes[..., :2] + 2)
        scores = torch.sort(torch.rand(N, 2, generator=gen), 0, descending=True)[0]
        for binary in (False, True):
            ref = RR.reg_labels(boxes, scores, gt, gl, (0.1, 0.5, 0.7), binary=binary)
            got = iou_regression_targets(boxes.to(cuda), scores.to(cuda), gt.to(cuda), gl.to(cuda), (0.1, 0.5, 0.7), binary=binary)
            got = got.cpu().numpy()
            assert np.array_equal(got > 0, ref > 0), (case, binary)
            np.testing.assert_allclose(got, ref, rtol=0, atol=2e-6)
        hits += int((ref > 0).sum())
    assert hits > 200


@pytest.mark.parametrize("math,tol", [("fp32", 2e-4), ("tf32", 3e-2)])
def test_mask_relation_refine_net_matches_reference(cuda, restore, math, tol):
    """Refined logits, row order and gradients of the reference's MaskRelationRefineNet (CIAM) on the golden inputs; then
    two images in one pass (masked attention) against the per-image passes."""
    from mmt_psm_b200 import kernels
    from mmt_psm_b200.modeling.relation import MaskRelationRefineNet
    from mmt_psm_b200.structures.bounding_box import BoxList
    from mmt_psm_b200.utils.synthetic_init import synthetic_state_dict
    from oracle import relation_ref as RR
    kernels.set_conv_math(math)
    g = golden("relation")
    net = MaskRelationRefineNet(_cfg(), None).to(cuda)
    _load_reference_layout(net, synthetic_state_dict(RR.mask_weight_shapes(), 0), RR.MASK_PREFIX)
    net.train()
    held = {}
    for tag in "ab":
        f = _t(g, "msk_%s_f" % tag, cuda).permute(0, 2, 3, 1).contiguous().requires_grad_(True)
        m = _t(g, "msk_%s_m" % tag, cuda).permute(0, 2, 3, 1).contiguous().requires_grad_(True)
        p = BoxList(_t(g, "msk_%s_boxes" % tag, cuda), (256, 256))
        p.add_field("labels", _t(g, "msk_%s_labels" % tag, cuda))
        p.add_field("objectness", _t(g, "msk_%s_obj" % tag, cuda))
        held[tag] = (f.detach(), m.detach(), p)
        net.zero_grad()
        logits, sp = net.refine(f, m, [p])
        # rows with equal objectness (the ground-truth boxes among the proposals, objectness 1) may come in any order:
        # align with the fixture by box
        gb, mb = g["msk_%s_sorted_boxes" % tag], sp[0].bbox.cpu().numpy()
        perm = [int(np.where((gb == b).all(1))[0][0]) for b in mb]
        assert sorted(perm) == list(range(len(gb)))
        sl = sp[0].get_field("labels").cpu().numpy()
        assert np.array_equal(sl, g["msk_%s_sorted_labels" % tag][perm]) and np.all(np.diff(sl) >= 0)
        ref = torch.from_numpy(g["msk_%s_logits2" % tag][perm]).permute(0, 2, 3, 1)
        w = torch.from_numpy(g["msk_%s_w" % tag][perm]).permute(0, 2, 3, 1).to(cuda)
        (logits * w).sum().backward()
        err = float((logits.detach().cpu() - ref).abs().max() / ref.abs().max())
        assert err <= tol, (math, tag, err)
        for got, name in ((m.grad.permute(0, 3, 1, 2), "dm"), (f.grad.sum((1, 2)), "df_sum"), (net.relation_module.gamma.grad, "dgamma")):
            r = g["msk_%s_%s" % (tag, name)]
            e = np.abs(got.cpu().numpy() - r).max() / np.abs(r).max()
            assert e <= 10 * tol, (math, tag, name, e)
    with torch.no_grad():
        (fa, ma, pa), (fb, mb_, pb) = held["a"], held["b"]
        both, sp = net.refine(torch.cat([fa, fb]), torch.cat([ma, mb_]), [pa, pb])
        la, _ = net.refine(fa, ma, [pa])
        lb, _ = net.refine(fb, mb_, [pb])
    assert [len(s) for s in sp] == [len(pa), len(pb)]
    assert float((both - torch.cat([la, lb])).abs().max()) <= 1e-4 * float(la.abs().max())


def _build(cuda):
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.data.synthetic import make_batch
    from mmt_psm_b200.modeling.balanced_positive_negative_sampler import BalancedPositiveNegativeSampler
    from mmt_psm_b200.modeling.detector import build_detection_model
    from mmt_psm_b200.utils.synthetic_init import init_model
    BalancedPositiveNegativeSampler.rng_device = "cpu"   # consume the reference's CPU random stream
    cfg = _cfg()
    student = build_detection_model(cfg, is_student=True).to(cuda)
    teacher = build_detection_model(cfg, is_teacher=True).to(cuda)
    init_model(student, 0)
    teacher.load_state_dict(student.state_dict())
    student.train()
    teacher.eval()
    g = golden("step_relation")
    images, targets, views = make_batch(1234, 2, int(g["size"]), 3, 6, 6)
    return cfg, student, teacher, images, targets, views, g


@pytest.mark.parametrize("math,tol_sup,tol_mt", [("fp32", 5e-4, 2e-3), ("tf32", 3e-2, 8e-2)])
def test_step_with_relation_matches_reference_golden(cuda, restore, math, tol_sup, tol_mt):
    """One supervised forward + teacher + student-unlabeled forward with RELATION_NMS and RELATION_MASK on (the reference's
    own yaml + train_mt.sh) against the patched reference model's numbers (step_relation.npz)."""
    from mmt_psm_b200 import kernels
    kernels.set_conv_math(math)
    cfg, student, teacher, images, targets, views, g = _build(cuda)
    assert len(student.state_dict()) == 349          # every reference state-dict entry is built
    seen = {}
    hook = student.relation_nms.register_forward_pre_hook(lambda mod, inp: seen.update(x=inp[0]))
    torch.manual_seed(7)
    sup = student(images.to(cuda), [t.to(cuda) for t in targets])
    hook.remove()
    for k in ("loss_classifier", "loss_box_reg", "loss_seg", "loss_objectness", "loss_rpn_box_reg"):
        ref, got = float(g["sup_" + k]), float(sup[k])
        assert abs(got - ref) <= tol_sup * max(abs(ref), 1e-2), (math, k, got, ref)
    # nms_loss hangs on DISCRETE choices made from near-tied numbers: the 90 best of 512 ROIs per class by a softmax score
    # whose neighbours are an fp32 ulp apart (random-init classifier), and per ground truth the best-scored overlapping box
    # — one flipped pair moves the loss by ~0.5 %.  So: (a) against the reference's number at a tolerance that allows a few
    # flips, and (b) tightly against the oracle evaluated on the very tensors the module was handed in this run.
    ref, got = float(g["sup_nms_loss"]), float(sup["nms_loss"])
    assert abs(got - ref) <= (3e-2 if math == "fp32" else 1.5e-1) * ref, (math, "nms_loss", got, ref)
    from mmt_psm_b200.modeling.detector import reference_state_dict
    from oracle import relation_ref as RR
    x, props, prob, reg, tg = seen["x"]
    sd = {k: v.cpu() for k, v in reference_state_dict(student).items() if k.startswith(RR.NMS_PREFIX)}
    n = [len(p) for p in props]
    per = [RR.relation_nms_loss(sd, RR.NMS_PREFIX, RR.RelationCfg(), xi.detach().cpu(), pi.detach().cpu(), ri.detach().cpu(),
                                p.bbox.cpu(), p.get_field("objectness").cpu(), p.size, t.bbox.cpu(), t.get_field("labels").cpu())[0]
           for xi, pi, ri, p, t in zip(x.split(n), prob.split(n), reg.split(n), props, tg)]
    same_inputs = float(torch.mean(torch.stack(per)))
    assert abs(got - same_inputs) <= (2e-3 if math == "fp32" else 5e-2) * same_inputs, (math, got, same_inputs)
    sum(sup.values()).backward()                     # the relation branches are differentiable end to end
    for name in ("relation_nms.relation_module.WK.weight", "relation_nms.nms_rank_fc.weight",
                 "mask_heads.mask.mask_relation_module.appearance_feature_extractor.mask_fcn1.weight",
                 "mask_heads.mask.mask_relation_module.relation_module.gamma"):
        p = dict(student.named_parameters())[name]
        assert p.grad is not None and float(p.grad.abs().sum()) > 0, name
    with torch.no_grad():
        infer = teacher(views[0].to(cuda))
        torch.manual_seed(11)
        tr = teacher.forward_teacher([views[0].to(cuda), views[1].to(cuda)])
    n_det = np.array([len(b) for b in infer])
    seg = np.packbits(torch.stack(list(tr["seg_mask"])).cpu().numpy() > 0)
    diff = (np.unpackbits(seg) != np.unpackbits(g["t_seg_mask"])).mean()
    assert np.array_equal(np.array([len(b) for b in tr["result_t"]]), g["t_n_rois"])
    if math == "fp32":
        assert np.abs(n_det - g["infer_n"]).max() <= 1, (n_det, g["infer_n"])
        if np.array_equal(n_det, g["infer_n"]):
            # detections come out in the refinement's order (class, objectness); equal-objectness rows may swap
            labels = torch.cat([b.get_field("labels") for b in infer]).cpu().numpy()
            assert np.array_equal(labels, g["infer_labels"])
            boxes = torch.cat([b.bbox for b in infer]).cpu().numpy()
            same = np.isclose(boxes, g["infer_boxes"], rtol=1e-4, atol=5e-3).all(1)
            assert same.mean() >= 0.95, same.mean()
        assert diff < 2e-3, diff
    else:
        assert np.abs(n_det - g["infer_n"]).max() <= 6, (n_det, g["infer_n"])
        assert diff < 0.04, diff
    mt = student.forward_student([views[2].to(cuda)], tr)
    for k in ("mt_fg_loss", "mt_classifier"):
        ref, got = float(g["mt_" + k]), float(mt[k])
        assert abs(got - ref) <= tol_mt * abs(ref), (math, k, got, ref)


def test_full_training_step_with_relation_matches_oracle(cuda, restore):
    """backward + fused SGD + EMA with the IR-Net branches on: after one mean-teacher iteration the updated student and
    teacher weights — the relation modules' included — equal the oracle trainer's (fp32 conv math; tolerance 2e-3 of
    each tensor's max-abs update, 2e-6 absolute floor: tensors whose whole update is ~2e-5 — a near-zero gradient — sit at
    a few fp32 ulps of the weights themselves, and the order in which concurrent streams add their partial weight
    gradients moves them by one ulp)."""
    from mmt_psm_b200 import kernels
    from mmt_psm_b200.engine import MTtrainer
    from mmt_psm_b200.modeling.detector import reference_state_dict
    from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer
    from oracle import model_cpu
    kernels.set_conv_math("fp32")
    cfg, student, teacher, images, targets, views, g = _build(cuda)
    opt = make_optimizer(cfg, student, teacher)
    trainer = MTtrainer(student, teacher, None, opt, make_lr_scheduler(cfg, opt), None, None, 10 ** 9, cfg)
    before = {k: v.clone() for k, v in reference_state_dict(student).items()}
    torch.manual_seed(7)
    ld, total = trainer.train_step(5, images, targets, views)
    assert "nms_loss" in ld

    oc = model_cpu.Cfg()
    oc.dropout, oc.start_mt, oc.relation = 0.0, 0, True
    ot = model_cpu.Trainer({k: v.cpu() for k, v in before.items()}, oc)
    tg = [dict(bbox=t.bbox, labels=t.get_field("labels"),
               polygons=[[p.numpy() for p in inst.polygons] for inst in t.get_field("masks").polygons]) for t in targets]
    torch.manual_seed(7)
    old, ototal = ot.step(5, images, tg, views)
    for k, v in old.items():
        assert abs(float(ld[k]) - float(v)) <= 2e-3 * max(abs(float(v)), 1e-2), (k, float(ld[k]), float(v))
    after_s = reference_state_dict(student)
    after_t = reference_state_dict(teacher)
    n_rel = 0
    for name in ot.student.trainable:
        ref = ot.student.sd[name].detach()
        got = after_s[name].cpu()
        upd = (ref - before[name].cpu()).abs().max().item()
        err = (got - ref).abs().max().item()
        assert err <= 2e-3 * upd + 2e-6, (name, err, upd)
        terr = (after_t[name].cpu() - ot.teacher.sd[name].detach()).abs().max().item()
        assert terr <= 2e-3 * upd + 2e-6, ("teacher", name, terr, upd)
        if "relation" in name:
            n_rel += 1
            # every relation parameter receives a gradient in this step — except WK's bias: a shift common to all keys
            # leaves every softmax row unchanged, its gradient is zero up to rounding
            assert upd > 0 or name.endswith("WK.bias"), name
    assert n_rel == 27


def test_graphed_iterations_with_relation_stay_finite(cuda, restore):
    """Five mean-teacher iterations of BASELINE config 4 through the CUDA-graph trunks (the shipped configuration, TF32):
    every loss term stays finite and the teacher's relation weights follow the student's by EMA."""
    from mmt_psm_b200.engine import MTtrainer
    from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer
    cfg, student, teacher, images, targets, views, g = _build(cuda)
    opt = make_optimizer(cfg, student, teacher)
    trainer = MTtrainer(student, teacher, None, opt, make_lr_scheduler(cfg, opt), None, None, 10 ** 9, cfg, cuda_graphs=True)
    w0 = teacher.relation_nms.classifier.weight.detach().clone()
    torch.manual_seed(7)
    for it in range(5, 10):
        ld, total = trainer.train_step(it, images, targets, views)
        vals = {k: float(v) for k, v in ld.items()}
        assert "nms_loss" in vals and all(np.isfinite(v) for v in vals.values()), (it, vals)
    assert float((teacher.relation_nms.classifier.weight - w0).abs().max()) > 0
