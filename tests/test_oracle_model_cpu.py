"""CPU suite: the oracle's full-model restatement (oracle/model_cpu.py) against the fixture produced by running the
patched reference model itself (tests/golden/step_small.npz): same weights, inputs, seed and call order."""
import sys

import numpy as np
import torch

from conftest import ROOT, golden

sys.path.insert(0, ROOT)


def _setup(size):
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.config import cfg
    from mmt_psm_b200.data.synthetic import make_batch
    from mmt_psm_b200.modeling.detector import build_detection_model, reference_state_dict
    from mmt_psm_b200.utils.synthetic_init import synthetic_state_dict
    sys.path.insert(0, ROOT + "/scripts")
    ns = {}
    exec(open(ROOT + "/scripts/bench_cfg.py").read(), ns)
    c = ns["bench_cfg"](0)
    m = build_detection_model(c, is_student=True)
    shapes = {k: tuple(v.shape) for k, v in reference_state_dict(m).items()}
    sd = synthetic_state_dict(shapes, 0)
    images, targets, views = make_batch(1234, 2, size, 3, 6, 6)
    tg = [dict(bbox=t.bbox, labels=t.get_field("labels"),
               polygons=[[p.numpy() for p in inst.polygons] for inst in t.get_field("masks").polygons]) for t in targets]
    return sd, images, tg, views


def test_oracle_model_matches_reference_step():
    from oracle import model_cpu
    g = golden("step_small")
    sd, images, tg, views = _setup(int(g["size"]))
    oc = model_cpu.Cfg()
    oc.dropout = 0.0
    student = model_cpu.Net(sd, oc, True)
    teacher = model_cpu.Net(sd, oc, False)
    torch.manual_seed(7)
    sup = student.forward_train(images, tg)
    for k, v in sup.items():
        ref = float(g["sup_" + k])
        assert abs(float(v) - ref) <= 1e-5 * max(abs(ref), 1e-3), (k, float(v), ref)
    t = teacher.forward_teacher([views[0], views[1]])
    assert np.array_equal(np.array([len(s["bbox"]) for s in t["samples"]]), g["t_n_rois"])
    assert np.array_equal(torch.cat([s["labels"] for s in t["samples"]]).numpy(), g["t_labels"])
    np.testing.assert_allclose(torch.cat([s["bbox"] for s in t["samples"]]).numpy(), g["t_boxes"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(torch.stack(t["class_logit_t"]).numpy(), g["t_logits"], rtol=1e-4, atol=1e-5)
    assert np.array_equal(np.packbits(t["seg_mask"].numpy() > 0), g["t_seg_mask"])
    mt = student.forward_student(views[2], t)
    for k, v in mt.items():
        ref = float(g["mt_" + k])
        assert abs(float(v) - ref) <= 1e-4 * abs(ref), (k, float(v), ref)


def test_oracle_model_matches_reference_step_with_relation():
    """BASELINE config 4 (IR-Net duplicate removal + mask refinement on): the restatement against the patched reference's
    own step (tests/golden/step_relation.npz) — supervised losses incl. nms_loss, the teacher's relation-NMS detections in
    the refinement's (class, objectness) order, sampled ROIs, logits, foreground map, student-unlabeled losses."""
    from oracle import model_cpu, relation_ref
    from mmt_psm_b200.utils.synthetic_init import synthetic_state_dict
    g = golden("step_relation")
    sd, images, tg, views = _setup(int(g["size"]))
    sd.update(synthetic_state_dict(relation_ref.nms_weight_shapes(), 0))
    sd.update(synthetic_state_dict(relation_ref.mask_weight_shapes(), 0))
    oc = model_cpu.Cfg()
    oc.dropout = 0.0
    oc.relation = True
    student = model_cpu.Net(sd, oc, True)
    teacher = model_cpu.Net(sd, oc, False)
    torch.manual_seed(7)
    sup = student.forward_train(images, tg)
    assert "nms_loss" in sup
    for k, v in sup.items():
        ref = float(g["sup_" + k])
        assert abs(float(v) - ref) <= 1e-5 * max(abs(ref), 1e-3), (k, float(v), ref)
    torch.manual_seed(11)
    t = teacher.forward_teacher([views[0], views[1]])
    assert np.array_equal(np.array([len(d["bbox"]) for d in t["dets"]]), g["infer_n"])
    assert np.array_equal(torch.cat([d["labels"] for d in t["dets"]]).numpy(), g["infer_labels"])
    np.testing.assert_allclose(torch.cat([d["bbox"] for d in t["dets"]]).numpy(), g["infer_boxes"], rtol=1e-5, atol=1e-4)
    assert np.array_equal(np.array([len(s["bbox"]) for s in t["samples"]]), g["t_n_rois"])
    assert np.array_equal(torch.cat([s["labels"] for s in t["samples"]]).numpy(), g["t_labels"])
    np.testing.assert_allclose(torch.cat([s["bbox"] for s in t["samples"]]).numpy(), g["t_boxes"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(torch.stack(t["class_logit_t"]).numpy(), g["t_logits"], rtol=1e-4, atol=1e-5)
    assert np.array_equal(np.packbits(t["seg_mask"].numpy() > 0), g["t_seg_mask"])
    mt = student.forward_student(views[2], t)
    for k, v in mt.items():
        ref = float(g["mt_" + k])
        assert abs(float(v) - ref) <= 1e-4 * abs(ref), (k, float(v), ref)
