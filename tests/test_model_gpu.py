"""GPU suite, step level: the CUDA path (this package's GeneralizedRCNN on libmmtpsm_b200) against
  (1) tests/golden/step_small.npz — losses / teacher outputs of the PATCHED REFERENCE MODEL itself on the same
      weights, inputs and CPU random stream (tests/golden/make_golden.py gen_step), and
  (2) the oracle's full training step (oracle/model_cpu.py Trainer.step: backward + SGD + EMA), run on the host at
      the same small size, comparing the UPDATED student / teacher weights.
Tolerances are written next to each assertion.  Conv math "fp32" (FFMA kernels) isolates the parity of the whole
path from TF32 operand rounding; "tf32" (the shipped default, = torch/cuDNN's default allow_tf32=True the reference
runs with) is checked at the looser tolerance that 10-bit-mantissa operands through ~60 conv layers allow."""
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden

sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def _build(cuda, dropout=0.0):
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.config.presets import mt_psm_cfg
    from mmt_psm_b200.data.synthetic import make_batch
    from mmt_psm_b200.modeling.balanced_positive_negative_sampler import BalancedPositiveNegativeSampler
    from mmt_psm_b200.modeling.detector import build_detection_model
    from mmt_psm_b200.utils.synthetic_init import init_model
    BalancedPositiveNegativeSampler.rng_device = "cpu"   # consume the reference's CPU random stream
    cfg = mt_psm_cfg(start_mt=0, extra=["MODEL.ROI_BOX_HEAD.DO", dropout])
    student = build_detection_model(cfg, is_student=True).to(cuda)
    teacher = build_detection_model(cfg, is_teacher=True).to(cuda)
    sd = init_model(student, 0)
    teacher.load_state_dict(student.state_dict())
    student.train()
    teacher.eval()
    g = golden("step_small")
    images, targets, views = make_batch(1234, 2, int(g["size"]), 3, 6, 6)
    return cfg, student, teacher, sd, images, targets, views, g


@pytest.fixture
def restore():
    from mmt_psm_b200 import kernels
    from mmt_psm_b200.modeling.balanced_positive_negative_sampler import BalancedPositiveNegativeSampler
    yield
    kernels.set_conv_math("tf32")
    BalancedPositiveNegativeSampler.rng_device = None


@pytest.mark.parametrize("math,tol_sup,tol_mt,tol_logit", [("fp32", 2e-4, 1e-3, 2e-4), ("tf32", 3e-2, 6e-2, 5e-2)])
def test_step_losses_match_reference_golden(cuda, restore, math, tol_sup, tol_mt, tol_logit):
    from mmt_psm_b200 import kernels, lib
    kernels.set_conv_math(math)
    cfg, student, teacher, sd, images, targets, views, g = _build(cuda)
    lib.reset_launch_count()
    torch.manual_seed(7)
    sup = student(images.to(cuda), [t.to(cuda) for t in targets])
    for k in ("loss_classifier", "loss_box_reg", "loss_seg", "loss_objectness", "loss_rpn_box_reg"):
        ref, got = float(g["sup_" + k]), float(sup[k])
        assert abs(got - ref) <= tol_sup * max(abs(ref), 1e-2), (math, k, got, ref)
    with torch.no_grad():
        tr = teacher.forward_teacher([views[0].to(cuda), views[1].to(cuda)])
    n_rois = np.array([len(b) for b in tr["result_t"]])
    assert np.array_equal(n_rois, g["t_n_rois"])
    labels = torch.cat([b.get_field("labels") for b in tr["result_t"]]).cpu().numpy()
    boxes = torch.cat([b.bbox for b in tr["result_t"]]).cpu().numpy()
    seg = np.packbits(torch.stack(list(tr["seg_mask"])).cpu().numpy() > 0)
    if math == "fp32":
        # the discrete decisions of the teacher pass (proposal ranking / NMS / sampling / pseudo-labels) are identical
        assert np.array_equal(labels, g["t_labels"])
        np.testing.assert_allclose(boxes, g["t_boxes"], rtol=1e-4, atol=2e-3)
        logits = torch.stack(list(tr["class_logit_t"])).cpu().numpy()
        np.testing.assert_allclose(logits, g["t_logits"], rtol=tol_logit, atol=tol_logit)
        diff = np.unpackbits(seg) != np.unpackbits(g["t_seg_mask"])
        assert diff.mean() < 1e-4, diff.mean()          # foreground map: at most a few boundary pixels (prob == 0.5)
    else:
        assert (labels == g["t_labels"]).mean() > 0.9   # TF32 perturbs near-tie rankings; most ROIs still coincide
    mt = student.forward_student([views[2].to(cuda)], tr)
    for k in ("mt_fg_loss", "mt_classifier"):
        ref, got = float(g["mt_" + k]), float(mt[k])
        assert abs(got - ref) <= tol_mt * abs(ref), (math, k, got, ref)
    assert lib.launch_count() > 300       # the CUDA extension did the work


def test_full_training_step_updates_match_oracle(cuda, restore):
    """backward + fused SGD + EMA: the updated student and teacher weights after one mean-teacher iteration equal
    the oracle trainer's (fp32 conv math; tolerance 2e-3 of each tensor's max-abs update, 1e-6 absolute floor)."""
    from mmt_psm_b200 import kernels
    from mmt_psm_b200.engine import MTtrainer
    from mmt_psm_b200.modeling.detector import reference_state_dict
    from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer
    from oracle import model_cpu
    kernels.set_conv_math("fp32")
    cfg, student, teacher, sd, images, targets, views, g = _build(cuda)
    opt = make_optimizer(cfg, student, teacher)
    trainer = MTtrainer(student, teacher, None, opt, make_lr_scheduler(cfg, opt), None, None, 10 ** 9, cfg)
    before = {k: v.clone() for k, v in reference_state_dict(student).items()}
    torch.manual_seed(7)
    ld, total = trainer.train_step(5, images, targets, views)

    oc = model_cpu.Cfg()
    oc.dropout, oc.start_mt = 0.0, 0
    ot = model_cpu.Trainer({k: v.cpu() for k, v in before.items()}, oc)
    tg = [dict(bbox=t.bbox, labels=t.get_field("labels"),
               polygons=[[p.numpy() for p in inst.polygons] for inst in t.get_field("masks").polygons]) for t in targets]
    torch.manual_seed(7)
    old, ototal = ot.step(5, images, tg, views)
    for k, v in old.items():
        assert abs(float(ld[k]) - float(v)) <= 1e-3 * max(abs(float(v)), 1e-2), (k, float(ld[k]), float(v))
    after_s = reference_state_dict(student)
    after_t = reference_state_dict(teacher)
    worst = 0.0
    for name in ot.student.trainable:
        ref = ot.student.sd[name].detach()
        got = after_s[name].cpu()
        upd = (ref - before[name].cpu()).abs().max().item()
        err = (got - ref).abs().max().item()
        assert err <= 2e-3 * upd + 1e-6, (name, err, upd)
        worst = max(worst, err / (upd + 1e-12))
        tref = ot.teacher.sd[name].detach()
        terr = (after_t[name].cpu() - tref).abs().max().item()
        assert terr <= 2e-3 * upd + 1e-6, ("teacher", name, terr, upd)
    print("worst relative update error %.2e" % worst)
