"""CPU suite: the oracle (oracle/ops_ref.py) against fixtures produced by the reference itself."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden

sys.path.insert(0, ROOT)
from oracle import ops_ref as R  # noqa: E402


def test_nms_oracle_matches_reference_cpu():
    g = golden("nms")
    for i in range(int(g["n_cases"])):
        b, s, thr = g["boxes%d" % i], g["scores%d" % i], float(g["thr%d" % i])
        # the two reference flavours (>= on CPU, > on CUDA) agree unless some IoU == thr exactly
        k_gt = R.nms_ref(b, s, thr, inclusive=False)
        k_ge = R.nms_ref(b, s, thr, inclusive=True)
        assert np.array_equal(k_gt, k_ge)
        assert np.array_equal(k_gt, g["keep%d" % i])


def test_nms_threshold_flavours_differ_at_exact_iou():
    # two boxes with IoU exactly 0.5 under the +1 convention: areas 4x2=8 and 4x2=8, inter 4x... build by hand
    b = np.array([[0, 0, 3, 1], [0, 0, 3, 1], [2, 0, 5, 1]], np.float32)  # box2 vs box0: inter 2x2=4, union 12 -> 1/3
    s = np.array([0.9, 0.8, 0.7], np.float32)
    assert list(R.nms_ref(b, s, 1.0, inclusive=False)) == [0, 1, 2]   # IoU == 1.0 is not > 1.0
    assert list(R.nms_ref(b, s, 1.0, inclusive=True)) == [0, 2]


def test_nms_oracle_vs_compiled_reference():
    ref = pytest.importorskip("oracle_ref_C", reason="oracle/_ref/_C not built", exc_type=ImportError)
    g = torch.Generator().manual_seed(5)
    for n in (0, 1, 50, 400):
        b = torch.rand(n, 4, generator=g) * 100
        b[:, 2:] += b[:, :2]
        s = torch.rand(n, generator=g)
        assert np.array_equal(ref.nms(b, s, 0.4).numpy(), R.nms_ref(b.numpy(), s.numpy(), 0.4, inclusive=True))


def test_roi_align_forward_oracle_bit_exact():
    g = golden("roi_align")
    for i in range(int(g["n_cases"])):
        scale, P, sr = g["cfg%d" % i]
        y = R.roi_align_forward_ref(g["x%d" % i], g["rois%d" % i], float(scale), int(P), int(P), int(sr))
        assert np.array_equal(y, g["y%d" % i]), "case %d max diff %g" % (i, np.abs(y - g["y%d" % i]).max())


def test_roi_align_backward_oracle_vs_torchvision():
    import torchvision.ops as tvo
    g = golden("roi_align")
    i = 0
    scale, P, sr = g["cfg%d" % i]
    x = torch.tensor(g["x%d" % i], requires_grad=True)
    rois = torch.tensor(g["rois%d" % i])
    y = tvo.roi_align(x, rois, int(P), float(scale), int(sr), aligned=False)
    go = torch.randn(y.shape, generator=torch.Generator().manual_seed(1))
    y.backward(go)
    mine = R.roi_align_backward_ref(go.numpy(), rois.numpy(), float(scale), int(P), int(P), *x.shape, int(sr))
    np.testing.assert_allclose(mine, x.grad.numpy(), rtol=1e-5, atol=1e-5)


def test_psm_oracle_matches_reference():
    g = golden("psm")
    for i in range(int(g["n_cases"])):
        s = torch.tensor(g["student%d" % i], requires_grad=True)
        loss, sel = R.psm_loss_ref(s, list(g["teacher%d" % i]), g["labels%d" % i], float(g["temperature"]), True, True,
                                   float(g["neg_weight"]))
        ref = float(g["loss%d" % i])
        if np.isnan(ref):
            assert torch.isnan(loss)
            continue
        assert abs(float(loss) - ref) <= 1e-6 * max(1, abs(ref))
        loss.backward()
        np.testing.assert_allclose(s.grad.numpy(), g["grad%d" % i], rtol=1e-5, atol=1e-8)


def test_mgd_oracle_matches_reference():
    g = golden("mgd")
    for i in range(int(g["n_cases"])):
        B, C, S, V = g["dims%d" % i]
        student = [torch.tensor(g["student%d_%d" % (i, l)], requires_grad=True) for l in range(5)]
        teachers = [[g["teacher%d_%d_%d" % (i, v, l)] for l in range(5)] for v in range(V)]
        loss, _ = R.mgd_loss_ref(teachers, student, g["mask%d" % i])
        assert abs(float(loss) - float(g["loss%d" % i])) <= 1e-6 * abs(float(g["loss%d" % i]))
        loss.backward()
        for l in range(5):
            np.testing.assert_allclose(student[l].grad.numpy(), g["grad%d_%d" % (i, l)], rtol=1e-5, atol=1e-9)


def test_ema_oracle_matches_reference():
    g = golden("ema_weights")
    tw, tb = g["t_w"], g["t_b"]
    for k in g["steps"]:
        a = R.ema_alpha(int(k), 0.99)
        tw = R.ema_ref(tw, g["s_w"], a).numpy()
        tb = R.ema_ref(tb, g["s_b"], a).numpy()
        assert np.array_equal(tw, g["t_w_after%d" % k]) and np.array_equal(tb, g["t_b_after%d" % k])


def test_level_mapper_oracle():
    g = golden("glue")
    assert np.array_equal(R.level_mapper_ref(g["lvl_boxes"]), g["lvl"])


def test_iou_and_matcher_oracle_match_reference():
    g = golden("glue")
    iou = R.boxlist_iou_ref(g["iou_gt"], g["iou_pr"])
    assert np.array_equal(iou.numpy(), g["iou"])
    assert np.array_equal(R.matcher_ref(iou, 0.7, 0.3, True), g["match_rpn"])
    assert np.array_equal(R.matcher_ref(iou, 0.5, 0.5, False), g["match_box"])


def test_decode_oracle_matches_reference_box_coder():
    """oracle decode == the reference's BoxCoder.decode on the golden pairs (both weight sets), bit for bit."""
    g = golden("glue")
    for name, w in (("rpn", (1., 1., 1., 1.)), ("box", (10., 10., 5., 5.))):
        codes = torch.from_numpy(g["coder_%s_enc" % name]) * 1.3 + 0.01
        out, valid = R.decode_clip_ref(codes, g["coder_prop"], w)
        assert np.array_equal(out, g["coder_%s_dec" % name]) and valid.all()
    # clip + size predicate: a box pushed outside a 100 x 80 image collapses onto the border and is filtered
    out, valid = R.decode_clip_ref(np.zeros((2, 4), np.float32), np.array([[10, 10, 49, 29], [150, 90, 190, 130]], np.float32),
                                   (1., 1., 1., 1.), image_wh=[[100, 80]], min_size=2)
    assert np.array_equal(out, np.array([[10, 10, 49, 29], [99, 79, 99, 79]], np.float32)) and valid.tolist() == [True, False]


def test_focal_oracle_vs_torchvision():
    import torchvision.ops as tvo
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(50, 4, generator=gen) * 3
    t = torch.randint(0, 5, (50,), generator=gen)
    mine = R.sigmoid_focal_loss_ref(x, t, 2.0, 0.25)
    onehot = torch.zeros(50, 4)
    for i, c in enumerate(t.tolist()):
        if c > 0:
            onehot[i, c - 1] = 1
    ref = tvo.sigmoid_focal_loss(x.double(), onehot.double(), alpha=0.25, gamma=2.0, reduction="none")
    np.testing.assert_allclose(mine.numpy(), ref.numpy(), rtol=1e-9, atol=1e-12)


def test_color_oracle_matches_pillow_and_reference_views():
    """oracle/color_ref.py: (1) its RGB <-> HSV restatement equals Pillow's on a quarter of all 2^24 colours (offsets
    cover every residue) and on every (h, s, v) the hue shift can produce; (2) brightness / contrast / hue equal
    torchvision's PIL ops (the reference's `tf.adjust_*` call path); (3) the whole view pipeline, with parameters
    drawn in the reference's order from seeded python / numpy streams, equals the reference's own transform classes
    (tests/golden/color_views.npz), bit for bit."""
    import importlib.util
    import random
    from PIL import Image
    import torchvision.transforms.functional as tf
    from conftest import ROOT
    from oracle import color_ref as C
    idx = np.arange(1, 1 << 24, 4, dtype=np.int64) + (np.arange((1 << 24) // 4) % 4) - 1
    cube = np.stack([(idx >> 16) & 255, (idx >> 8) & 255, idx & 255], -1).astype(np.uint8).reshape(2048, 2048, 3)
    assert np.array_equal(C.rgb_to_hsv(cube), np.array(Image.fromarray(cube, "RGB").convert("HSV")))
    assert np.array_equal(C.hsv_to_rgb(cube), np.array(Image.fromarray(cube, "HSV").convert("RGB")))
    g = golden("color_views")
    img = g["image"]
    pil = Image.fromarray(img, "RGB")
    for f in (0.85, 0.97, 1.0, 1.15):
        assert np.array_equal(C.adjust_brightness(img, f), np.array(tf.adjust_brightness(pil, f)))
        assert np.array_equal(C.adjust_contrast(img, f), np.array(tf.adjust_contrast(pil, f)))
    for h in (-0.05, -0.017, 0.0, 0.033, 0.05):
        assert np.array_equal(C.adjust_hue(img, h), np.array(tf.adjust_hue(pil, h)))
    spec = importlib.util.spec_from_file_location("uv_params", os.path.join(ROOT, "mmt-psm_b200", "data", "unlabeled_views.py"))
    src = open(spec.origin).read().replace("from .. import kernels as K", "K = None")
    ns = {}
    exec(compile(src, spec.origin, "exec"), ns)          # the parameter drawing only: no CUDA needed
    random.seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    for v in range(g["views"].shape[0]):
        p = ns["draw_view_params"](img.shape[0], img.shape[1])
        out = C.colour_view(img, p["brightness"], p["contrast"], p["hue"], p["rects"])
        assert np.array_equal(out, g["views"][v]), v


def _gold_t(g, k):
    a = g[k]
    return torch.from_numpy(a.astype(np.float32) if a.dtype == np.float16 else a)


def test_relation_nms_oracle_matches_reference():
    """IR-Net duplicate removal (reference relation_module.py): ranking, IoU-regression labels, loss + gradients and the
    eval-mode detections, bit-exact against the patched reference's module on seeded inputs (golden relation.npz)."""
    from oracle import relation_ref as RR
    from mmt_psm_b200.utils.synthetic_init import synthetic_state_dict
    g = golden("relation")
    cfg = RR.RelationCfg()
    sd = synthetic_state_dict(RR.nms_weight_shapes(), 0)
    for tag in "ab":
        a = {k: _gold_t(g, "nms_%s_%s" % (tag, k)) for k in ("x", "prob", "reg", "rois", "obj", "gt", "gt_labels")}
        x = a["x"].requires_grad_(True)
        for v in sd.values():
            v.requires_grad_(True)
            v.grad = None
        loss, ranked = RR.relation_nms_loss(sd, RR.NMS_PREFIX, cfg, x, a["prob"], a["reg"], a["rois"], a["obj"], (256, 256),
                                            a["gt"], a["gt_labels"])
        loss.backward()
        assert np.array_equal(ranked["idx"].numpy(), g["nms_%s_sorted_idx" % tag])
        assert np.array_equal(ranked["labels_reg"].numpy(), g["nms_%s_reg_label" % tag])
        assert (g["nms_%s_reg_label" % tag] > 0).sum() >= 2
        np.testing.assert_allclose(loss.item(), g["nms_%s_loss" % tag], rtol=1e-6)
        np.testing.assert_allclose(x.grad.numpy(), g["nms_%s_dx" % tag], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(sd[RR.NMS_PREFIX + "relation_module.WG.weight"].grad.numpy(), g["nms_%s_dWG" % tag],
                                   rtol=1e-5, atol=1e-8)
        det = RR.relation_nms_detect({k: v.detach() for k, v in sd.items()}, RR.NMS_PREFIX, cfg, x.detach(), a["prob"],
                                     a["reg"], a["rois"], a["obj"], (256, 256))
        assert len(det["labels"]) > 10
        for f in ("bbox", "scores", "labels", "objectness", "all_scores"):
            assert np.array_equal(det[f].numpy(), g["nms_%s_det_%s" % (tag, f)]), f


def test_relation_mask_oracle_matches_reference():
    """IR-Net mask refinement (reference mask_relation_module.py, CIAM): refined logits, row order and gradients."""
    from oracle import relation_ref as RR
    from mmt_psm_b200.utils.synthetic_init import synthetic_state_dict
    g = golden("relation")
    sd = synthetic_state_dict(RR.mask_weight_shapes(), 0)
    for tag in "ab":
        f = _gold_t(g, "msk_%s_f" % tag).requires_grad_(True)
        m = _gold_t(g, "msk_%s_m" % tag).requires_grad_(True)
        for v in sd.values():
            v.requires_grad_(True)
            v.grad = None
        out, order = RR.mask_relation(sd, RR.MASK_PREFIX, f, m, _gold_t(g, "msk_%s_labels" % tag), _gold_t(g, "msk_%s_obj" % tag))
        (out * _gold_t(g, "msk_%s_w" % tag)).sum().backward()
        np.testing.assert_allclose(out.detach().numpy(), g["msk_%s_logits2" % tag], rtol=1e-5, atol=1e-6)
        assert np.array_equal(g["msk_%s_boxes" % tag][order.numpy()], g["msk_%s_sorted_boxes" % tag])
        assert np.array_equal(g["msk_%s_labels" % tag][order.numpy()], g["msk_%s_sorted_labels" % tag])
        np.testing.assert_allclose(m.grad.numpy(), g["msk_%s_dm" % tag], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(f.grad.sum((2, 3)).numpy(), g["msk_%s_df_sum" % tag], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(sd[RR.MASK_PREFIX + "relation_module.gamma"].grad.numpy(), g["msk_%s_dgamma" % tag], rtol=1e-4)


# ------------------------------------------------------------------------------------------- size-independent properties
def test_oracle_properties_hypothesis():
    """Domain properties the path offers, on random cases: NMS is idempotent and its survivors are pairwise below the
    threshold; BoxCoder decode(encode(gt, anchor), anchor) returns gt (legacy +1 extents); the IR-Net label preparation
    gives every ground truth at most one winner, only boxes above the threshold win, and a winner's IoU-regression target is
    its IoU with a ground truth of that class."""
    from hypothesis import given, settings, strategies as st
    from oracle import relation_ref as RR

    def boxes_from(seed, n, size=300.0):
        g = torch.Generator().manual_seed(seed)
        c = torch.rand(n, 2, generator=g) * size
        wh = torch.rand(n, 2, generator=g) * 60 + 4
        return torch.cat([c - wh / 2, c + wh / 2], 1)

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(st.integers(0, 10 ** 6), st.integers(1, 120), st.sampled_from([0.3, 0.5, 0.7]))
    def nms_props(seed, n, thr):
        b = boxes_from(seed, n).numpy()
        s = torch.rand(n, generator=torch.Generator().manual_seed(seed + 1)).numpy()
        keep = np.asarray(R.nms_ref(b, s, thr))
        assert np.all(np.diff(keep) > 0)                                   # ascending original indices
        again = np.asarray(R.nms_ref(b[keep], s[keep], thr))
        assert np.array_equal(again, np.arange(len(keep)))                 # idempotent
        iou = R.boxlist_iou_ref(b[keep], b[keep]).numpy().copy()
        np.fill_diagonal(iou, 0)
        assert (iou <= thr + 1e-6).all()

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(st.integers(0, 10 ** 6), st.integers(1, 60))
    def coder_roundtrip(seed, n):
        from oracle import model_cpu as M
        gt, an = boxes_from(seed, n), boxes_from(seed + 7, n)
        w = (10., 10., 5., 5.)
        back = M.decode(M.encode(gt, an, w), an, w)
        assert torch.allclose(back, gt, rtol=1e-4, atol=2e-3)

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(st.integers(0, 10 ** 6), st.integers(1, 90), st.integers(0, 30))
    def label_props(seed, n, m):
        g = torch.Generator().manual_seed(seed)
        gt = boxes_from(seed, m)
        gl = torch.randint(1, 3, (m,), generator=g)
        if m:
            boxes = gt[torch.randint(0, m, (n, 2), generator=g)] + torch.randn(n, 2, 4, generator=g) * 4
        else:
            boxes = boxes_from(seed + 3, 2 * n).view(n, 2, 4)
        boxes[..., 2:] = torch.max(boxes[..., 2:], boxes[..., :2] + 1)
        scores = torch.sort(torch.rand(n, 2, generator=g), 0, descending=True)[0]
        lab = RR.reg_labels(boxes, scores, gt, gl, (0.1,))
        hot = RR.reg_labels(boxes, scores, gt, gl, (0.1,), binary=True)
        assert lab.shape == (n, 2, 1) and (lab >= 0).all() and (lab <= 1 + 1e-6).all()
        assert np.array_equal(lab > 0, (hot > 0) & (lab > 0))
        for c in range(2):
            k = int((gl == c + 1).sum())
            assert (hot[:, c, 0] > 0).sum() <= max(k, 0)                    # at most one winner per ground truth
            if k:
                iou = RR._iou_plus1(boxes[:, c], gt[gl == c + 1]).numpy()
                win = lab[:, c, 0] > 0
                assert (iou.max(1)[win] > 0.1).all()
                assert np.allclose(lab[win, c, 0], iou.max(1)[win], atol=1e-6)

    nms_props()
    coder_roundtrip()
    label_props()


def _mask_target_fixture():
    """tests/golden/mask_targets.npz -> (instances: per instance a list of flat fp32 polygons, which, boxes, masks)."""
    g = golden("mask_targets")
    xy, poly_len, inst_npoly = g["xy"], g["poly_len"], g["inst_npoly"]
    polys, pos = [], 0
    for n in poly_len:
        polys.append(xy[pos:pos + 2 * int(n)].astype(np.float32))
        pos += 2 * int(n)
    inst, pi = [], 0
    for n in inst_npoly:
        inst.append(polys[pi:pi + int(n)])
        pi += int(n)
    return inst, g["which"], g["boxes"], g["masks"]


def test_polygon_oracle_bit_exact_reference_mask_targets():
    """oracle.polygon_ref (rleFrPoly / decode / merge restated) reproduces all 96 28x28 mask targets the reference's
    project_masks_on_boxes + its own pycocotools build produced (mask_head/loss.py:37-75, structures/segmentation_mask.py:
    94-137, pycoco/maskApi.c:166-206) — bit for bit, including the sub-pixel box (row 3: w, h < 1 clamp to 1), the
    integer box (row 5) and the two-polygon instances."""
    from oracle import polygon_ref
    inst, which, boxes, masks = _mask_target_fixture()
    assert masks.shape == (96, 28, 28) and any(len(p) == 2 for p in inst)
    got = polygon_ref.project_masks_on_boxes_ref([inst[i] for i in which], boxes, 28)
    assert got.dtype == np.float32 and set(np.unique(got)) <= {0.0, 1.0}
    bad = [i for i in range(len(which)) if not np.array_equal(got[i].astype(np.uint8), masks[i])]
    assert not bad, bad
    assert masks.any(axis=(1, 2)).mean() > 0.9 and 0.2 < masks.mean() < 0.9     # the fixture is not trivially empty / full


def test_balanced_sample_restatement_properties():
    """oracle.ops_ref.balanced_sample_ref (the restated draw the shipped device sampler is pinned to on the GPU): exact class
    sizes of the reference's sampler (balanced_positive_negative_sampler.py:41-54), only eligible rows, per-image keys,
    seed-dependent, and uniform over seeds."""
    rng = np.random.default_rng(3)
    lab = np.where(rng.random(4000) < 0.12, 1.0, 0.0).astype(np.float32)
    lab[rng.integers(0, 4000, 300)] = -1.0
    small = np.array([1.0] * 5 + [0.0] * 20 + [-1.0] * 12, dtype=np.float32)
    labels = np.concatenate([lab, small, lab])
    offs = [0, 4000, 4037, 8037]
    p, n, unique = R.balanced_sample_ref(labels, offs, 512, 128, 42)
    assert unique and not (p & (labels < 1)).any() and not (n & (labels != 0)).any()
    for a, b in zip(offs[:-1], offs[1:]):
        want_p = min(int((labels[a:b] >= 1).sum()), 128)
        assert int(p[a:b].sum()) == want_p and int(n[a:b].sum()) == min(int((labels[a:b] == 0).sum()), 512 - want_p)
    assert not np.array_equal(p[:4000], p[4037:])                      # same labels, different image index: different draw
    p2, _, _ = R.balanced_sample_ref(labels, offs, 512, 128, 43)
    assert not np.array_equal(p, p2)
    hits = np.zeros(4000)
    for s in range(300):
        hits += R.balanced_sample_ref(lab, [0, 4000], 512, 128, 1000 + s)[0]
    f = hits[lab >= 1] / 300.0
    exp = 128.0 / float((lab >= 1).sum())
    assert abs(f.mean() - exp) < 1e-9 and np.abs(f - exp).max() < 4.5 * (exp * (1 - exp) / 300.0) ** 0.5
