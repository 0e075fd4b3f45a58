"""GPU parity tests proper: every CUDA op, called through the C ABI (ctypes), against
(a) fixtures produced by the reference itself (tests/golden) and (b) the CPU oracle on seeded inputs."""
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden

sys.path.insert(0, ROOT)
from oracle import ops_ref as R  # noqa: E402

pytestmark = pytest.mark.gpu


def K():
    from mmt_psm_b200 import kernels
    return kernels


# ------------------------------------------------------------------------------------------ NMS
def test_nms_golden_bit_exact(cuda):
    g = golden("nms")
    for i in range(int(g["n_cases"])):
        b = torch.tensor(g["boxes%d" % i], device=cuda)
        s = torch.tensor(g["scores%d" % i], device=cuda)
        keep = K().nms(b, s, float(g["thr%d" % i]))
        assert keep.dtype == torch.int64 and keep.is_cuda
        assert np.array_equal(keep.cpu().numpy(), g["keep%d" % i])


def test_nms_full_size_ties_and_empty(cuda):
    gen = torch.Generator().manual_seed(21)
    # RPN-sized problem (2000 boxes) with saturated, tied scores (sigmoid == 1.0f) — SURVEY "NMS ties"
    n = 2000
    c = torch.rand(n, 2, generator=gen) * 780
    wh = torch.rand(n, 2, generator=gen) * 120 + 4
    b = torch.cat([c, c + wh], 1)
    s = torch.sigmoid(torch.randn(n, generator=gen) * 12)
    assert (s == 1).sum() > 10
    keep = K().nms(b.to(cuda), s.to(cuda), 0.7).cpu().numpy()
    assert np.array_equal(keep, R.nms_ref(b.numpy(), s.numpy(), 0.7))
    assert np.all(np.diff(keep) > 0)
    # idempotence: NMS of the survivors keeps everything
    k2 = K().nms(b[keep].to(cuda), s[keep].to(cuda), 0.7).cpu().numpy()
    assert np.array_equal(k2, np.arange(len(keep)))
    # exact-threshold case distinguishes the CUDA (>) from the CPU (>=) flavour
    bb = torch.tensor([[0, 0, 3, 1], [0, 0, 3, 1], [2, 0, 5, 1.]], device=cuda)
    ss = torch.tensor([0.9, 0.8, 0.7], device=cuda)
    assert K().nms(bb, ss, 1.0).tolist() == [0, 1, 2]
    # empty input: empty int64 CPU tensor (reference csrc/nms.h:17-18)
    e = K().nms(torch.zeros(0, 4, device=cuda), torch.zeros(0, device=cuda), 0.5)
    assert e.numel() == 0 and e.dtype == torch.int64 and e.device.type == "cpu"
    # single box
    assert K().nms(bb[:1], ss[:1], 0.5).tolist() == [0]


def test_nms_segmented_matches_per_segment(cuda):
    gen = torch.Generator().manual_seed(22)
    sizes = [2000, 0, 1875, 507, 1, 64, 65, 2000, 999, 130]
    boxes, scores = [], []
    for n in sizes:
        c = torch.rand(n, 2, generator=gen) * 700
        wh = torch.rand(n, 2, generator=gen) * 90 + 2
        boxes.append(torch.cat([c, c + wh], 1))
        scores.append(torch.rand(n, generator=gen))
    offs = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    keep, counts = K().nms_segmented(torch.cat(boxes).to(cuda), torch.cat(scores).to(cuda), offs, 0.7)
    keep, counts = keep.cpu().numpy(), counts.cpu().numpy()
    for i, n in enumerate(sizes):
        ref = R.nms_ref(boxes[i].numpy(), scores[i].numpy(), 0.7)
        assert counts[i] == len(ref)
        assert np.array_equal(keep[offs[i]: offs[i] + counts[i]], ref)


# ------------------------------------------------------------------------------------- ROIAlign
def test_roi_align_forward_golden_bit_exact(cuda):
    g = golden("roi_align")
    for i in range(int(g["n_cases"])):
        scale, P, sr = g["cfg%d" % i]
        x = torch.tensor(g["x%d" % i], device=cuda)
        rois = torch.tensor(g["rois%d" % i], device=cuda)
        y = K().roi_align_forward(x, rois, float(scale), int(P), int(P), int(sr))
        assert np.array_equal(y.cpu().numpy(), g["y%d" % i]), "NCHW case %d" % i
        y2 = K().roi_align_forward(x.contiguous(memory_format=torch.channels_last), rois, float(scale), int(P), int(P), int(sr))
        assert np.array_equal(y2.contiguous().cpu().numpy(), g["y%d" % i]), "NHWC case %d" % i


def test_roi_align_backward_vs_oracle(cuda):
    g = golden("roi_align")
    for i in (0, 2, 3):
        scale, P, sr = g["cfg%d" % i]
        x = g["x%d" % i]
        rois = g["rois%d" % i]
        go = torch.randn(rois.shape[0], x.shape[1], int(P), int(P), generator=torch.Generator().manual_seed(i))
        ref = R.roi_align_backward_ref(go.numpy(), rois, float(scale), int(P), int(P), *x.shape, int(sr))
        for fmt in (torch.contiguous_format, torch.channels_last):
            gi = K().roi_align_backward(go.to(cuda).contiguous(memory_format=fmt), torch.tensor(rois, device=cuda),
                                        float(scale), int(P), int(P), *x.shape, int(sr))
            np.testing.assert_allclose(gi.contiguous().cpu().numpy(), ref, rtol=1e-5, atol=2e-5)


def test_roi_align_fpn_fused_equals_per_level(cuda):
    gen = torch.Generator().manual_seed(23)
    B, C = 2, 256
    sizes = [(200, 200), (100, 100), (50, 50), (25, 25)]
    scales = [1 / 4., 1 / 8., 1 / 16., 1 / 32.]
    feats = [torch.randn(B, h, w, C, generator=gen).to(cuda) for h, w in sizes]
    n = 1024
    c = torch.rand(n, 2, generator=gen) * 800
    wh = torch.exp(torch.rand(n, 2, generator=gen) * 5.5 + 1)
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 799)
    ids = torch.randint(0, B, (n, 1), generator=gen).float()
    rois = torch.cat([ids, boxes], 1)
    levels = torch.tensor(R.level_mapper_ref(boxes.numpy()), dtype=torch.int32)
    assert len(set(levels.tolist())) == 4
    for P in (7, 14):
        out = K().roi_align_fpn_forward(feats, scales, rois.to(cuda), levels.to(cuda), P, 2)
        go = torch.randn(out.shape, generator=gen).to(cuda)
        grads = [torch.zeros_like(f) for f in feats]
        K().roi_align_fpn_backward(go, scales, rois.to(cuda), levels.to(cuda), P, 2, grads)
        for l in range(4):
            idx = torch.nonzero(levels == l).squeeze(1)
            x_nchw = feats[l].permute(0, 3, 1, 2)  # channels_last view
            ref = K().roi_align_forward(x_nchw, rois[idx].to(cuda), scales[l], P, P, 2)
            assert torch.equal(out[idx.to(cuda)], ref.permute(0, 2, 3, 1))
            gref = K().roi_align_backward(go[idx.to(cuda)].permute(0, 3, 1, 2), rois[idx].to(cuda), scales[l], P, P, B, C,
                                          sizes[l][0], sizes[l][1], 2)
            torch.testing.assert_close(grads[l], gref.permute(0, 2, 3, 1), rtol=1e-4, atol=1e-4)
    # small level checked against the CPU oracle too
    l = 3
    idx = torch.nonzero(levels == l).squeeze(1)[:40]
    ref = R.roi_align_forward_ref(feats[l].permute(0, 3, 1, 2).cpu().numpy(), rois[idx].numpy(), scales[l], 7, 7, 2)
    out = K().roi_align_fpn_forward(feats, scales, rois[idx].to(cuda), levels[idx].to(cuda), 7, 2)
    assert np.array_equal(out.permute(0, 3, 1, 2).cpu().numpy(), ref)


def test_roi_pool_vs_torchvision(cuda):
    import torchvision.ops as tvo
    gen = torch.Generator().manual_seed(24)
    x = torch.randn(2, 6, 30, 40, generator=gen)
    c = torch.rand(20, 2, generator=gen) * torch.tensor([300., 220.])
    wh = torch.rand(20, 2, generator=gen) * 100 + 4
    rois = torch.cat([torch.randint(0, 2, (20, 1), generator=gen).float(), c, c + wh], 1)
    out, arg = K().roi_pool_forward(x.to(cuda), rois.to(cuda), 0.125, 7, 7)
    ref = tvo.roi_pool(x, rois, 7, 0.125)
    torch.testing.assert_close(out.cpu(), ref)
    go = torch.randn(out.shape, generator=gen)
    xr = x.clone().requires_grad_(True)
    tvo.roi_pool(xr, rois, 7, 0.125).backward(go)
    gi = K().roi_pool_backward(go.to(cuda), x.to(cuda), rois.to(cuda), arg, 0.125, 7, 7, 2, 6, 30, 40)
    torch.testing.assert_close(gi.cpu(), xr.grad, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------- losses
def test_sigmoid_focal_loss_vs_oracle(cuda):
    gen = torch.Generator().manual_seed(25)
    x = torch.randn(4000, 3, generator=gen) * 4
    t = torch.randint(-1, 4, (4000,), generator=gen)
    ref = R.sigmoid_focal_loss_ref(x, t, 2.0, 0.25)
    out = K().sigmoid_focal_loss_forward(x.to(cuda), t.to(cuda), 3, 2.0, 0.25)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-7)  # fp32 vs fp64: 1e-4 rel
    xd = x.double().requires_grad_(True)
    w = torch.randn(4000, 3, generator=gen)
    (R.sigmoid_focal_loss_ref(xd, t, 2.0, 0.25) * w.double()).sum().backward()
    gx = K().sigmoid_focal_loss_backward(x.to(cuda), t.to(cuda), w.to(cuda), 3, 2.0, 0.25)
    np.testing.assert_allclose(gx.cpu().numpy(), xd.grad.numpy(), rtol=1e-4, atol=1e-6)


def test_psm_golden_and_oracle(cuda):
    g = golden("psm")
    for i in range(int(g["n_cases"])):
        s, t, lab = g["student%d" % i], g["teacher%d" % i], g["labels%d" % i]
        loss, grad, sel = K().psm_loss(torch.tensor(s, device=cuda), torch.tensor(t, device=cuda),
                                       torch.tensor(lab, device=cuda), float(g["temperature"]), True, True,
                                       float(g["neg_weight"]))
        ref = float(g["loss%d" % i])
        _, sel_ref = R.psm_loss_ref(s, list(t), lab, float(g["temperature"]), True, True, float(g["neg_weight"]))
        assert np.array_equal(sel.cpu().numpy(), sel_ref.numpy())  # selection indices exact
        if np.isnan(ref):
            assert torch.isnan(loss).all()
            continue
        assert abs(float(loss) - ref) <= 1e-4 * abs(ref)  # fp32 losses within 1e-4 rel (north_star)
        np.testing.assert_allclose(grad.cpu().numpy(), g["grad%d" % i], rtol=1e-4, atol=1e-8)


def test_mgd_golden(cuda):
    g = golden("mgd")
    for i in range(int(g["n_cases"])):
        B, C, S, V = [int(v) for v in g["dims%d" % i]]
        mask = torch.tensor(g["mask%d" % i], device=cuda).float()
        student = [torch.tensor(g["student%d_%d" % (i, l)], device=cuda).permute(0, 2, 3, 1).contiguous() for l in range(5)]
        sizes = [(s.shape[1], s.shape[2]) for s in student]
        masks, sums = K().mgd_masks(mask, sizes)
        _, ref_masks = R.mgd_loss_ref([[g["teacher%d_%d_%d" % (i, v, l)] for l in range(5)] for v in range(V)],
                                      [g["student%d_%d" % (i, l)] for l in range(5)], g["mask%d" % i])
        partials = torch.zeros(5, V, device=cuda)
        flip = sum(1 << v for v in range(V) if v % 2 == 1)
        grads = []
        for l in range(5):
            assert np.array_equal(masks[l].cpu().numpy(), ref_masks[l][:, 0].numpy().astype(np.uint8))
            teachers = [torch.tensor(g["teacher%d_%d_%d" % (i, v, l)], device=cuda).permute(0, 2, 3, 1).contiguous() for v in range(V)]
            grads.append(K().mgd_level(student[l], teachers, flip, masks[l], sums[l:l + 1], 1.0 / (5 * V), partials[l], True))
        loss = K().mgd_finalize(partials, sums, 5, V, C)
        ref = float(g["loss%d" % i])
        assert abs(float(loss) - ref) <= 1e-4 * abs(ref)
        for l in range(5):
            np.testing.assert_allclose(grads[l].permute(0, 3, 1, 2).cpu().numpy(), g["grad%d_%d" % (i, l)], rtol=1e-4, atol=1e-8)


def test_ema_golden_and_sgd(cuda):
    g = golden("ema_weights")
    s = torch.tensor(np.concatenate([g["s_w"].ravel(), g["s_b"].ravel()]), device=cuda)
    t = torch.tensor(np.concatenate([g["t_w"].ravel(), g["t_b"].ravel()]), device=cuda)
    for k in g["steps"]:
        K().ema_(t, s, R.ema_alpha(int(k), 0.99))
        ref = np.concatenate([g["t_w_after%d" % k].ravel(), g["t_b_after%d" % k].ravel()])
        got = t.cpu().numpy()
        ulp = np.abs(got.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
        assert ulp.max() <= 2, ulp.max()   # reference add_ may or may not be FMA-contracted
    # fused SGD(+momentum, wd) + EMA over a 3-tensor arena with padding between tensors
    gen = torch.Generator().manual_seed(26)
    sizes = [1000, 37, 4096]
    begins, total = [], 0
    for n in sizes:
        begins.append(total)
        total += (n + 3) // 4 * 4
    p = torch.randn(total, generator=gen)
    tt = torch.randn(total, generator=gen)
    gr = torch.randn(total, generator=gen)
    buf = torch.randn(total, generator=gen)
    hyper = [(0.005, 1e-4, 1), (0.01, 0.0, 1), (0.005, 1e-4, 0)]
    segs = K().make_segments([(b, b + n, lr, wd, tr) for (b, n, (lr, wd, tr)) in zip(begins, sizes, hyper)], cuda)
    for first in (True, False):
        pd, td, bd = p.to(cuda), tt.to(cuda), buf.to(cuda)
        K().sgd_ema_(pd, td, gr.to(cuda), bd, segs, 3, 0.9, 0.99, first, 1.0)
        for b, n, (lr, wd, tr) in zip(begins, sizes, hyper):
            sl = slice(b, b + n)
            if tr:
                p_ref, b_ref = R.sgd_ref(p[sl], gr[sl], buf[sl], lr, 0.9, wd, first)
            else:
                p_ref, b_ref = p[sl], buf[sl]
            t_ref = R.ema_ref(tt[sl], p_ref, 0.99)
            torch.testing.assert_close(pd[sl].cpu(), p_ref, rtol=1e-6, atol=1e-7)
            torch.testing.assert_close(td[sl].cpu(), t_ref, rtol=1e-6, atol=1e-7)
            if tr:
                torch.testing.assert_close(bd[sl].cpu(), b_ref, rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------ elementwise
def test_elementwise_vs_torch(cuda):
    import torch.nn.functional as Fn
    gen = torch.Generator().manual_seed(27)
    x = torch.randn(2, 37, 50, 64, generator=gen).to(cuda)  # NHWC
    xn = x.permute(0, 3, 1, 2)
    assert torch.equal(K().maxpool3x3s2(x), Fn.max_pool2d(xn, 3, 2, 1).permute(0, 2, 3, 1))
    assert torch.equal(K().subsample2(x), Fn.max_pool2d(xn, 1, 2, 0).permute(0, 2, 3, 1))
    top = torch.randn(2, 10, 12, 16, generator=gen).to(cuda)
    lat = torch.randn(2, 20, 24, 16, generator=gen).to(cuda)
    ref = lat.permute(0, 3, 1, 2) + Fn.interpolate(top.permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    assert torch.equal(K().upsample2_add(lat, top), ref.permute(0, 2, 3, 1))
    go = torch.randn(2, 20, 24, 16, generator=gen).to(cuda)
    gt = torch.zeros_like(top)
    K().upsample2_backward_(go, gt)
    torch.testing.assert_close(gt, go.view(2, 10, 2, 12, 2, 16).sum((2, 4)))
    gx = torch.zeros_like(x)
    gy = torch.randn(2, 19, 25, 64, generator=gen).to(cuda)
    K().subsample2_backward_(gy, gx)
    assert torch.equal(gx[:, ::2, ::2], gy) and float(gx.abs().sum()) == float(gy.abs().sum())
    y = torch.randn(2, 9, 11, 32, generator=gen).to(cuda)
    gyy = torch.randn(2, 9, 11, 32, generator=gen).to(cuda)
    sc = torch.rand(32, generator=gen).to(cuda) + 0.5
    torch.testing.assert_close(K().bn_relu_backward(gyy, y, sc, True), gyy * (y > 0) * sc)
    a = torch.randn(3, 5, 17, 33, generator=gen).to(cuda)
    assert torch.equal(K().nchw_to_nhwc(a), a.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(K().nhwc_to_nchw(K().nchw_to_nhwc(a)), a)


# ----------------------------------------------------------------------------- colour views
def test_color_views_bit_exact(cuda):
    """mmt_color_views_u8 == the reference's host pipeline: (1) the golden views produced by the reference's own
    transform classes, with the parameters re-drawn from the same seeds; (2) the oracle on a batch of two larger
    random crops x 3 views with erase rectangles, including saturated / grey / black pixels.  Bit-exact."""
    import random
    from mmt_psm_b200.data.unlabeled_views import colour_views, draw_view_params
    from oracle import color_ref as C
    g = golden("color_views")
    img = g["image"]
    random.seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    params = [[draw_view_params(img.shape[0], img.shape[1])] for _ in range(g["views"].shape[0])]
    views = colour_views(torch.from_numpy(img)[None].to(cuda), params)
    for v, out in enumerate(views):
        assert np.array_equal(out[0].cpu().numpy(), g["views"][v]), v
    rng = np.random.default_rng(3)
    batch = rng.integers(0, 256, (2, 150, 210, 3), dtype=np.uint8)
    batch[0, :20] = 0
    batch[0, 20:40] = 255
    batch[1, :30, :, :] = batch[1, :30, :, :1]                # grey rows (s == 0)
    batch[1, 60:90, :, 2] = 0
    random.seed(11)
    np.random.seed(11)
    params = [[draw_view_params(150, 210) for _ in range(2)] for _ in range(3)]
    params[1][0]["brightness"], params[2][1]["hue"] = 1.15, -0.05
    views = colour_views(torch.from_numpy(batch).to(cuda), params)
    for v in range(3):
        for b in range(2):
            p = params[v][b]
            ref = C.colour_view(batch[b], p["brightness"], p["contrast"], p["hue"], p["rects"])
            assert np.array_equal(views[v][b].cpu().numpy(), ref), (v, b)


def test_unlabeled_dataset_and_device_collator(cuda):
    """UnlabeledCrops + DeviceViewCollator == the reference's unlabeled item pipeline (transpart1 on the host, aug_k x
    transpart2, zero-padded collation) restated with PIL + the oracle, for a same-size and a ragged batch."""
    import random
    from PIL import Image
    import torchvision.transforms.functional as tf
    from mmt_psm_b200.data.unlabeled import DeviceViewCollator, UnlabeledCrops, reference_resize_size
    from mmt_psm_b200.data.unlabeled_views import draw_view_params
    from oracle import color_ref as C
    assert reference_resize_size(1000, 1000, 800, 1333) == (800, 800)
    assert reference_resize_size(2000, 1000, 800, 1333) == (666, 1332)      # the long side caps the scale
    rng = np.random.default_rng(8)
    for shapes in ([(100, 120)] * 2, [(100, 120), (90, 150)]):
        crops = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
        ds = UnlabeledCrops(crops, min_size=64, max_size=120, aug_k=3)
        random.seed(3)
        np.random.seed(3)
        views, idx = DeviceViewCollator(cuda, size_divisible=32)([ds[i] for i in range(len(crops))])
        random.seed(3)
        np.random.seed(3)
        ref_items = []
        for c in crops:   # the reference's order of draws: flip, then per view the colour parameters and rectangles
            img = Image.fromarray(c, "RGB")
            img = tf.resize(img, reference_resize_size(img.size[0], img.size[1], 64, 120))
            if random.random() < 0.5:
                img = tf.hflip(img)
            a = np.array(img)
            ref_items.append([C.colour_view(a, **{k: v for k, v in p.items()}) for p in
                              [draw_view_params(a.shape[0], a.shape[1]) for _ in range(3)]])
        assert idx == tuple(range(len(crops))) and len(views) == 3
        for k in range(3):
            v = views[k].cpu().numpy()
            assert v.shape[2] % 32 == 0 and v.shape[3] % 32 == 0
            for b, item in enumerate(ref_items):
                r = item[k]
                assert np.array_equal(v[b, :, :r.shape[1], :r.shape[2]], r), (k, b)
                assert not v[b, :, r.shape[1]:, :].any() and not v[b, :, :, r.shape[2]:].any()   # zero padding


# --------------------------------------------------------------------------------- RPN loss
def test_rpn_loss_fused_matches_reference_formula(cuda):
    """functional.rpn_loss over fused per-level head maps == the reference's formula (rpn/loss.py:138-196: BCE with
    logits over the sampled anchors, smooth-L1(beta = 1/9) over the positives, both / |sampled|) evaluated with torch
    ops on the flattened maps — values to 1e-6 relative, gradients w.r.t. the head maps to 1e-6 absolute."""
    import torch.nn.functional as Fn
    from mmt_psm_b200 import functional as Fk
    gen = torch.Generator().manual_seed(21)
    N, A, Cp = 2, 3, 32
    sizes = [(12, 10), (6, 5), (3, 3)]
    fused = [torch.randn(N, h, w, Cp, generator=gen).to(cuda).requires_grad_(True) for h, w in sizes]
    total = N * sum(h * w * A for h, w in sizes)
    labels = torch.randint(-1, 2, (total,), generator=gen).float().to(cuda)
    targets = (torch.randn(total, 4, generator=gen) * 0.3).to(cuda)
    r = torch.rand(total, generator=gen).to(cuda)
    pos = (labels == 1) & (r < 0.5)
    neg = (labels == 0) & (r < 0.4)
    lo, lb = Fk.rpn_loss(fused, A, labels, targets, pos, neg, 1.0 / 9)
    (2.0 * lo + 3.0 * lb).backward()
    got = [f.grad.clone() for f in fused]
    ref_in = [f.detach().clone().requires_grad_(True) for f in fused]
    obj = torch.cat([f[..., :A].reshape(N, -1) for f in ref_in], 1).reshape(-1)
    reg = torch.cat([f[..., A:5 * A].reshape(N, -1, 4) for f in ref_in], 1).reshape(-1, 4)
    sampled = pos | neg
    n = torch.abs(reg[pos] - targets[pos])
    ref_b = torch.where(n < 1.0 / 9, 0.5 * n ** 2 * 9, n - 0.5 / 9).sum() / sampled.sum()
    ref_o = Fn.binary_cross_entropy_with_logits(obj[sampled], labels[sampled])
    (2.0 * ref_o + 3.0 * ref_b).backward()
    assert abs(float(lo) - float(ref_o)) <= 1e-6 * abs(float(ref_o)) + 1e-7
    assert abs(float(lb) - float(ref_b)) <= 1e-6 * abs(float(ref_b)) + 1e-7
    for g, f in zip(got, ref_in):
        assert float((g - f.grad).abs().max()) <= 1e-6
        assert float(g[..., 5 * A:].abs().max()) == 0.0


# ------------------------------------------------------------------------- balanced sampler
def test_balanced_sample_counts_and_uniformity(cuda):
    """mmt_balanced_sample_f32: exact class sizes per image (min(#pos, quota), min(#neg, batch - taken)), only
    eligible rows selected, ignore rows never; and the draw is uniform: over many seeds every positive of an
    over-full image is picked with frequency quota / #pos (4-sigma band)."""
    kn = K()
    gen = torch.Generator().manual_seed(9)
    sizes = [5000, 37, 160000]
    labs = []
    labs.append(torch.where(torch.rand(5000, generator=gen) < 0.1, 1.0, 0.0))            # 500 pos > quota 128
    labs.append(torch.tensor([1.0] * 5 + [0.0] * 20 + [-1.0] * 12))                      # fewer than either quota
    big = torch.zeros(160000)
    big[torch.randperm(160000, generator=gen)[:60]] = 2.0                                 # 60 pos (< quota), label 2
    big[torch.randperm(160000, generator=gen)[:5000]] = -1.0
    labs.append(big)
    labels = torch.cat(labs).to(cuda)
    offs = torch.tensor([0, 5000, 5037, 165037], dtype=torch.int32, device=cuda)
    pos, neg = kn.balanced_sample(labels, offs, 3, 512, 128, seed=1234)
    pos, neg = pos.cpu(), neg.cpu()
    lab = labels.cpu()
    assert not (pos & (lab < 1)).any() and not (neg & (lab != 0)).any()
    a = 0
    for n, l in zip(sizes, labs):
        p, q = pos[a:a + n], neg[a:a + n]
        want_p = min(int((l >= 1).sum()), 128)
        want_n = min(int((l == 0).sum()), 512 - want_p)
        assert int(p.sum()) == want_p and int(q.sum()) == want_n, (n, int(p.sum()), want_p, int(q.sum()), want_n)
        a += n
    # different seeds give different draws; the same seed the same draw
    pos2, _ = kn.balanced_sample(labels, offs, 3, 512, 128, seed=1234)
    pos3, _ = kn.balanced_sample(labels, offs, 3, 512, 128, seed=99)
    assert torch.equal(pos2.cpu(), pos) and not torch.equal(pos3.cpu(), pos)
    # uniformity on image 0: 500 positives, 128 drawn, 400 seeds -> each picked ~ Binomial(400, 0.256)
    hits = torch.zeros(5000)
    l0 = labs[0].to(cuda)
    o0 = torch.tensor([0, 5000], dtype=torch.int32, device=cuda)
    for s in range(400):
        p, _ = kn.balanced_sample(l0, o0, 1, 512, 128, seed=1000 + s)
        hits += p.cpu().float()
    f = hits[labs[0] >= 1] / 400.0
    exp = 128.0 / float((labs[0] >= 1).sum())
    sigma = (exp * (1 - exp) / 400.0) ** 0.5
    assert abs(float(f.mean()) - exp) < 1e-6 and float((f - exp).abs().max()) < 4.5 * sigma, (float(f.min()), float(f.max()), exp, sigma)


def test_balanced_sample_matches_restated_draw(cuda):
    """The benchmarked sampler path, pinned draw for draw: `mmt_balanced_sample_f32` == oracle.ops_ref.balanced_sample_ref
    (the same counter-based keys, the n smallest of each class) on RPN-sized (160 K anchors, prefiltered key list), box-head
    sized (2 K proposals, whole class in shared memory), under-full and empty classes, for several seeds."""
    kn = K()
    gen = torch.Generator().manual_seed(10)
    labs = [torch.where(torch.rand(159882, generator=gen) < 0.002, 1.0, 0.0),     # ~320 pos > quota, ~159 K neg
            torch.where(torch.rand(2040, generator=gen) < 0.4, 2.0, 0.0),         # box head: 816 pos > 128, 1224 neg > 384
            torch.tensor([1.0] * 5 + [0.0] * 20 + [-1.0] * 12),                   # fewer than either quota
            torch.full((300,), -1.0)]                                            # nothing eligible
    labs[0][torch.randperm(159882, generator=gen)[:4000]] = -1.0
    sizes = [int(l.numel()) for l in labs]
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    labels = torch.cat(labs)
    offs = torch.tensor(offsets, dtype=torch.int32, device=cuda)
    for seed, batch, quota in ((1234, 256, 128), (0x9E3779B97F4A7C15, 512, 128), (7, 512, 128)):
        pos, neg = kn.balanced_sample(labels.to(cuda), offs, len(labs), batch, quota, seed=seed)
        rp, rn, unique = R.balanced_sample_ref(labels.numpy(), offsets, batch, quota, seed)
        assert unique          # no exactly-equal keys at a cut in these draws (a 2^-32 event per pair)
        assert np.array_equal(pos.cpu().numpy(), rp), (seed, int(pos.sum()), int(rp.sum()))
        assert np.array_equal(neg.cpu().numpy(), rn), (seed, int(neg.sum()), int(rn.sum()))


# --------------------------------------------------------------------- encode / FPN ROI table
def test_encode_boxes_and_fpn_rois_golden(cuda):
    kn = K()
    g = golden("glue")
    ref, prop = torch.from_numpy(g["coder_ref"]).to(cuda), torch.from_numpy(g["coder_prop"]).to(cuda)
    for name, w in (("rpn", (1., 1., 1., 1.)), ("box", (10., 10., 5., 5.))):
        enc = kn.encode_boxes(ref, None, prop, w).cpu().numpy()
        np.testing.assert_allclose(enc, g["coder_%s_enc" % name], rtol=2e-6, atol=2e-6)   # logf vs the CPU's log
    # gather form: shuffled ground truth addressed through matched indices; negative codes clamp to entry 0
    perm = torch.randperm(200, generator=torch.Generator().manual_seed(5))
    midx = torch.empty(200, dtype=torch.int64)
    midx[perm] = torch.arange(200)
    shuffled = ref.cpu()[perm]                       # shuffled[midx[i]] == ref[i]
    midx2 = midx.clone()
    midx2[:5] = torch.tensor([-1, -2, -1, -2, -1])
    enc = kn.encode_boxes(shuffled.to(cuda), midx2.to(cuda), prop, (1., 1., 1., 1.)).cpu().numpy()
    np.testing.assert_allclose(enc[5:], g["coder_rpn_enc"][5:], rtol=2e-6, atol=2e-6)
    want0 = kn.encode_boxes(shuffled[:1].repeat(5, 1).to(cuda), None, prop[:5], (1., 1., 1., 1.)).cpu().numpy()
    assert np.array_equal(enc[:5], want0)
    # ROI table + level mapper: golden levels of the reference's LevelMapper(2, 5); image index column
    boxes = torch.from_numpy(g["lvl_boxes"]).to(cuda)
    rois, levels = kn.fpn_rois(boxes, [100, 0, 205], 2.0, 5.0)
    assert np.array_equal(levels.cpu().numpy().astype(np.int64), g["lvl"])
    r = rois.cpu().numpy()
    assert np.array_equal(r[:, 1:], g["lvl_boxes"]) and np.array_equal(r[:, 0], np.r_[np.zeros(100), np.full(205, 2.0)])


# ------------------------------------------------------------------------------ fused matcher
def test_match_boxes_golden_and_matrix_path(cuda):
    """mmt_match_boxes_f32 == Matcher(boxlist_iou(gt, pr)) of the reference: the golden assignments for both Matcher
    configurations (glue.npz, produced by the reference), and the in-package matrix path on a larger random problem
    with duplicated boxes (ties) and a ground truth nobody overlaps (the IoU == 0 low-quality quirk)."""
    from mmt_psm_b200.modeling.matcher import Matcher
    from mmt_psm_b200.structures.bounding_box import BoxList
    from mmt_psm_b200.structures.boxlist_ops import boxlist_iou
    kn = K()
    g = golden("glue")
    gt, pr = torch.from_numpy(g["iou_gt"]).to(cuda), torch.from_numpy(g["iou_pr"]).to(cuda)
    for key, (hi, lo, lq) in (("match_rpn", (0.7, 0.3, True)), ("match_box", (0.5, 0.5, False))):
        got = kn.match_boxes(gt, pr, hi, lo, lq).cpu().numpy()
        assert np.array_equal(got, g[key]), key
    gen = torch.Generator().manual_seed(77)
    base = torch.rand(60, 2, generator=gen) * 700
    gts = torch.cat([base, base + 20 + torch.rand(60, 2, generator=gen) * 80], 1)
    gts[7] = gts[3]                                            # duplicate ground truth: arg-max tie
    gts[11] = torch.tensor([5000., 5000., 5040., 5040.])       # overlaps nothing
    ctr = torch.rand(30000, 2, generator=gen) * 800
    boxes = torch.cat([ctr, ctr + 8 + torch.rand(30000, 2, generator=gen) * 120], 1)
    boxes[:60] = gts                                           # exact hits
    a, b = BoxList(gts.to(cuda), (800, 800)), BoxList(boxes.to(cuda), (800, 800))
    for hi, lo, lq in ((0.7, 0.3, True), (0.5, 0.5, False)):
        m = Matcher(hi, lo, allow_low_quality_matches=lq)
        got = m.match(a, b)
        assert torch.equal(got, m(boxlist_iou(a, b)))
        assert np.array_equal(got.cpu().numpy(), R.matcher_ref(R.boxlist_iou_ref(gts, boxes), hi, lo, lq))
    with pytest.raises(ValueError):
        Matcher(0.5, 0.5).match(BoxList(torch.zeros(0, 4, device=cuda), (800, 800)), b)


# ------------------------------------------------------------------------- proposal decoding
def test_decode_clip_boxes_golden_and_oracle(cuda):
    kn = K()
    g = golden("glue")
    for name, w in (("rpn", (1., 1., 1., 1.)), ("box", (10., 10., 5., 5.))):
        codes = torch.from_numpy(g["coder_%s_enc" % name]) * 1.3 + 0.01
        boxes, s_nms, valid = kn.decode_clip_boxes(codes.to(cuda), torch.from_numpy(g["coder_prop"]).to(cuda), w,
                                                   R.math.log(1000. / 16))
        # expf on the device vs the reference's CPU exp: last-ulp differences only
        np.testing.assert_allclose(boxes.cpu().numpy(), g["coder_%s_dec" % name], rtol=2e-6, atol=2e-5)
        assert s_nms is None and valid is None
    gen = torch.Generator().manual_seed(41)
    N, per = 3, 700
    anchors = torch.rand(N * per, 4, generator=gen) * 300
    anchors[:, 2:] = anchors[:, :2] + torch.rand(N * per, 2, generator=gen) * 120
    deltas = torch.randn(N * per, 4, generator=gen) * torch.tensor([0.5, 0.5, 1.5, 1.5])
    deltas[::50, 2:] = 9.0                                     # exercises the dw / dh clamp
    scores = torch.rand(N * per, generator=gen)
    wh = torch.tensor([[320., 256.], [300., 300.], [200., 333.]])
    ref, ref_valid = R.decode_clip_ref(deltas, anchors, (1., 1., 1., 1.), wh, per, min_size=6.0)
    boxes, s_nms, valid = kn.decode_clip_boxes(deltas.to(cuda), anchors.to(cuda), (1., 1., 1., 1.), R.math.log(1000. / 16),
                                               wh.to(cuda), per, scores.to(cuda), 6.0)
    np.testing.assert_allclose(boxes.cpu().numpy(), ref, rtol=2e-6, atol=1e-3)
    v = valid.cpu().numpy()
    near = np.abs((ref[:, 2:] - ref[:, :2] + 1) - 6.0).min(1) < 1e-2      # predicate may flip only at the threshold
    assert np.array_equal(v[~near], ref_valid[~near])
    assert np.array_equal(s_nms.cpu().numpy(), np.where(v, scores.numpy(), np.float32(-1)))
    assert 0 < ref_valid.sum() < ref_valid.size


def test_rpn_topk_gather_matches_torch_topk(cuda):
    """mmt_rpn_topk_gather_f32 == per level `scores.topk(k, dim=1, sorted=True)` (reference rpn/inference.py:92-94) +
    the delta / anchor gathers: the selected SCORES are identical; the selected anchors are identical wherever scores are
    unique, and among equal scores the lowest anchor indices are taken in ascending order (torch leaves ties unspecified).
    Levels as at 800x800 (120000 .. 507 anchors), a row of all-equal scores, and a row with 5000 copies of the k-th value."""
    kn = K()
    gen = torch.Generator().manual_seed(44)
    A, Cp, N = 3, 16, 2
    hw = [200 * 200, 100 * 100, 50 * 50, 25 * 25, 13 * 13]
    scores, fused, anchors = [], [], []
    for l, p in enumerate(hw):
        s_ = torch.sigmoid(torch.randn(N, p * A, generator=gen) * 3 - 3)
        if l == 1:
            s_[1] = 0.5                                   # every score equal: the first k anchor indices win
        if l == 0:
            v = float(s_[0].sort(descending=True)[0][1999])
            pos = torch.randperm(p * A, generator=gen)[:5000]
            s_[0][pos] = torch.where(s_[0][pos] < v, torch.full_like(s_[0][pos], v), s_[0][pos])   # 5000 ties at the cut
        scores.append(s_.to(cuda))
        fused.append(torch.randn(N, p, Cp, generator=gen).to(cuda))
        anchors.append(torch.rand(p * A, 4, generator=gen).to(cuda))
    ks = [min(2000, p * A) for p in hw]
    K_ = sum(ks)
    deltas = torch.full((N, K_, 4), float("nan"), device=cuda)
    anch = torch.full((N, K_, 4), float("nan"), device=cuda)
    out_s = torch.full((N, K_), float("nan"), device=cuda)
    assert kn.rpn_topk_gather(scores, fused, A, anchors, ks, deltas, anch, out_s)
    col = 0
    for l, (p, k) in enumerate(zip(hw, ks)):
        vals, idx = scores[l].topk(k, dim=1, sorted=True)
        got_s = out_s[:, col:col + k]
        assert torch.equal(got_s, vals), l
        # recover the selected indices through the (unique) anchor rows
        for n in range(N):
            s_row = scores[l][n]
            # reference selection with the documented tie rule: sort by (score desc, index asc)
            key = (-s_row.double()).cpu().numpy()
            ref_idx = torch.from_numpy(np.lexsort((np.arange(p * A), key))[:k].copy()).to(cuda)
            assert torch.equal(anch[n, col:col + k], anchors[l][ref_idx]), (l, n)
            pix, a = ref_idx // A, ref_idx % A
            ref_d = torch.stack([fused[l][n, pix, A + 4 * a + c] for c in range(4)], 1)
            assert torch.equal(deltas[n, col:col + k], ref_d), (l, n)
        col += k
    assert not torch.isnan(out_s).any() and not torch.isnan(deltas).any()
    # capacity: k > 2048 is refused (the caller falls back to torch.topk + the gather kernel)
    assert kn.rpn_topk_gather(scores[:1], fused[:1], A, anchors[:1], [3000], torch.empty(N, 3000, 4, device=cuda),
                              torch.empty(N, 3000, 4, device=cuda), torch.empty(N, 3000, device=cuda)) is False


# ------------------------------------------------------------------------------------ convs
CONV_CASES = [
    # B, H, W, Cin, Cout, k, stride, pad
    (2, 25, 25, 64, 64, 1, 1, 0),
    (2, 50, 50, 256, 64, 1, 1, 0),
    (1, 56, 40, 64, 64, 3, 1, 1),
    (2, 25, 25, 128, 128, 3, 1, 1),
    (2, 13, 13, 256, 256, 3, 1, 1),
    (2, 50, 50, 256, 512, 1, 2, 0),
    (2, 25, 27, 512, 128, 1, 2, 0),
    (3, 14, 14, 256, 256, 3, 1, 1),
    (1, 20, 20, 256, 32, 1, 1, 0),
    (2, 9, 9, 2048, 256, 1, 1, 0),
    (1, 1, 1024, 12544, 1024, 1, 1, 0),   # FC6 as the box head runs it: 1024 ROIs x (7*7*256) -> 1024 (392 K blocks / tile)
    (5, 14, 14, 256, 1024, 1, 1, 0),      # mask predictor's ConvTranspose2d(256, 256, 2, 2) as ONE 1x1 conv with 4*256 outputs
]
# heads whose channel count is padded to the TMA granularity (15 -> 16: fused RPN cls+bbox, fused box predictor)
HEAD_CASES = [
    (2, 50, 50, 256, 16, 1, 1, 0),
    (1, 1, 1000, 1024, 16, 1, 1, 0),
]


def _conv_inputs(case, gen, cuda):
    B, H, W, Ci, Co, k, s, p = case
    x = torch.randn(B, Ci, H, W, generator=gen)
    w = torch.randn(Co, Ci, k, k, generator=gen) / (Ci * k * k) ** 0.5
    return x, w


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_conv_fprop(cuda, impl):
    import torch.nn.functional as Fn
    kn = K()
    gen = torch.Generator().manual_seed(28)
    for case in CONV_CASES:
        B, H, W, Ci, Co, k, s, p = case
        x, w = _conv_inputs(case, gen, cuda)
        scale = torch.rand(Co, generator=gen) + 0.5
        bias = torch.randn(Co, generator=gen)
        shape = kn.conv_shape(B, H, W, Ci, Co, k, s, p)
        res = torch.randn(B, Co, shape.out_h, shape.out_w, generator=gen)
        xd = x.permute(0, 2, 3, 1).contiguous().to(cuda)
        wd = w.permute(0, 2, 3, 1).contiguous().to(cuda)
        rd = res.permute(0, 2, 3, 1).contiguous().to(cuda)
        y = kn.conv_fprop(xd, wd, shape, scale.to(cuda), bias.to(cuda), rd, True,
                          kn.CONV_SIMT if impl == "simt" else kn.CONV_TCGEN05)
        y = y.permute(0, 3, 1, 2).cpu()
        exact = R.conv_bn_act_ref(x, w, scale, bias, res, True, s, p, operands="fp32")
        if impl == "simt":
            assert _rel(y, exact) < 1e-5, (case, _rel(y, exact))
        else:
            trunc = R.conv_bn_act_ref(x, w, scale, bias, res, True, s, p, operands="tf32")
            e_t, e_x = _rel(y, trunc), _rel(y, exact)
            print("conv_fprop tc", case, "vs tf32-truncated %.2e  vs fp32 %.2e" % (e_t, e_x))
            assert e_x < 3e-3, (case, e_x)          # TF32 operands, fp32 accumulate
            assert e_t < 1e-4 or e_x < 1e-3, (case, e_t, e_x)
        # no-epilogue variant
        y2 = kn.conv_fprop(xd, wd, shape, None, None, None, False,
                           kn.CONV_SIMT if impl == "simt" else kn.CONV_TCGEN05).permute(0, 3, 1, 2).cpu()
        assert _rel(y2, Fn.conv2d(x, w, None, s, p)) < (1e-5 if impl == "simt" else 3e-3)


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_conv_dgrad_wgrad(cuda, impl):
    import torch.nn.functional as Fn
    kn = K()
    gen = torch.Generator().manual_seed(29)
    flag = kn.CONV_SIMT if impl == "simt" else kn.CONV_TCGEN05
    tol = 1e-5 if impl == "simt" else 3e-3
    for case in CONV_CASES:
        B, H, W, Ci, Co, k, s, p = case
        x, w = _conv_inputs(case, gen, cuda)
        shape = kn.conv_shape(B, H, W, Ci, Co, k, s, p)
        dy = torch.randn(B, Co, shape.out_h, shape.out_w, generator=gen)
        xr = x.double().requires_grad_(True)
        wr = w.double().requires_grad_(True)
        Fn.conv2d(xr, wr, None, s, p).backward(dy.double())
        xd = x.permute(0, 2, 3, 1).contiguous().to(cuda)
        wd = w.permute(0, 2, 3, 1).contiguous().to(cuda)
        dyd = dy.permute(0, 2, 3, 1).contiguous().to(cuda)
        dx = kn.conv_dgrad(dyd, wd, shape, flag).permute(0, 3, 1, 2).cpu()
        assert _rel(dx, xr.grad.float()) < tol, ("dgrad", case, _rel(dx, xr.grad.float()))
        dw, db = kn.conv_wgrad(xd, dyd, shape, True, flag)
        dw = dw.permute(0, 3, 1, 2).cpu()
        assert _rel(dw, wr.grad.float()) < tol, ("wgrad", case, _rel(dw, wr.grad.float()))
        assert _rel(db.cpu(), dy.sum((0, 2, 3))) < 1e-5


PAIR_CASES = [
    # B, H, W, Cin, Cout, k, stride, pad              pixel tiles (128 px)            column tiles
    (1, 48, 40, 128, 256, 3, 1, 1),       # 15 (odd: the last pair computes its tile twice)     1 x 256
    (1, 29, 30, 256, 384, 1, 1, 0),       # 7 flat-pixel tiles (odd)           2 x 256, the second half empty
    (2, 50, 50, 256, 512, 1, 2, 0),       # strided 1x1, 10 tiles              2 x 256
    (2, 25, 25, 128, 128, 3, 1, 1),       # BN = 128 pair (64 filter rows per CTA)
    (4, 50, 50, 256, 256, 3, 1, 1),       # 112 tiles: more pairs than one round of 74 clusters -> persistent loop, both TMEM buffers
    (1, 1, 1024, 12544, 1024, 1, 1, 0),   # FC6: 392 K blocks per tile
]


@pytest.mark.parametrize("cg", [1, 2])
def test_conv_cta_pair_matches_reference(cuda, cg):
    """The cta_group::2 kernel (256 x BN tile over a CTA pair, leader-issued MMA, multicast commits, peer-signalled TMA)
    against the TF32-truncated fp64 reference and against the one-CTA kernel: fprop with the full epilogue (scale, shift,
    residual, ReLU) and dgrad (same kernel, transposed filter)."""
    import torch.nn.functional as Fn
    kn = K()
    gen = torch.Generator().manual_seed(34)
    try:
        for case in PAIR_CASES:
            B, H, W, Ci, Co, k, s, p = case
            x, w = _conv_inputs(case, gen, cuda)
            scale = torch.rand(Co, generator=gen) + 0.5
            bias = torch.randn(Co, generator=gen)
            shape = kn.conv_shape(B, H, W, Ci, Co, k, s, p)
            res = torch.randn(B, Co, shape.out_h, shape.out_w, generator=gen)
            dy = torch.randn(B, Co, shape.out_h, shape.out_w, generator=gen)
            xd = x.permute(0, 2, 3, 1).contiguous().to(cuda)
            wd = w.permute(0, 2, 3, 1).contiguous().to(cuda)
            rd = res.permute(0, 2, 3, 1).contiguous().to(cuda)
            dyd = dy.permute(0, 2, 3, 1).contiguous().to(cuda)
            kn.conv_force_cta_group(cg)
            y = kn.conv_fprop(xd, wd, shape, scale.to(cuda), bias.to(cuda), rd, True, kn.CONV_TCGEN05)
            dx = kn.conv_dgrad(dyd, wd, shape, kn.CONV_TCGEN05)
            torch.cuda.synchronize()
            kn.conv_force_cta_group(1)
            y1 = kn.conv_fprop(xd, wd, shape, scale.to(cuda), bias.to(cuda), rd, True, kn.CONV_TCGEN05)
            dx1 = kn.conv_dgrad(dyd, wd, shape, kn.CONV_TCGEN05)
            # same operands, same K order inside a tile: the pair kernel reproduces the one-CTA kernel to accumulation noise
            assert _rel(y, y1) < 2e-6, (case, "fprop pair vs single", _rel(y, y1))
            assert _rel(dx, dx1) < 2e-6, (case, "dgrad pair vs single", _rel(dx, dx1))
            trunc = R.conv_bn_act_ref(x, w, scale, bias, res, True, s, p, operands="tf32")
            assert _rel(y.permute(0, 3, 1, 2).cpu(), trunc) < 1e-4, (case, _rel(y.permute(0, 3, 1, 2).cpu(), trunc))
            if Ci * k * k <= 4096:
                xr = x.double().requires_grad_(True)
                Fn.conv2d(xr, w.double(), None, s, p).backward(dy.double())
                assert _rel(dx.permute(0, 3, 1, 2).cpu(), xr.grad.float()) < 3e-3, case
    finally:
        kn.conv_force_cta_group(0)


def test_conv_fused_backward_ops(cuda):
    """mmt_conv_dgrad_fused_f32 / mmt_conv_wgrad_scaled_f32 (BN scale folded into the transformed filter / the wgrad
    epilogue; skip-add or ReLU mask in the dgrad epilogue) against their definition composed from exact fp64 pieces."""
    import torch.nn.functional as Fn
    kn = K()
    gen = torch.Generator().manual_seed(35)
    cases = [(2, 25, 25, 128, 128, 3, 1, 1), (2, 50, 50, 256, 64, 1, 1, 0), (2, 50, 50, 256, 512, 1, 2, 0),
             (1, 48, 40, 128, 256, 3, 1, 1), (4, 50, 50, 1024, 256, 1, 1, 0)]
    for case in cases:
        B, H, W, Ci, Co, k, s, p = case
        x, w = _conv_inputs(case, gen, cuda)
        shape = kn.conv_shape(B, H, W, Ci, Co, k, s, p)
        dy = torch.randn(B, Co, shape.out_h, shape.out_w, generator=gen)
        scale = torch.rand(Co, generator=gen) + 0.5
        aux = torch.randn(B, Ci, H, W, generator=gen)
        xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
        Fn.conv2d(xr, wr, None, s, p).backward(dy.double() * scale.double().view(1, -1, 1, 1))
        dx_ref, dw_ref = xr.grad.float(), wr.grad.float()
        xd = x.permute(0, 2, 3, 1).contiguous().to(cuda)
        wd = w.permute(0, 2, 3, 1).contiguous().to(cuda)
        dyd = dy.permute(0, 2, 3, 1).contiguous().to(cuda)
        auxd = aux.permute(0, 2, 3, 1).contiguous().to(cuda)
        nchw = lambda t: t.permute(0, 3, 1, 2).cpu()
        got = nchw(kn.conv_dgrad_fused(dyd, wd, shape, scale.to(cuda)))
        assert _rel(got, dx_ref) < 3e-3, ("scaled dgrad", case, _rel(got, dx_ref))
        if s == 1:
            got = nchw(kn.conv_dgrad_fused(dyd, wd, shape, scale.to(cuda), auxd, kn.DGRAD_AUX_ADD))
            assert _rel(got, dx_ref + aux) < 3e-3, ("dgrad + aux", case)
            got = nchw(kn.conv_dgrad_fused(dyd, wd, shape, scale.to(cuda), auxd, kn.DGRAD_AUX_MASK))
            ref = torch.where(aux > 0, dx_ref, torch.zeros_like(dx_ref))
            assert _rel(got, ref) < 3e-3, ("masked dgrad", case)
            assert bool((got[aux <= 0] == 0).all())            # the mask is exact, not approximately zero
        dw = kn.conv_wgrad_scaled(xd, dyd, shape, scale.to(cuda)).permute(0, 3, 1, 2).cpu()
        assert _rel(dw, dw_ref) < 3e-3, ("scaled wgrad", case, _rel(dw, dw_ref))
        acc = torch.full((Co, k, k, Ci), 0.25, device=cuda)
        assert kn.conv_wgrad_scaled(xd, dyd, shape, scale.to(cuda), into=acc) is None
        assert _rel(acc.permute(0, 3, 1, 2).cpu() - 0.25, dw_ref) < 3e-3, ("accumulating scaled wgrad", case)


@pytest.mark.parametrize("downsample,stride", [(False, 1), (True, 2)])
def test_bottleneck_fused_backward_matches_layerwise(cuda, downsample, stride):
    """One bottleneck, forward + backward: the block-level autograd node with the fused backward kernels against the
    layer-by-layer autograd path (same forward kernels; backward differs in WHERE the BN scale meets TF32 rounding)."""
    from mmt_psm_b200.modeling.backbone import BottleneckWithFixedBatchNorm
    torch.manual_seed(3)
    cin, mid, cout = (256, 128, 512) if downsample else (512, 128, 512)
    blk = BottleneckWithFixedBatchNorm(cin, mid, cout, 1, True, stride).to(cuda)
    for m in blk.modules():
        if hasattr(m, "running_var"):
            m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.1); m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    x0 = torch.randn(4, 50, 50, cin, device=cuda).relu()
    gy = torch.randn(4, 50 // stride, 50 // stride, cout, device=cuda)
    outs = []
    for fused in (True, False):
        blk.fused_backward = fused
        for p_ in blk.parameters():
            p_.grad = None
        x = x0.clone().requires_grad_(True)
        y = blk(x)
        y.backward(gy)
        outs.append((y.detach(), x.grad, {n: p_.grad.clone() for n, p_ in blk.named_parameters()}))
    blk.fused_backward = True
    (y0, dx0, g0), (y1, dx1, g1) = outs
    assert torch.equal(y0, y1)
    assert _rel(dx0, dx1) < 3e-3, _rel(dx0, dx1)
    for n in g0:
        assert _rel(g0[n], g1[n]) < 3e-3, (n, _rel(g0[n], g1[n]))


def test_conv_padded_heads_tc(cuda):
    """16-channel (15 padded) heads: fprop + wgrad on the tcgen05 kernels (BN = 32 tile, TMA clips the stores); dgrad of
    these shapes (GEMM K = 16) is routed to the FFMA kernel by CONV_AUTO and must agree as well."""
    import torch.nn.functional as Fn
    kn = K()
    gen = torch.Generator().manual_seed(33)
    for case in HEAD_CASES:
        B, H, W, Ci, Co, k, s, p = case
        x, w = _conv_inputs(case, gen, cuda)
        w[15:] = 0                                     # the padding row
        bias = torch.randn(Co, generator=gen)
        shape = kn.conv_shape(B, H, W, Ci, Co, k, s, p)
        dy = torch.randn(B, Co, shape.out_h, shape.out_w, generator=gen)
        xd = x.permute(0, 2, 3, 1).contiguous().to(cuda)
        wd = w.permute(0, 2, 3, 1).contiguous().to(cuda)
        dyd = dy.permute(0, 2, 3, 1).contiguous().to(cuda)
        y = kn.conv_fprop(xd, wd, shape, None, bias.to(cuda), None, False, kn.CONV_TCGEN05).permute(0, 3, 1, 2).cpu()
        ref = R.conv_bn_act_ref(x, w, None, bias, None, False, s, p, operands="tf32")
        assert _rel(y, ref) < 1e-4, (case, _rel(y, ref))
        xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
        Fn.conv2d(xr, wr, None, s, p).backward(dy.double())
        dw, db = kn.conv_wgrad(xd, dyd, shape, True, kn.CONV_TCGEN05)
        assert _rel(dw.permute(0, 3, 1, 2).cpu(), wr.grad.float()) < 3e-3, case
        assert _rel(db.cpu(), dy.sum((0, 2, 3))) < 1e-5
        dx = kn.conv_dgrad(dyd, wd, shape, kn.CONV_AUTO).permute(0, 3, 1, 2).cpu()
        assert _rel(dx, xr.grad.float()) < 3e-3, case


def test_conv_stem_7x7_simt(cuda):
    import torch.nn.functional as Fn
    kn = K()
    gen = torch.Generator().manual_seed(30)
    x = torch.randn(2, 3, 64, 96, generator=gen)
    w = torch.randn(64, 3, 7, 7, generator=gen) * 0.1
    shape = kn.conv_shape(2, 64, 96, 3, 64, 7, 2, 3)
    y = kn.conv_fprop(x.permute(0, 2, 3, 1).contiguous().to(cuda), w.permute(0, 2, 3, 1).contiguous().to(cuda), shape,
                      None, None, None, True, kn.CONV_AUTO)
    assert _rel(y.permute(0, 3, 1, 2).cpu(), torch.relu(Fn.conv2d(x, w, None, 2, 3))) < 1e-5


def test_conv_stem_7x7_tensor_core(cuda):
    """The space-to-depth tensor-core stem == Conv2d(3, 64, 7, 2, 3) + frozen-BN affine + ReLU (TF32 operands), on an
    even and an odd-sized image (the odd last row / column pairs with a zero)."""
    import torch.nn.functional as Fn
    kn = K()
    gen = torch.Generator().manual_seed(31)
    for (B, H, W) in ((2, 64, 96), (1, 67, 95)):
        x = torch.randn(B, 3, H, W, generator=gen) * 50
        w = torch.randn(64, 3, 7, 7, generator=gen) / (3 * 49) ** 0.5
        sc, sh = torch.rand(64, generator=gen) * 0.02 + 0.01, torch.randn(64, generator=gen) * 0.1
        shape = kn.conv_shape(B, H, W, 3, 64, 7, 2, 3)
        y = kn.conv_stem_fprop(x.permute(0, 2, 3, 1).contiguous().to(cuda), w.permute(0, 2, 3, 1).contiguous().to(cuda),
                               shape, sc.to(cuda), sh.to(cuda), relu=True).permute(0, 3, 1, 2).cpu()
        ref = torch.relu(Fn.conv2d(x.double(), w.double(), None, 2, 3) * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1))
        assert y.shape == ref.shape, (y.shape, ref.shape)
        assert _rel(y.double(), ref) < 3e-3, _rel(y.double(), ref)
        # the same stem reading the reference's NCHW batch directly: bit-identical to the NHWC entry
        y2 = kn.conv_stem_fprop(x.contiguous().to(cuda), w.permute(0, 2, 3, 1).contiguous().to(cuda), shape, sc.to(cuda),
                                sh.to(cuda), relu=True, nchw=True).permute(0, 3, 1, 2).cpu()
        assert torch.equal(y, y2)


def test_conv_full_size_tc_vs_simt(cuda):
    """BASELINE-size layer (fpn_layer1: 2x200x200x256 -> 256, 3x3): tensor-core path vs the fp32 SIMT kernel."""
    kn = K()
    gen = torch.Generator(device="cuda").manual_seed(31)
    x = torch.randn(2, 200, 200, 256, generator=gen, device=cuda)
    w = torch.randn(256, 3, 3, 256, generator=gen, device=cuda) / 48
    b = torch.randn(256, generator=gen, device=cuda)
    shape = kn.conv_shape(2, 200, 200, 256, 256, 3, 1, 1)
    y_tc = kn.conv_fprop(x, w, shape, None, b, None, False, kn.CONV_TCGEN05)
    y_ref = kn.conv_fprop(x, w, shape, None, b, None, False, kn.CONV_SIMT)
    assert _rel(y_tc, y_ref) < 3e-3
    # linearity (size-independent property): conv(2x) - 2 conv(x) == -bias
    y2 = kn.conv_fprop(2 * x, w, shape, None, b, None, False, kn.CONV_TCGEN05)
    assert _rel(y2 - 2 * y_tc, (-b).expand_as(y_tc)) < 1e-3


# ------------------------------------------------------------------------- mask targets / paste
def _mask_target_fixture():
    g = golden("mask_targets")
    xy, poly_len, inst_npoly = g["xy"], g["poly_len"], g["inst_npoly"]
    polys, pos = [], 0
    for n in poly_len:
        polys.append(torch.from_numpy(xy[pos:pos + 2 * int(n)].astype(np.float32)))
        pos += 2 * int(n)
    inst, pi = [], 0
    for n in inst_npoly:
        inst.append(polys[pi:pi + int(n)])
        pi += int(n)
    return inst, g["which"], g["boxes"], g["masks"]


def test_rasterize_polygons_bit_exact_reference_golden(cuda):
    """mmt_rasterize_polygons_f32 == the reference's project_masks_on_boxes + pycocotools (mask_head/loss.py:37-75,
    segmentation_mask.py:94-137, pycoco/maskApi.c:166-206) on all 96 golden ROIs — bit for bit, incl. the sub-pixel
    box, the integer box and two-polygon instances — and == oracle.polygon_ref on 300 more random ROIs."""
    from mmt_psm_b200.structures.segmentation_mask import SegmentationMask
    from oracle import polygon_ref
    kn = K()
    inst, which, boxes, masks = _mask_target_fixture()
    sm = SegmentationMask(inst, (800, 800))
    got = kn.rasterize_polygons(sm.packed(cuda), torch.from_numpy(boxes).to(cuda), torch.from_numpy(which).to(cuda), 28)
    got = got.cpu().numpy()
    assert got.shape == (96, 28, 28) and set(np.unique(got)) <= {0.0, 1.0}
    bad = [i for i in range(96) if not np.array_equal(got[i].astype(np.uint8), masks[i])]
    assert not bad, bad
    # random ROIs (boxes partly outside the instance, tiny and huge boxes, image-border clamps) against the oracle
    gen = torch.Generator().manual_seed(91)
    K2 = 300
    w2 = torch.randint(0, len(inst), (K2,), generator=gen)
    b2 = []
    for i in w2.tolist():
        pts = torch.cat(inst[i]).view(-1, 2)
        lo, hi = pts.min(0)[0], pts.max(0)[0]
        jit = (torch.rand(4, generator=gen) - 0.5) * 1.2 * torch.cat([hi - lo, hi - lo])
        b2.append((torch.cat([lo, hi]) + jit).clamp(0, 799))
    b2 = torch.stack(b2)
    b2[:, 2:] = torch.max(b2[:, 2:], b2[:, :2])
    b2[::7] = b2[::7].round()
    got2 = kn.rasterize_polygons(sm.packed(cuda), b2.to(cuda), w2.to(cuda), 28).cpu().numpy()
    ref2 = polygon_ref.project_masks_on_boxes_ref([[p.numpy() for p in inst[i]] for i in w2.tolist()], b2.numpy(), 28)
    bad = [i for i in range(K2) if not np.array_equal(got2[i], ref2[i])]
    assert not bad, bad


def test_paste_masks_sum_vs_oracle(cuda):
    """mmt_paste_masks_sum_f32 == Masker paste (pad 1, expand, int box, bilinear resize, > 0.5, clip to the image) summed
    over the detections of each image (mask_head/inference.py:169-231, generalized_rcnn.py:129-132) as restated by
    oracle.model_cpu.paste_sum: integral foreground map identical except where the interpolated probability sits within
    float rounding of the 0.5 threshold (<= 1e-4 of the pixels), boxes partly / wholly outside the image, 1-pixel boxes,
    and an image without detections."""
    from oracle.model_cpu import paste_sum
    kn = K()
    gen = torch.Generator().manual_seed(92)
    H, W, M, n_img = 300, 420, 28, 3
    D = 60
    c = torch.rand(D, 2, generator=gen) * torch.tensor([W, H])
    wh = torch.rand(D, 2, generator=gen) * 120 + 2
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
    boxes[0] = torch.tensor([-30., -20., 40., 35.])       # sticks out top-left
    boxes[1] = torch.tensor([400., 280., 470., 330.])     # sticks out bottom-right
    boxes[2] = torch.tensor([500., 100., 560., 150.])     # wholly outside
    boxes[3] = torch.tensor([100.3, 100.6, 100.9, 101.2])  # sub-pixel
    yy, xx = torch.meshgrid(torch.arange(M, dtype=torch.float32), torch.arange(M, dtype=torch.float32), indexing="ij")
    r = ((xx - 13.5) ** 2 + (yy - 13.5) ** 2).sqrt()
    probs = torch.sigmoid((torch.rand(D, 1, 1, generator=gen) * 8 + 5 - r) * 0.7) * (0.6 + 0.4 * torch.rand(D, M, M, generator=gen))
    img = torch.randint(0, 2, (D,), generator=gen)        # image 2 gets no detection
    got = kn.paste_masks_sum(probs.to(cuda), boxes.to(cuda), img.to(cuda), n_img, H, W).cpu()
    assert got.shape == (n_img, H, W) and float(got[2].abs().sum()) == 0.0
    total_diff = 0
    for i in range(2):
        sel = img == i
        ref = paste_sum(probs[sel], boxes[sel], H, W)
        total_diff += int((got[i] != ref).sum())
        assert float((got[i] - ref).abs().max()) <= 1.0      # a disagreeing pixel differs by ONE detection's vote
        assert float(ref.max()) >= 2                              # overlapping detections are summed, not OR-ed
    assert total_diff <= 1e-4 * 2 * H * W, total_diff
