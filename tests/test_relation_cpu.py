"""CPU suite: host-side logic of the IR-Net relation modules (modeling/relation.py — device-agnostic torch code around the
CUDA entry points) against the oracle's restatement of the reference (oracle/relation_ref.py, itself pinned bit-exact to the
patched reference by tests/golden/relation.npz): label preparation, embeddings, row order, masked instance attention, cfg."""
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)
import mmt_psm_b200  # noqa: E402,F401

def test_relation_label_preparation_matches_numpy_restatement():
    """`iou_regression_targets` (device-agnostic torch, no read-back) against the oracle's restatement of the reference's
    numpy argmax / intersect1d label preparation (relation_module.py:333-473), incl. empty classes and the rank-0 fallback."""
    import torch
    from mmt_psm_b200.modeling.relation import iou_regression_targets
    from oracle import relation_ref as RR
    gen = torch.Generator().manual_seed(3)
    hits = 0
    for case in range(25):
        M = int(torch.randint(0, 30, (1,), generator=gen))
        N = 90 if case % 3 else int(torch.randint(1, 90, (1,), generator=gen))
        ctr = torch.rand(M, 2, generator=gen) * 600 + 50
        wh = torch.rand(M, 2, generator=gen) * 80 + 10
        gt = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        gl = torch.randint(1, 3, (M,), generator=gen) if case % 4 else torch.full((M,), 2, dtype=torch.int64)
        if M:
            own = torch.randint(0, M, (N, 2), generator=gen)
            boxes = gt[own] + torch.randn(N, 2, 4, generator=gen) * (2 + case % 6)
        else:
            boxes = torch.rand(N, 2, 4, generator=gen) * 300
        boxes[..., 2:] = torch.max(boxes[..., 2:], boxes[..., :2] + 2)
        scores = torch.sort(torch.rand(N, 2, generator=gen), 0, descending=True)[0]
        for binary in (False, True):
            ref = RR.reg_labels(boxes, scores, gt, gl, (0.1, 0.6), binary=binary)
            got = iou_regression_targets(boxes, scores, gt, gl, (0.1, 0.6), binary=binary).numpy()
            assert np.array_equal(got, ref), (case, binary)
        hits += int((ref > 0).sum())
    assert hits > 100


def test_relation_embeddings_and_row_order_match_oracle():
    import torch
    from mmt_psm_b200.modeling.relation import (CIAM_Module, MaskRelationRefineNet, extract_multi_position_matrix,
                                                 extract_rank_embedding)
    from mmt_psm_b200.config.presets import mt_psm_cfg
    from mmt_psm_b200.structures.bounding_box import BoxList
    from oracle import relation_ref as RR
    gen = torch.Generator().manual_seed(9)
    b = torch.rand(2, 17, 2, 4, generator=gen) * 200
    b[..., 2:] += b[..., :2]
    for i in range(2):   # leading image dimension = the oracle's per-image call
        assert torch.equal(extract_multi_position_matrix(b, None, 64, 1000)[i], RR.position_embedding(b[i], 64))
    assert torch.equal(extract_rank_embedding(90, 1024), RR.rank_embedding(90, 1024))
    # row order of the mask refinement: per image class 1 by objectness (descending), then class 2
    props, labels, objs = [], [], []
    for n in (9, 0, 6):
        p = BoxList(torch.rand(n, 4, generator=gen) * 100, (256, 256))
        p.add_field("labels", torch.randint(1, 3, (n,), generator=gen))
        p.add_field("objectness", torch.rand(n, generator=gen))
        props.append(p)
    order, gid, sl = MaskRelationRefineNet.sort_rows(props, 2)
    off, want = 0, []
    for p in props:
        if len(p):
            o, _ = RR.sort_by_class(p.get_field("labels"), p.get_field("objectness"), 2)
            want.append(o + off)
        off += len(p)
    assert torch.equal(order, torch.cat(want))
    assert torch.all(gid[1:] >= gid[:-1]) and torch.equal(sl, torch.cat([p.get_field("labels") for p in props])[order])
    # masked instance attention over all rows == the reference's per-group attention
    ciam = CIAM_Module(mt_psm_cfg(relation=True))
    with torch.no_grad():
        ciam.gamma.fill_(0.7)
    x = torch.randn(len(order), 14, 14, 16, generator=gen)
    got = ciam(x, gid)
    for g in gid.unique():
        rows = (gid == g).nonzero()[:, 0]
        ref = RR.ciam(x[rows].permute(0, 3, 1, 2).contiguous(), ciam.gamma.detach()).permute(0, 2, 3, 1)
        assert torch.allclose(got[rows], ref, rtol=1e-4, atol=1e-4)


def test_relation_cfg_of_the_reference_yaml():
    from mmt_psm_b200.config.presets import mt_psm_cfg
    from mmt_psm_b200.modeling.detector import build_detection_model
    c = mt_psm_cfg(relation=True)
    r, m = c.MODEL.RELATION_NMS, c.MODEL.RELATION_MASK
    assert (r.FIRST_N, tuple(r.THREAD), tuple(r.HID_DIM), r.TOPK, r.FG_THREAD, r.POS_NMS, r.REG_IOU) == \
        (90, (0.1,), (1024, 1024, 128), 40, 0.1, 0.55, True)
    assert (m.TYPE, m.EXTRACTOR_CHANNEL, m.NORM, m.DEEP_SUPER, m.USE_RELATION) == ("CIAM", 16, -1, True, True)
    model = build_detection_model(c, is_student=True)
    assert len(model.state_dict()) == 349     # every entry of the reference's state dict
    assert len(build_detection_model(mt_psm_cfg(relation=False), is_student=True).state_dict()) == 349 - 27
