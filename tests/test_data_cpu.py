"""CPU suite: the labeled-crop item pipeline (data/labeled.py: host-side Resize / flip of image AND target, parameter draws
in the reference's order) against items produced by the reference's own transform classes and BoxList / SegmentationMask
containers (tests/golden/source_items.npz); the pixels are rendered here with the oracle's Pillow-exact colour pipeline
(oracle/color_ref.py) — the device rendering of the same parameters is checked in the GPU suite."""
import random
import sys

import numpy as np
import torch

from conftest import ROOT, golden

sys.path.insert(0, ROOT)
import mmt_psm_b200  # noqa: E402,F401


def labeled_dataset(g, **kw):
    from mmt_psm_b200.data.labeled import LabeledCrops
    crops, ann = [], []
    for i in range(int(g["n"])):
        crops.append(g["img_%d" % i])
        flat, lens, per = g["poly_flat_%d" % i], g["poly_len_%d" % i], g["poly_%d" % i]
        pieces, off = [], 0
        for n in lens:
            pieces.append(flat[off:off + n].tolist())
            off += n
        polys, k = [], 0
        for c in per:
            polys.append(pieces[k:k + c])
            k += c
        ann.append({"boxes": g["boxes_%d" % i], "labels": g["labels_%d" % i], "polygons": polys})
    return LabeledCrops(crops, ann, min_size=64, max_size=120, **kw)


def test_labeled_items_match_reference_transforms():
    from oracle import color_ref as C
    g = golden("source_items")
    ds = labeled_dataset(g)
    random.seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    flips = 0
    for i in range(len(ds)):
        arr, target, params, idx = ds[i]
        assert idx == i and tuple(target.size) == tuple(g["out_size_%d" % i]) == (arr.shape[1], arr.shape[0])
        np.testing.assert_allclose(target.bbox.numpy(), g["out_boxes_%d" % i], rtol=0, atol=1e-4)
        assert np.array_equal(target.get_field("labels").numpy(), g["labels_%d" % i])
        flat = np.concatenate([p.numpy().astype(np.float64) for m in target.get_field("masks").polygons for p in m.polygons])
        np.testing.assert_allclose(flat, g["out_poly_flat_%d" % i], rtol=0, atol=1e-4)
        image = C.colour_view(arr, params["brightness"], params["contrast"], params["hue"], params["rects"])
        assert np.array_equal(image, g["out_image_%d" % i]), i
        flips += int(target.bbox[0, 0] != torch.as_tensor(g["boxes_%d" % i][0, 0] * arr.shape[1] / g["img_%d" % i].shape[1]))
    assert 0 < flips < len(ds)        # the fixture holds flipped and un-flipped items


def test_labeled_loader_sampler_chain():
    """The loader's batch plan: aspect-ratio groups never mix inside a batch, exactly max_iter batches, resumable, and
    rank shards are disjoint — checked on the sampler chain (the device collation is covered by the GPU suite)."""
    from mmt_psm_b200.data.loaders import LabeledCropLoader, aspect_groups
    g = golden("source_items")
    ds = labeled_dataset(g)
    sizes = [(g["img_%d" % i].shape[1], g["img_%d" % i].shape[0]) for i in range(len(ds))]
    groups = aspect_groups(sizes)
    assert groups == [0, 0, 1, 0]                         # (h / w): 100/120, 90/150, 128/96, 100/120
    ld = LabeledCropLoader(ds, "cpu", images_per_batch=2, max_iter=7, shuffle=False)
    plan = list(ld.batch_sampler)
    assert len(ld) == len(plan) == 7 and plan[:3] == [[0, 1], [2], [3]] and plan[3] == [0, 1]
    assert all(len({groups[i] for i in b}) == 1 for b in plan)
    resumed = list(LabeledCropLoader(ds, "cpu", images_per_batch=2, max_iter=7, start_iter=5, shuffle=False).batch_sampler)
    assert len(resumed) == 2
    a = LabeledCropLoader(ds, "cpu", images_per_batch=1, max_iter=2, world=2, rank=0, aspect_grouping=False).batch_sampler
    b = LabeledCropLoader(ds, "cpu", images_per_batch=1, max_iter=2, world=2, rank=1, aspect_grouping=False).batch_sampler
    assert not ({i for x in a for i in x} & {i for x in b for i in x})
