"""GPU suite, labeled-crop pipeline: data/labeled.py's device collation against items produced by the reference's own
transform classes (tests/golden/source_items.npz).  Host-side parity of the same items: tests/test_data_cpu.py."""
import sys

import numpy as np
import pytest

from conftest import ROOT, golden

sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def test_labeled_dataset_and_device_collator(cuda):
    """LabeledCrops + DeviceSourceCollator == the reference's `source` item pipeline + BatchCollator: pixels bit-exact with
    the items the reference's own transform classes produced (tests/golden/source_items.npz), zero-padded ragged batch,
    targets (resized / flipped boxes and polygons) passed through."""
    import random
    from mmt_psm_b200.data.labeled import DeviceSourceCollator
    from test_data_cpu import labeled_dataset
    g = golden("source_items")
    ds = labeled_dataset(g)
    random.seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    images, targets, idx = DeviceSourceCollator(cuda, size_divisible=32)([ds[i] for i in range(len(ds))])
    t = images.tensors.cpu().numpy()
    assert idx == tuple(range(len(ds))) and t.shape[2] % 32 == 0 and t.shape[3] % 32 == 0
    for i in range(len(ds)):
        ref = g["out_image_%d" % i]
        assert tuple(images.image_sizes[i]) == ref.shape[1:]
        assert np.array_equal(t[i, :, :ref.shape[1], :ref.shape[2]], ref), i
        assert not t[i, :, ref.shape[1]:, :].any() and not t[i, :, :, ref.shape[2]:].any()
        np.testing.assert_allclose(targets[i].bbox.numpy(), g["out_boxes_%d" % i], rtol=0, atol=1e-4)

