"""Labeled-crop dataset + collator: the reference's `source` item pipeline with the colour work on the device.

Reference: data/transforms/build.py:37-50 (`Resize`, `RandomHorizontalFlip(0.5)`, `AdjustBrightness(0.15)`,
`AdjustContrast(0.15)`, `AdjustHue(0.05)`, `RandomErasing(0.7)`, `ToTensor`, `Normalize(BGR255 - PIXEL_MEAN)`) applied by the
dataset's `__getitem__` to (PIL image, BoxList target), then data/collate_batch.py:8-24 (`BatchCollator`: `to_image_list`
with SIZE_DIVISIBILITY, targets as a tuple).  Split as for the unlabeled crops (data/unlabeled.py): the host does what
must stay there — decode, the PIL bilinear `Resize`, the flip, and the same transforms on the target (BoxList.resize /
transpose carry the polygon masks along) — and draws the colour / erase parameters in the reference's order; the collator
uploads the uint8 crop once and renders the normalised tensor with the colour kernels (bit-exact with the PIL pipeline),
zero-padded to SIZE_DIVISIBILITY.

Annotations are plain arrays (the reference's Pap-smear file readers, data/datasets/Pap.py, are out of scope):
    {"boxes": [n, 4] xyxy, "labels": [n] (1 = cytoplasm, 2 = nuclei), "polygons": per instance a list of flat [x0, y0, ...]}"""
import random

import numpy as np
import torch

from ..structures.bounding_box import BoxList
from ..structures.image_list import ImageList
from ..structures.segmentation_mask import SegmentationMask
from .unlabeled import DeviceViewCollator, reference_resize_size
from .unlabeled_views import draw_view_params


class LabeledCrops(object):
    def __init__(self, crops, annotations, min_size=800, max_size=1333, flip_prob=0.5, erase_prob=0.7):
        """crops: sequence of uint8 RGB arrays [H,W,3], PIL images, or image file paths; annotations: see module doc."""
        assert len(crops) == len(annotations)
        self.crops, self.annotations = crops, annotations
        self.min_size, self.max_size, self.flip_prob, self.erase_prob = min_size, max_size, flip_prob, erase_prob

    def __len__(self):
        return len(self.crops)

    def target(self, index, image_size):
        a = self.annotations[index]
        t = BoxList(torch.as_tensor(np.asarray(a["boxes"], dtype=np.float32)).reshape(-1, 4), image_size, mode="xyxy")
        t.add_field("labels", torch.as_tensor(np.asarray(a["labels"], dtype=np.int64)))
        if "polygons" in a:
            t.add_field("masks", SegmentationMask([[list(map(float, p)) for p in inst] for inst in a["polygons"]], image_size))
        return t

    def __getitem__(self, index):
        from PIL import Image
        import torchvision.transforms.functional as tf
        src = self.crops[index]
        img = Image.open(src).convert("RGB") if isinstance(src, str) else (
            src if isinstance(src, Image.Image) else Image.fromarray(np.asarray(src, dtype=np.uint8), "RGB"))
        target = self.target(index, img.size)
        img = tf.resize(img, reference_resize_size(img.size[0], img.size[1], self.min_size, self.max_size))
        target = target.resize(img.size)
        if random.random() < self.flip_prob:
            img = tf.hflip(img)
            target = target.transpose(0)
        arr = np.array(img)
        params = draw_view_params(arr.shape[0], arr.shape[1], erase_prob=self.erase_prob)
        return arr, target, params, index


class DeviceSourceCollator(object):
    """(ImageList on the device, tuple of targets, indices) — what `MTtrainer.forward_source` consumes."""

    def __init__(self, device, size_divisible=0, pixel_mean=(102.9801, 115.9465, 122.7717)):
        self.views = DeviceViewCollator(device, size_divisible, pixel_mean)

    def __call__(self, batch):
        arrs, targets, params, idx = zip(*batch)
        views, _ = self.views([(a, [p], i) for a, p, i in zip(arrs, params, idx)])
        return ImageList(views[0], [tuple(a.shape[:2]) for a in arrs]), tuple(targets), idx
