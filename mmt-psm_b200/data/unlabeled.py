"""Unlabeled-crop dataset + collator with the colour views produced on the device.

Mirrors reference data/datasets/Pap.py:818-831 (`PapNucleiUnlabelDataset.__getitem__`: transpart1 once, then aug_k x
transpart2 on deep copies) and data/collate_batch.py:58-77 (`BatchCollatorWoLabelK`), with the work split differently:
  * `UnlabeledCrops.__getitem__` does what must stay on the host — decode, `Resize(min, max)` (PIL bilinear, exactly the
    reference's torchvision call) and `RandomHorizontalFlip(0.5)` of transpart1 (data/transforms/build.py:23-26) — and
    DRAWS the parameters of the aug_k views in the reference's order (data/unlabeled_views.draw_view_params);
  * `DeviceViewCollator` uploads each uint8 crop once (3 bytes / pixel instead of aug_k x 12) and renders all views with
    two launches per batch (libmmtpsm_b200 color kernels, bit-exact with the PIL pipeline), zero-padded to
    SIZE_DIVISIBILITY like `to_image_list`.
The collator returns what the trainer's `forward_unlabel` consumes: ([aug_k tensors [B,3,H,W]], indices)."""
import math
import random

import numpy as np
import torch

from .unlabeled_views import colour_views, draw_view_params


def reference_resize_size(w, h, min_size, max_size):
    """Resize.get_size (reference data/transforms/transforms.py:35-58), returns (out_h, out_w)."""
    size = min_size
    if max_size is not None:
        if w == max_size and h == max_size:
            return (h, w)
        lo, hi = float(min(w, h)), float(max(w, h))
        if hi / lo * size > max_size:
            size = int(round(max_size * lo / hi))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        return (int(size * h / w), size)
    return (size, int(size * w / h))


class UnlabeledCrops(object):
    def __init__(self, crops, min_size=800, max_size=1333, aug_k=3, flip_prob=0.5):
        """crops: sequence of uint8 RGB arrays [H,W,3], PIL images, or image file paths."""
        self.crops, self.min_size, self.max_size = crops, min_size, max_size
        self.aug_k, self.flip_prob = aug_k, flip_prob

    def __len__(self):
        return len(self.crops)

    def __getitem__(self, index):
        from PIL import Image
        import torchvision.transforms.functional as tf
        src = self.crops[index]
        img = Image.open(src).convert("RGB") if isinstance(src, str) else (
            src if isinstance(src, Image.Image) else Image.fromarray(np.asarray(src, dtype=np.uint8), "RGB"))
        img = tf.resize(img, reference_resize_size(img.size[0], img.size[1], self.min_size, self.max_size))
        if random.random() < self.flip_prob:
            img = tf.hflip(img)
        arr = np.array(img)
        params = [draw_view_params(arr.shape[0], arr.shape[1]) for _ in range(self.aug_k)]
        return arr, params, index


class DeviceViewCollator(object):
    def __init__(self, device, size_divisible=0, pixel_mean=(102.9801, 115.9465, 122.7717)):
        self.device, self.size_divisible, self.pixel_mean = torch.device(device), size_divisible, pixel_mean

    def __call__(self, batch):
        arrs, params, idx = zip(*batch)
        aug_k = len(params[0])
        H, W = max(a.shape[0] for a in arrs), max(a.shape[1] for a in arrs)
        if self.size_divisible > 0:
            H = int(math.ceil(H / self.size_divisible) * self.size_divisible)
            W = int(math.ceil(W / self.size_divisible) * self.size_divisible)
        same = all(a.shape == arrs[0].shape for a in arrs)
        if same:   # one upload, one launch pair for the whole batch
            dev = torch.from_numpy(np.stack(arrs)).to(self.device, non_blocking=True)
            views = colour_views(dev, [[p[k] for p in params] for k in range(aug_k)], self.pixel_mean)
            if (H, W) == arrs[0].shape[:2]:
                return list(views), idx
            out = [v.new_zeros((len(arrs), 3, H, W)) for v in views]
            for o, v in zip(out, views):
                o[:, :, :v.shape[2], :v.shape[3]].copy_(v)
            return out, idx
        out = [torch.zeros((len(arrs), 3, H, W), dtype=torch.float32, device=self.device) for _ in range(aug_k)]
        for b, (a, p) in enumerate(zip(arrs, params)):   # ragged batch: per image, zero-padded like to_image_list
            views = colour_views(torch.from_numpy(a)[None].to(self.device, non_blocking=True), [[pk] for pk in p],
                                 self.pixel_mean)
            for k in range(aug_k):
                out[k][b, :, :a.shape[0], :a.shape[1]].copy_(views[k][0])
        return out, idx
