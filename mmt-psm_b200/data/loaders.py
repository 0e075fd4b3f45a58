"""Minimal loaders with the iteration contract of the reference's `make_mt_data_loader` (data/build.py:118-230):
the labeled ('source') loader yields `(images, targets, indices)` for exactly `max_iter` iterations, the unlabeled
('no_label') loader yields `([aug_k view batches], indices)`.  The reference's Pap-smear dataset classes (openslide
readers, annotation parsing) are out of scope; these loaders take in-memory crops or the seeded synthetic stream the
benchmark uses, and are what `tools/train_mean_teacher.py --synthetic` drives."""
import numpy as np
import torch

from .samplers import DistributedSampler  # noqa: F401  (re-exported for callers that shard real crop lists)
from .synthetic import PIXEL_MEAN, make_batch
from .unlabeled import DeviceViewCollator, UnlabeledCrops


class SyntheticSourceLoader(object):
    """Labeled batches of the synthetic 2-class crop stream, `max_iter` of them, sharded by rank through the seed."""

    def __init__(self, max_iter, batch=2, size=800, seed=1234, rank=0, pool=2, n_objects=(20, 20), pin=True):
        self.max_iter = max_iter
        self.pool = [make_batch(seed + 17 * rank + 1000 * j, batch, size, 0, n_objects[0], n_objects[1], pin=pin)[:2]
                     for j in range(pool)]

    def __len__(self):
        return self.max_iter

    def __iter__(self):
        for it in range(self.max_iter):
            images, targets = self.pool[it % len(self.pool)]
            yield images, targets, tuple(range(len(targets)))


def synthetic_unlabeled_crops(n, crop=1000, seed=99):
    """uint8 RGB crops (noise + blobs) standing in for the unlabeled Pap-smear patches."""
    g = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        img = g.integers(150, 211, (crop, crop, 3)).astype(np.uint8)
        for _ in range(40):
            cy, cx, r = g.integers(0, crop), g.integers(0, crop), g.integers(10, 80)
            y0, y1, x0, x1 = max(cy - r, 0), min(cy + r, crop), max(cx - r, 0), min(cx + r, crop)
            img[y0:y1, x0:x1] -= np.uint8(g.integers(30, 90))
        out.append(img)
    return out


class UnlabeledViewLoader(object):
    """Endless iterator over `crops` in shuffled order: host-side Resize / flip, device-side colour views."""

    def __init__(self, crops, cfg, device, batch=2, seed=0):
        self.ds = UnlabeledCrops(crops, cfg.INPUT.MIN_SIZE_TRAIN if isinstance(cfg.INPUT.MIN_SIZE_TRAIN, int)
                                 else cfg.INPUT.MIN_SIZE_TRAIN[0], cfg.INPUT.MAX_SIZE_TRAIN,
                                 aug_k=cfg.MT.AUG_K + cfg.MT.AUG_S)
        self.collate = DeviceViewCollator(device, cfg.DATALOADER.SIZE_DIVISIBILITY, tuple(cfg.INPUT.PIXEL_MEAN))
        self.batch, self.gen = batch, torch.Generator().manual_seed(seed)

    def __iter__(self):
        while True:
            order = torch.randperm(len(self.ds), generator=self.gen).tolist()
            for i in range(0, len(order) - self.batch + 1, self.batch):
                yield self.collate([self.ds[j] for j in order[i:i + self.batch]])


def aspect_groups(sizes_wh, bins=(1,)):
    """Group id per image from its height / width ratio (reference data/build.py:80-93, `_quantize` over
    DATALOADER.ASPECT_RATIO_GROUPING's single bin edge 1: portrait-or-square = 1, landscape = 0)."""
    import bisect
    edges = sorted(bins)
    return [bisect.bisect_right(edges, float(h) / float(w)) for w, h in sizes_wh]


class LabeledCropLoader(object):
    """The labeled ('source') loader over real crops: the reference's sampler chain (data/build.py:71-116 —
    Random / Sequential / Distributed sampler -> GroupedBatchSampler by aspect ratio -> IterationBasedBatchSampler for
    exactly `max_iter` batches, resumable at `start_iter`) around `LabeledCrops`, with the collation on the device
    (`DeviceSourceCollator`).  Yields `(ImageList, targets, indices)`; host-side item work can run in `num_workers`
    DataLoader workers (the device collator runs in the consuming process)."""

    def __init__(self, dataset, device, images_per_batch=2, max_iter=1, start_iter=0, shuffle=True, aspect_grouping=True,
                 size_divisible=32, pixel_mean=PIXEL_MEAN, world=1, rank=0, num_workers=0, sizes_wh=None):
        from .labeled import DeviceSourceCollator
        from .samplers import GroupedBatchSampler, IterationBasedBatchSampler
        if world > 1:
            sampler = DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=shuffle)
        elif shuffle:
            sampler = torch.utils.data.RandomSampler(dataset)
        else:
            sampler = torch.utils.data.SequentialSampler(dataset)
        if aspect_grouping:
            if sizes_wh is None:
                sizes_wh = [self._size_of(dataset.crops[i]) for i in range(len(dataset))]
            batches = GroupedBatchSampler(sampler, aspect_groups(sizes_wh), images_per_batch, drop_uneven=False)
        else:
            batches = torch.utils.data.BatchSampler(sampler, images_per_batch, drop_last=False)
        self.batch_sampler = IterationBasedBatchSampler(batches, max_iter, start_iter)
        self.loader = torch.utils.data.DataLoader(dataset, batch_sampler=self.batch_sampler, num_workers=num_workers,
                                                  collate_fn=list)
        self.collate = DeviceSourceCollator(device, size_divisible, tuple(pixel_mean))

    @staticmethod
    def _size_of(src):
        if isinstance(src, str):
            from PIL import Image
            with Image.open(src) as im:
                return im.size
        if hasattr(src, "size") and not hasattr(src, "shape"):
            return src.size                                   # PIL image: (w, h)
        return (src.shape[1], src.shape[0])

    def __len__(self):
        return len(self.batch_sampler)

    def __iter__(self):
        for items in self.loader:
            yield self.collate(items)
