"""Fixed-size positive/negative sampling (reference modeling/balanced_positive_negative_sampler.py:16-71).
`rng_device`: where `torch.randperm` draws (the reference draws on the tensors' device; parity tests set "cpu" so
that the CPU reference and the CUDA path consume the same random stream)."""
import torch


class BalancedPositiveNegativeSampler(object):
    rng_device = None  # None = the labels' device

    def __init__(self, batch_size_per_image, positive_fraction):
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction

    def __call__(self, matched_idxs):
        pos_idx, neg_idx = [], []
        for m in matched_idxs:
            positive = torch.nonzero(m >= 1).squeeze(1)
            negative = torch.nonzero(m == 0).squeeze(1)
            num_pos = min(positive.numel(), int(self.batch_size_per_image * self.positive_fraction))
            num_neg = min(negative.numel(), self.batch_size_per_image - num_pos)
            dev = positive.device if self.rng_device is None else self.rng_device
            perm1 = torch.randperm(positive.numel(), device=dev)[:num_pos].to(positive.device)
            perm2 = torch.randperm(negative.numel(), device=dev)[:num_neg].to(positive.device)
            pos_mask = torch.zeros_like(m, dtype=torch.bool)
            neg_mask = torch.zeros_like(m, dtype=torch.bool)
            pos_mask[positive[perm1]] = True
            neg_mask[negative[perm2]] = True
            pos_idx.append(pos_mask)
            neg_idx.append(neg_mask)
        return pos_idx, neg_idx
