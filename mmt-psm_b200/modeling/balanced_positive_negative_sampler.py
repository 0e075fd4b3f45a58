"""Fixed-size positive/negative sampling (reference modeling/balanced_positive_negative_sampler.py:16-71).

`rng_device = None` (default): CUDA inputs are sampled by ONE kernel launch for the whole batch
(`mmt_balanced_sample_f32`: counter-based random keys, no count read-back) — the reference's distribution, a
different random stream.  `rng_device = "cpu"` (parity tests): the reference's own `torch.randperm(count)[:n]`
sequence is reproduced on the host, so that the CPU reference and the CUDA path consume the same random stream."""
import torch


class BalancedPositiveNegativeSampler(object):
    rng_device = None  # None = the labels' device

    _instances = 0
    _offset_cache = {}

    def __init__(self, batch_size_per_image, positive_fraction):
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        # the device-side sampler's random stream: (torch.initial_seed(), this sampler's ordinal, its own call count) —
        # per instance, so that the draws do not depend on how the passes of an iteration interleave (the trainer runs
        # the teacher pass on a second host thread)
        BalancedPositiveNegativeSampler._instances += 1
        self._ordinal = BalancedPositiveNegativeSampler._instances
        self._calls = 0

    def _sample_on_device(self, matched_idxs):
        """One launch for the whole batch (libmmtpsm_b200 `mmt_balanced_sample_f32`): per image a uniformly random
        min(#pos, quota) positives and min(#neg, batch - #pos_taken) negatives — the distribution of the reference's
        randperm(...)[:n] draws (balanced_positive_negative_sampler.py:41-54) without learning the counts on the host.
        Counter-based keys: reproducible under torch.manual_seed, a different stream than torch.randperm."""
        from .. import kernels as K
        sizes = tuple(int(m.numel()) for m in matched_idxs)
        dev = matched_idxs[0].device
        key = (sizes, str(dev))
        offs = self._offset_cache.get(key)
        if offs is None:
            if len(self._offset_cache) > 32:
                self._offset_cache.clear()
            acc = [0]
            for n in sizes:
                acc.append(acc[-1] + n)
            offs = torch.tensor(acc, dtype=torch.int32, device=dev)
            BalancedPositiveNegativeSampler._offset_cache[key] = offs
        labels = matched_idxs[0] if len(matched_idxs) == 1 else torch.cat(list(matched_idxs))
        self._calls += 1
        seed = (torch.initial_seed() * 0x9E3779B1 + self._calls * 0x85EBCA77 + self._ordinal * 0xC2B2AE3D27D4EB4F) & 0xFFFFFFFFFFFFFFFF
        pos, neg = K.balanced_sample(labels.to(torch.float32), offs, len(sizes), self.batch_size_per_image,
                                     int(self.batch_size_per_image * self.positive_fraction), seed)
        return list(pos.split(sizes)), list(neg.split(sizes))

    def __call__(self, matched_idxs):
        pos_idx, neg_idx = [], []
        if self.rng_device is None and matched_idxs and matched_idxs[0].is_cuda:
            return self._sample_on_device(matched_idxs)
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
