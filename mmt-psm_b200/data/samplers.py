"""Rank sharding of the labeled / unlabeled streams (semantics of reference
data/samplers/distributed.py:26-66: an epoch-seeded permutation, padded by wrap-around to a multiple of the world
size, one contiguous slice per rank) and the iteration-based batch wrapper the trainer's loaders use
(data/samplers/iteration_based_batch_sampler.py)."""
import torch


def shard_indices(n_items, world, rank, epoch=0, shuffle=True):
    """The indices rank `rank` of `world` visits in epoch `epoch` over a dataset of `n_items` items."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    if shuffle:
        g = torch.Generator()
        g.manual_seed(epoch)
        order = torch.randperm(n_items, generator=g).tolist()
    else:
        order = list(range(n_items))
    per_rank = -(-n_items // world)
    order += order[: per_rank * world - n_items]
    return order[rank * per_rank:(rank + 1) * per_rank]


class DistributedSampler(torch.utils.data.Sampler):
    def __init__(self, dataset, num_replicas=None, rank=None, shuffle=True):
        import torch.distributed as dist
        if num_replicas is None:
            num_replicas = dist.get_world_size()
        if rank is None:
            rank = dist.get_rank()
        self.dataset, self.num_replicas, self.rank, self.shuffle = dataset, num_replicas, rank, shuffle
        self.epoch = 0
        self.num_samples = -(-len(dataset) // num_replicas)
        self.total_size = self.num_samples * num_replicas

    def __iter__(self):
        return iter(shard_indices(len(self.dataset), self.num_replicas, self.rank, self.epoch, self.shuffle))

    def __len__(self):
        return self.num_samples

    def set_epoch(self, epoch):
        self.epoch = epoch


class IterationBasedBatchSampler(torch.utils.data.BatchSampler):
    """Re-iterates a batch sampler until `num_iterations` batches were produced (resumable from `start_iter`)."""

    def __init__(self, batch_sampler, num_iterations, start_iter=0):
        self.batch_sampler, self.num_iterations, self.start_iter = batch_sampler, num_iterations, start_iter

    def __iter__(self):
        iteration = self.start_iter
        while iteration <= self.num_iterations:
            if hasattr(self.batch_sampler.sampler, "set_epoch"):
                self.batch_sampler.sampler.set_epoch(iteration)
            for batch in self.batch_sampler:
                iteration += 1
                if iteration > self.num_iterations:
                    break
                yield batch

    def __len__(self):
        return self.num_iterations


class GroupedBatchSampler(torch.utils.data.BatchSampler):
    """Mini-batches whose elements share a group id (aspect-ratio bucket), in an order as close as possible to the
    wrapped sampler's (the batch sampler the reference's train loader builds when DATALOADER.ASPECT_RATIO_GROUPING is
    on, data/build.py + data/samplers/grouped_batch_sampler.py; behaviour pinned by the reference's own
    tests/test_data_samplers.py).  Every group's ids are taken in sampled order and cut into chunks of `batch_size`; the
    chunks of all groups are emitted ordered by where their first element sits in the sampled order; with
    `drop_uneven` the short chunks are dropped."""

    def __init__(self, sampler, group_ids, batch_size, drop_uneven=False):
        if not isinstance(sampler, torch.utils.data.Sampler):
            raise ValueError("sampler should be an instance of torch.utils.data.Sampler, but got sampler=%r" % (sampler,))
        self.sampler, self.group_ids = sampler, list(group_ids)
        self.batch_size, self.drop_uneven = batch_size, drop_uneven
        self._cached = None

    def _batches(self):
        order = list(self.sampler)
        position = {idx: p for p, idx in enumerate(order)}
        per_group = {}
        for idx in order:
            per_group.setdefault(self.group_ids[idx], []).append(idx)
        chunks = [ids[i:i + self.batch_size] for ids in per_group.values() for i in range(0, len(ids), self.batch_size)]
        chunks.sort(key=lambda c: position[c[0]])
        if self.drop_uneven:
            chunks = [c for c in chunks if len(c) == self.batch_size]
        return chunks

    def __iter__(self):
        batches, self._cached = (self._cached if self._cached is not None else self._batches()), None
        return iter(batches)

    def __len__(self):
        if self._cached is None:
            self._cached = self._batches()      # the next __iter__ hands out exactly these
        return len(self._cached)
