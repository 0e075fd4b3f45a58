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
