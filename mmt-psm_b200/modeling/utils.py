import torch


def cat(tensors, dim=0):
    """torch.cat that skips the copy for a single tensor (reference modeling/utils.py:9-16)."""
    assert isinstance(tensors, (list, tuple))
    return tensors[0] if len(tensors) == 1 else torch.cat(tensors, dim)
