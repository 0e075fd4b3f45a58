"""Plug-in registries under the reference's names and string keys (reference modeling/registry.py:5-7)."""
from ..utils.registry import Registry

BACKBONES = Registry()
ROI_BOX_FEATURE_EXTRACTORS = Registry()
RPN_HEADS = Registry()
