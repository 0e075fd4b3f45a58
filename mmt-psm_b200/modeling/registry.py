"""Plug-in registries under the reference's names and string keys (`registry.BACKBONES["R-50-FPN"]`,
`registry.RPN_HEADS["SingleConvRPNHead"]`, `registry.ROI_BOX_FEATURE_EXTRACTORS["FPN2MLPFeatureExtractor"]`):
the lookups `build_backbone`, `build_rpn` and `make_roi_box_feature_extractor` resolve cfg strings through."""
from ..utils.registry import Registry

__all__ = ["BACKBONES", "RPN_HEADS", "ROI_BOX_FEATURE_EXTRACTORS"]

for _table in __all__:
    globals()[_table] = Registry()
del _table
