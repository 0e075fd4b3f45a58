"""Drop-in for the reference's pybind module `maskrcnn_benchmark._C`
(maskrcnn_benchmark/csrc/vision.cpp:7-13): the same five functions with the same Python
signatures and return conventions, served by libmmtpsm_b200.so.

    from mmt_psm_b200 import _C      # instead of: from maskrcnn_benchmark import _C
"""
from .kernels import nms, roi_align_backward, roi_align_forward, roi_pool_backward, roi_pool_forward

__all__ = ["nms", "roi_align_forward", "roi_align_backward", "roi_pool_forward", "roi_pool_backward"]
