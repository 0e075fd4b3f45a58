from .bounding_box import BoxList, FLIP_LEFT_RIGHT, FLIP_TOP_BOTTOM  # noqa: F401
from .image_list import ImageList, to_image_list  # noqa: F401
from .segmentation_mask import SegmentationMask, Polygons  # noqa: F401
