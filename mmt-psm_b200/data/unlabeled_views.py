"""Colour views of an unlabeled crop on the device.

The reference builds the K + S views of every unlabeled image on the host, one PIL pass per op and view
(data/datasets/Pap.py:818-831: deepcopy + `transpart2` of data/transforms/build.py:27-35 = AdjustBrightness(0.15),
AdjustContrast(0.15), AdjustHue(0.05), RandomErasing(0.9), ToTensor, Normalize(BGR255 - PIXEL_MEAN); transforms.py:84-206).
Here the resized uint8 crop is uploaded once and `colour_views` produces all views with two kernel launches per view
batch (libmmtpsm_b200 `mmt_color_luma_sums_u8`, `mmt_color_views_u8`), bit-exact with Pillow's arithmetic.

`draw_view_params` draws the random parameters in the reference's order from the same global streams it uses
(python `random` for the three factors and the erase count, `numpy.random` for the rectangles and their fill), so a
seeded run reproduces the reference's views exactly."""
import random

import numpy as np
import torch

from .. import kernels as K


def draw_view_params(height, width, brightness=0.15, contrast=0.15, hue=0.05, erase_prob=0.9,
                     s_l=0.001, s_h=0.004, r_1=0.2, r_2=1 / 0.2, v_l=0, v_h=255):
    """One view's parameters: dict(brightness, contrast, hue, rects=[(top, left, fill[h, w, 3] float64)])."""
    p = {"brightness": random.uniform(1 - brightness, 1 + brightness),      # transforms.py:129-135
         "contrast": random.uniform(1 - contrast, 1 + contrast),            # :136-143
         "hue": random.uniform(-hue, hue), "rects": []}                     # :120-127
    for _ in range(random.randint(0, 10)):                                   # RandomErasing.__call__, :197-206
        if np.random.rand() > erase_prob:
            continue
        while True:
            s = np.random.uniform(s_l, s_h) * height * width
            r = np.random.uniform(r_1, r_2)
            w, h = int(np.sqrt(s / r)), int(np.sqrt(s * r))
            left, top = np.random.randint(0, width), np.random.randint(0, height)
            if left + w <= width and top + h <= height:
                break
        p["rects"].append((top, left, np.random.uniform(v_l, v_h, (h, w, 3))))
    return p


def colour_views(images_u8, params, pixel_mean=(102.9801, 115.9465, 122.7717)):
    """images_u8: CUDA uint8 [B, H, W, 3] (RGB, after the reference's Resize / flip); params: list over views of lists
    over images of `draw_view_params` dicts.  Returns a list over views of fp32 [B, 3, H, W] tensors = what the
    reference's collator hands to the trainer (BGR * 255 - PIXEL_MEAN)."""
    return K.color_views(images_u8, params, pixel_mean)
