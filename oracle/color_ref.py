"""TEST INFRASTRUCTURE — CPU oracle of the unlabeled-view colour pipeline.  Never imported by the product.

Restates, in numpy, what the reference applies to every one of the K + S views of an unlabeled crop
(maskrcnn_benchmark/data/transforms/build.py:27-35 `transpart2`; data/datasets/Pap.py:818-831):

    AdjustBrightness(0.15) -> AdjustContrast(0.15) -> AdjustHue(0.05) -> RandomErasing(0.9) -> ToTensor -> Normalize(BGR255)
    (transforms.py:120-143, 146-206, 84-99)

The three colour ops are `torchvision.transforms.functional.adjust_*` on PIL images, i.e. Pillow's integer / float32
arithmetic: ImageEnhance = Image.blend(degenerate, image, factor) (ImagingBlend: float32 `in1 + alpha * (in2 - in1)`,
clipped, truncated), the contrast pivot = int(mean(L) + 0.5) with L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16, and
the hue shift = uint8 wrap-around add on the H plane of Pillow's RGB <-> HSV conversion (Convert.c rgb2hsv_row /
hsv2rgb: colorsys formulas evaluated in float with double-promoted constants, truncating / rounding to uint8).
Pinning: tests/test_oracle_cpu.py checks `rgb_to_hsv` / `hsv_to_rgb` against Pillow itself over ALL 2^24 colours and
the whole chain against the reference's own transform classes (tests/golden/color_views.npz)."""
import numpy as np

F32, F64 = np.float32, np.float64
PIXEL_MEAN = (102.9801, 115.9465, 122.7717)


def blend_u8(degenerate, image_u8, factor):
    """ImageEnhance.*.enhance(factor) = Image.blend(degenerate, image, factor) on uint8 data."""
    x = image_u8.astype(F32)
    d = np.asarray(degenerate, dtype=F32)
    t = d + F32(factor) * (x - d)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, np.floor(t))).astype(np.uint8)


def adjust_brightness(image_u8, factor):
    return blend_u8(0.0, image_u8, factor)


def luma_u8(image_u8):
    r, g, b = (image_u8[..., i].astype(np.int64) for i in range(3))
    return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16


def adjust_contrast(image_u8, factor):
    mean = int(luma_u8(image_u8).mean() + 0.5)
    return blend_u8(float(mean), image_u8, factor)


def rgb_to_hsv(rgb_u8):
    rgb = rgb_u8.reshape(-1, 3)
    r, g, b = (rgb[:, i].astype(np.int32) for i in range(3))
    maxc, minc = np.maximum(r, np.maximum(g, b)), np.minimum(r, np.minimum(g, b))
    cr = (maxc - minc).astype(F32)
    crs = np.where(cr == 0, F32(1), cr)
    s = (cr / np.where(maxc == 0, 1, maxc).astype(F32)).astype(F32)
    rc = ((maxc - r).astype(F32) / crs).astype(F32)
    gc = ((maxc - g).astype(F32) / crs).astype(F32)
    bc = ((maxc - b).astype(F32) / crs).astype(F32)
    h = np.where(r == maxc, (bc - gc).astype(F32),
                 np.where(g == maxc, (2.0 + rc.astype(F64) - bc.astype(F64)).astype(F32),
                          (4.0 + gc.astype(F64) - rc.astype(F64)).astype(F32)))
    h = np.fmod(h.astype(F64) / 6.0 + 1.0, 1.0).astype(F32)
    uh = np.clip((h.astype(F64) * 255.0).astype(np.int64), 0, 255)
    us = np.clip((s.astype(F64) * 255.0).astype(np.int64), 0, 255)
    gray = minc == maxc
    out = np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], -1).astype(np.uint8)
    return out.reshape(rgb_u8.shape)


def hsv_to_rgb(hsv_u8):
    hsv = hsv_u8.reshape(-1, 3)
    h, s, v = (hsv[:, i] for i in range(3))
    hf = h.astype(F32).astype(F64) * 6.0 / 255.0
    i = np.floor(hf).astype(np.int64)
    f = (hf - i.astype(F32).astype(F64)).astype(F32).astype(F64)
    fs = (s.astype(F32).astype(F64) / 255.0).astype(F32).astype(F64)
    vf = v.astype(F32).astype(F64)
    rnd = lambda x: np.floor(x + 0.5).astype(np.int64)     # arguments are >= 0: C round() == floor(x + 0.5)
    p = np.clip(rnd(vf * (1.0 - fs)), 0, 255)
    q = np.clip(rnd(vf * (1.0 - fs * f)), 0, 255)
    t = np.clip(rnd(vf * (1.0 - fs * (1.0 - f))), 0, 255)
    vi, k = v.astype(np.int64), i % 6
    r = np.choose(k, [vi, q, p, p, t, vi])
    g = np.choose(k, [t, vi, vi, q, p, p])
    b = np.choose(k, [p, p, t, vi, vi, q])
    gray = s == 0
    out = np.stack([np.where(gray, vi, r), np.where(gray, vi, g), np.where(gray, vi, b)], -1).astype(np.uint8)
    return out.reshape(hsv_u8.shape)


def hue_shift_u8(hue_factor):
    """torchvision F_pil.adjust_hue: `np_h += np.uint8(hue_factor * 255)` with wrap-around."""
    return int(np.trunc(hue_factor * 255.0)) % 256


def adjust_hue(image_u8, hue_factor):
    hsv = rgb_to_hsv(image_u8)
    hsv[..., 0] = (hsv[..., 0].astype(np.int64) + hue_shift_u8(hue_factor)).astype(np.uint8)
    return hsv_to_rgb(hsv)


def erase(image_u8, rects):
    """RandomErasing applied: rects = list of (top, left, fill [h, w, 3] float64 in [0, 255)); the uint8 image takes the
    truncated fill (numpy float -> uint8 assignment)."""
    out = image_u8.copy()
    for top, left, fill in rects:
        h, w = fill.shape[:2]
        out[top:top + h, left:left + w, :] = fill.astype(np.uint8)
    return out


def to_tensor_normalize(image_u8, mean=PIXEL_MEAN):
    """ToTensor (/255, CHW) -> Normalize(to_bgr255): image[[2,1,0]] * 255 - mean (std 1).  float32 throughout."""
    x = (image_u8.astype(F32) / F32(255)).transpose(2, 0, 1)
    x = x[[2, 1, 0]] * F32(255)
    return (x - np.asarray(mean, dtype=F32)[:, None, None]) / F32(1.0)


def colour_view(image_u8, brightness, contrast, hue, rects=()):
    x = adjust_brightness(image_u8, brightness)
    x = adjust_contrast(x, contrast)
    x = adjust_hue(x, hue)
    x = erase(x, rects)
    return to_tensor_normalize(x)
