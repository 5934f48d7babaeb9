"""TEST INFRASTRUCTURE ONLY — CPU oracle of the image transform in front of the embed path.

Restates, in plain numpy, the transform the reference applies to every image before `ClipMapper`
(`clip_retrieval/clip_inference/reader.py:98-106,158-165` call `preprocess(PIL.Image)`, the object
`load_clip` returns — `mapper.py:36-41`).  For OpenAI/open_clip models that object is torchvision's
    Compose[Resize(n_px, BICUBIC), CenterCrop(n_px), convert("RGB"), ToTensor(), Normalize(mean, std)]
whose arithmetic lives in third-party code that is not under /root/reference:

* Pillow (12.2.0 in this image) `src/libImaging/Resample.c`: `precompute_coeffs`, `bicubic_filter`
  (a = -0.5), `normalize_coeffs_8bpc` (PRECISION_BITS = 22), `ImagingResampleHorizontal_8bpc`,
  `ImagingResampleVertical_8bpc`, `clip8` — horizontal pass first, uint8 intermediate;
* torchvision (0.26) `transforms/functional.py`: `_compute_resized_output_size` (shorter side → n_px,
  longer = int(n_px·long/short)), `center_crop` (`int(round((h - n_px) / 2.0))`, Python banker's
  rounding), `to_tensor` (uint8 → float32, `/ 255`), `normalize` (`(x - mean) / std`, float32).

Pinned (tests/test_preprocess_cpu.py) against Pillow + torchvision themselves running in this container
and against the reference's own fixtures `tests/test_clip_inference/test_images/*.jpg` →
`tests/test_clip_inference/test_tensors/*.pkl` (bit-exact; committed as tests/golden/preprocess_ref.npz
by tests/golden/make_preprocess_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import math

import numpy as np

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)
PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x):
    """Pillow Resample.c `bicubic_filter`, a = -0.5 (Keys), evaluated in the same operation order."""
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size, support_base=2.0):
    """Pillow `precompute_coeffs` + `normalize_coeffs_8bpc` for the full-image box [0, in_size).
    Returns (bounds int32 [out,2] = (xmin, count), coeffs int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = support_base * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_axis(img, out_size, axis):
    """One Pillow 8bpc pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 [H,W,C] image."""
    in_size = img.shape[axis]
    bounds, kk = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        k = kk[xx, :cnt].astype(np.int64)
        acc = np.tensordot(k, src[xmin:xmin + cnt], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resize_bicubic(img, out_h, out_w):
    """Pillow `ImagingResample`: horizontal pass (skipped when the width is unchanged), then vertical."""
    if img.shape[1] != out_w:
        img = resample_axis(img, out_w, 1)
    if img.shape[0] != out_h:
        img = resample_axis(img, out_h, 0)
    return img


def resized_shape(h, w, n_px):
    """torchvision `_compute_resized_output_size` for an int size: shorter side -> n_px."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = n_px, int(n_px * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)   # (new_h, new_w)


def crop_origin(new_h, new_w, n_px):
    """torchvision `center_crop` (Python round = half-to-even)."""
    return int(round((new_h - n_px) / 2.0)), int(round((new_w - n_px) / 2.0))


def preprocess(img_u8, n_px=224, mean=OPENAI_MEAN, std=OPENAI_STD):
    """uint8 RGB [H,W,3] -> float32 [3,n_px,n_px], bit-for-bit what the reference's transform yields."""
    img_u8 = np.ascontiguousarray(img_u8)
    assert img_u8.dtype == np.uint8 and img_u8.ndim == 3 and img_u8.shape[2] == 3
    h, w = img_u8.shape[:2]
    new_h, new_w = resized_shape(h, w, n_px)
    if (new_h, new_w) != (h, w):                       # torchvision short-circuits an already-sized image
        img_u8 = resize_bicubic(img_u8, new_h, new_w)
    top, left = crop_origin(new_h, new_w, n_px)
    crop = img_u8[top:top + n_px, left:left + n_px]
    x = crop.astype(np.float32) / np.float32(255)
    x = (x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def preprocess_batch(images, n_px=224, mean=OPENAI_MEAN, std=OPENAI_STD):
    return np.stack([preprocess(im, n_px, mean, std) for im in images])
