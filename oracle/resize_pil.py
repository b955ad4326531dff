"""ORACLE (test infrastructure, not product code): CPU restatement of `scipy.misc.imresize(image, shape)` as the
reference calls it at read time (dataset_.py:238,484,491; serialize.py:425: default `interp='bilinear'`).

The arithmetic lives in a third-party dependency that is not vendored under /root/reference: scipy.misc.imresize ->
`PIL.Image.resize(size, resample=BILINEAR)` -> Pillow's `ImagingResample` (src/libImaging/Resample.c, 8 bits per
channel path).  The reference pins neither scipy nor Pillow (dependencies.txt:1-4); this restatement follows the
algorithm of Pillow's Resample.c as published since Pillow 3.x and is PINNED against the Pillow installed in the build
container (12.2.0): `tests/golden/make_golden_resize.py` writes PIL's own outputs to
`tests/golden/resize_bilinear_golden.npz`, `tests/test_host_resize.py` holds this file bit-exact to them.

Algorithm (per axis, horizontal pass first, each pass rounded to uint8):
  scale = in / out; filterscale = max(scale, 1); support = 1.0 * filterscale            (bilinear support 1.0)
  for every output index xx: center = (xx + 0.5) * scale
      xmin = max(int(center - support + 0.5), 0); xmax = min(int(center + support + 0.5), in)
      w[x] = triangle((x + xmin - center + 0.5) / filterscale), normalised to sum 1
      kk[x] = int(w[x] * 2^22 +- 0.5)                                                    (PRECISION_BITS = 32 - 8 - 2)
      out[xx] = clip8((2^21 + sum_x in[xmin + x] * kk[x]) >> 22)
An axis whose size does not change is skipped.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _triangle(x):
    if x < 0.0:
        x = -x
    if x < 1.0:
        return 1.0 - x
    return 0.0


def precompute_coeffs(in_size, out_size):
    """(bounds int32 [out, 2] = (first input index, tap count), coeffs int32 [out, ksize]) of one axis."""
    scale = float(in_size) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, ksize), np.int32)
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
        k = [0.0] * ksize
        ww = 0.0
        for x in range(xmax):
            w = _triangle((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x]
            coeffs[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coeffs


def _resample_axis(img, out_size, axis):
    img = np.moveaxis(img, axis, 0)
    bounds, coeffs = precompute_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt):
            acc += img[xmin + x].astype(np.int64) * int(coeffs[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def imresize_bilinear(image, out_h, out_w):
    """uint8 [..., H, W, C] -> uint8 [..., out_h, out_w, C]; leading axes are independent images."""
    image = np.asarray(image)
    assert image.dtype == np.uint8 and image.ndim >= 3
    h_axis, w_axis = image.ndim - 3, image.ndim - 2
    if out_w != image.shape[w_axis]:
        image = _resample_axis(image, out_w, w_axis)
    if out_h != image.shape[h_axis]:
        image = _resample_axis(image, out_h, h_axis)
    return image
