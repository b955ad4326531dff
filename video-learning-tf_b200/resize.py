"""Read-time image resampling of the reference on the device.

`scipy.misc.imresize(image, shape)` (dataset_.py:238,484,491; serialize.py:425; default interp 'bilinear') is Pillow's
`Image.resize(size, BILINEAR)`: two separable passes over uint8 data with 22-bit fixed-point coefficients.  The host
computes the coefficient tables exactly as Pillow's `precompute_coeffs` / `normalize_coeffs_8bpc` do (double
arithmetic, same operation order); `vl_resize_bilinear_u8` (csrc/resize.cu) applies them.  Bit-exact against PIL
(tests/golden/resize_bilinear_golden.npz).
"""
import math

import numpy as np
import torch

from . import _native as nv

PRECISION_BITS = 32 - 8 - 2


def _triangle(x):
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


def pil_bilinear_coeffs(in_size, out_size):
    """Pillow's coefficient table of one axis: (bounds int32 [out, 2] = first input index and tap count,
    coeffs int32 [out, ksize])."""
    in_size, out_size = int(in_size), int(out_size)
    if in_size < 1 or out_size < 1:
        raise ValueError("resize extents must be positive (%d -> %d)" % (in_size, out_size))
    scale = float(in_size) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    one = float(1 << PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_triangle((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            coeffs[xx, x] = int(-0.5 + v * one) if v < 0 else int(0.5 + v * one)
        bounds[xx] = (xmin, xmax)
    return bounds, coeffs


class DeviceResizer(object):
    """Caches the coefficient tables (device) and scratch buffers of the resizes an Engine is asked for."""

    def __init__(self, device):
        self.device = device
        self._tables = {}
        self._buffers = {}

    def _table(self, in_size, out_size):
        key = (int(in_size), int(out_size))
        t = self._tables.get(key)
        if t is None:
            b, c = pil_bilinear_coeffs(*key)
            t = (torch.from_numpy(b).to(self.device), torch.from_numpy(c).to(self.device), int(c.shape[1]))
            self._tables[key] = t
        return t

    def _buffer(self, tag, shape):
        buf = self._buffers.get(tag)
        if buf is None or buf.numel() < int(np.prod(shape)):
            buf = torch.empty(int(np.prod(shape)), dtype=torch.uint8, device=self.device)
            self._buffers[tag] = buf
        return buf[:int(np.prod(shape))].view(*shape)

    def resize(self, frames, out_h, out_w):
        """uint8 device tensor [n, h, w, c] -> uint8 device tensor [n, out_h, out_w, c] (a cached buffer: consume it
        before the next call)."""
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.is_contiguous()
        n, h, w, c = (int(x) for x in frames.shape)
        if (h, w) == (out_h, out_w):
            return frames
        out = self._buffer("out", (n, out_h, out_w, c))
        bw = cw = bh = ch = None
        kw = kh = 0
        if w != out_w:
            bw, cw, kw = self._table(w, out_w)
        if h != out_h:
            bh, ch, kh = self._table(h, out_h)
        tmp = self._buffer("tmp", (n, h, out_w, c)) if (w != out_w and h != out_h) else None
        nv.call("vl_resize_bilinear_u8", frames, out, tmp, n, h, w, out_h, out_w, c, bw, cw, kw, bh, ch, kh)
        return out
