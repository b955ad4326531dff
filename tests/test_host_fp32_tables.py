"""CPU test of the operand tables of the fp32-accuracy mode (shadow_table.split3_*): the K-major [W_hi | W_hi | W_lo]
convolution filters equal a direct construction, for a grouped 5x5 filter and for the space-to-depth conv1."""
import numpy as np
import torch

import vlb200  # noqa: F401
from vlb200 import shadow_table as ST


def _hi_lo(x):
    hi = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    return hi, x - hi


def _apply(table, src):
    hi, lo = _hi_lo(src.reshape(-1))
    idx = np.where(table >= 0, table & ~ST.LO_FLAG, 0)
    val = np.where(table & ST.LO_FLAG, lo[idx], hi[idx])
    return np.where(table < 0, 0.0, val).astype(np.float32)


def test_split3_kmajor_equals_direct_construction():
    taps, cg, cout, grp = 25, 48, 256, 192
    w = np.random.default_rng(0).standard_normal((taps, cg, cout)).astype(np.float32)
    got = _apply(ST.split3_kmajor(taps, cg, cout, grp), w).reshape(cout, taps, grp)
    hi, lo = _hi_lo(w)
    exp = np.zeros((cout, taps, grp), np.float32)
    exp[:, :, 0:cg] = hi.transpose(2, 0, 1)
    exp[:, :, cg:2 * cg] = hi.transpose(2, 0, 1)
    exp[:, :, 2 * cg:3 * cg] = lo.transpose(2, 0, 1)
    assert np.array_equal(got, exp)
    # hi + lo restores the weight to 2^-16 relative
    assert np.abs(hi + lo.astype(np.float32) - w).max() <= 2.0 ** -16 * np.abs(w).max()


def test_split3_s2d_kmajor_equals_direct_construction():
    kh = kw = 11
    cin, cout, s, grp = 3, 96, 4, 192
    w = np.random.default_rng(1).standard_normal((kh, kw, cin, cout)).astype(np.float32)
    got = _apply(ST.split3_s2d_kmajor(kh, kw, cin, cout, s, grp), w).reshape(cout, 9, grp)
    hi, lo = _hi_lo(w)
    exp = np.zeros((cout, 9, grp), np.float32)
    for tr in range(3):
        for ts in range(3):
            for dy in range(s):
                for dx in range(s):
                    r, q = s * tr + dy, s * ts + dx
                    if r >= kh or q >= kw:
                        continue
                    j = (dy * s + dx) * cin
                    for part, srcw in enumerate((hi, hi, lo)):
                        exp[:, tr * 3 + ts, part * 48 + j:part * 48 + j + cin] = srcw[r, q].T
    assert np.array_equal(got, exp)
