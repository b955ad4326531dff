"""Index tables of the bf16 operand copies ("shadows") of the filters.

Every tensor-core operand the kernels read is a fixed permutation / zero padding of an fp32 master variable.  For each
layout this module computes, once at engine construction, the flat source index (into the parameter arena) of every
destination element (-1 = zero padding); `vl_gather_bf16` then refreshes all of them in one launch after each
optimiser step.  Each builder restates the layout of the stand-alone packing kernel it replaces (include/vlb200.h);
tests/test_gpu_parity.py::test_shadow_gather_equals_the_pack_kernels holds them bit-identical.
"""
import numpy as np


def identity(numel):
    """vl_cast_f32_to_bf16: HWIO filter seen as [taps * cin_g, cout] (data-gradient operand)."""
    return np.arange(numel, dtype=np.int64)


def s2d_filter_kmajor(kh, kw, cin, cout, s, chunk):
    """vl_s2d_pack_filter(transpose=1): K-major filter of the space-to-depth conv1,
    dst[o][(tr*kb+ts)*chunk + (dy*s+dx)*cin + c] = src[s*tr+dy][s*ts+dx][c][o]."""
    kb_h, kb_w = -(-kh // s), -(-kw // s)
    rows = kb_h * kb_w * chunk
    o, row = np.meshgrid(np.arange(cout), np.arange(rows), indexing="ij")
    tap, j = row // chunk, row % chunk
    tr, ts = tap // kb_w, tap % kb_w
    c, dd = j % cin, j // cin
    dy, dx = dd // s, dd % s
    r, q = s * tr + dy, s * ts + dx
    valid = (j < s * s * cin) & (r < kh) & (q < kw)
    src = ((r * kw + q) * cin + c) * cout + o
    return np.where(valid, src, -1).reshape(-1)


def kmajor_padded(taps, cin_g, cout, dst_grp):
    """vl_pack_bf16_t: dst[col][(r // cin_g) * dst_grp + r % cin_g] = src[r][col] with src = [taps * cin_g, cout]."""
    dst_ld = taps * dst_grp
    col, kk = np.meshgrid(np.arange(cout), np.arange(dst_ld), indexing="ij")
    grp, rr = kk // dst_grp, kk % dst_grp
    r = grp * cin_g + rr
    valid = rr < cin_g
    return np.where(valid, r * cout + col, -1).reshape(-1)


def dgrad_d2s(kh, kw, cin_g, cout_g, groups, sh, sw):
    """vl_pack_dgrad_d2s: dst[g*sh*sw*cin_g + (dy*sw+dx)*cin_g + c][(ty*kw2+tx)*kpad + k] =
    src[kh-1+dy-ty][kw-1+dx-tx][c][g*cout_g + k]."""
    kpad = -(-cout_g // 64) * 64
    kh2, kw2 = kh + sh - 1, kw + sw - 1
    ld = kh2 * kw2 * kpad
    rows = groups * sh * sw * cin_g
    row, col = np.meshgrid(np.arange(rows), np.arange(ld), indexing="ij")
    tap, k = col // kpad, col % kpad
    ty, tx = tap // kw2, tap % kw2
    per_g = sh * sw * cin_g
    g, rr = row // per_g, row % per_g
    seg, c = rr // cin_g, rr % cin_g
    dy, dx = seg // sw, seg % sw
    r, q = kh - 1 + dy - ty, kw - 1 + dx - tx
    valid = (k < cout_g) & (r >= 0) & (r < kh) & (q >= 0) & (q < kw)
    src = ((r * kw + q) * cin_g + c) * (groups * cout_g) + g * cout_g + k
    return np.where(valid, src, -1).reshape(-1)


def col_padded(rows, cols, dst_ld):
    """vl_pack_bf16 with one group: dst[r][c] = src[r][c] for c < cols, zero padded to dst_ld columns."""
    r, c = np.meshgrid(np.arange(rows), np.arange(dst_ld), indexing="ij")
    return np.where(c < cols, r * cols + c, -1).reshape(-1)


# ------------------------------------------------------------------------------------------------
# fp32-accuracy mode (fp32_path.py): K-major filters [W_hi | W_hi | W_lo] along the input-channel axis.
# Table entries index the fp32 source; bit 30 selects the lo part (vl_gather_split_bf16).
# ------------------------------------------------------------------------------------------------
LO_FLAG = 1 << 30


def split3_kmajor(taps, cin_g, cout, dst_grp):
    """dst[col][tap * dst_grp + part * cin_g + c] = {hi, hi, lo}[part] of src[tap * cin_g + c][col] (src = HWIO seen as
    [taps * cin_g, cout]); columns beyond 3 * cin_g of a tap are zero padding."""
    assert dst_grp >= 3 * cin_g
    col, kk = np.meshgrid(np.arange(cout), np.arange(taps * dst_grp), indexing="ij")
    tap, rr = kk // dst_grp, kk % dst_grp
    part, c = rr // cin_g, rr % cin_g
    valid = rr < 3 * cin_g
    idx = (tap * cin_g + c) * cout + col
    idx = np.where(part == 2, idx | LO_FLAG, idx)
    return np.where(valid, idx, -1).reshape(-1)


def split3_s2d_kmajor(kh, kw, cin, cout, s, dst_grp):
    """The same for the space-to-depth conv1: dst[o][tap * dst_grp + part * (s*s*cin) + (dy*s+dx)*cin + c] =
    {hi, hi, lo}[part] of src[s*tr+dy][s*ts+dx][c][o] (-1 where the tap leaves the kh x kw filter)."""
    kb_h, kb_w = -(-kh // s), -(-kw // s)
    blk = s * s * cin
    assert dst_grp >= 3 * blk
    o, row = np.meshgrid(np.arange(cout), np.arange(kb_h * kb_w * dst_grp), indexing="ij")
    tap, rr = row // dst_grp, row % dst_grp
    part, j = rr // blk, rr % blk
    tr, ts = tap // kb_w, tap % kb_w
    c, dd = j % cin, j // cin
    dy, dx = dd // s, dd % s
    r, q = s * tr + dy, s * ts + dx
    valid = (rr < 3 * blk) & (r < kh) & (q < kw)
    idx = ((r * kw + q) * cin + c) * cout + o
    idx = np.where(part == 2, idx | LO_FLAG, idx)
    return np.where(valid, idx, -1).reshape(-1)
