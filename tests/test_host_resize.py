"""Read-time resize (scipy.misc.imresize -> PIL BILINEAR; dataset_.py:238,484,491, serialize.py:425): the oracle
restatement and the product's coefficient tables against PIL's own outputs (tests/golden/resize_bilinear_golden.npz,
written by tests/golden/make_golden_resize.py with the Pillow of the build container)."""
import importlib
import os

import numpy as np
import pytest

from oracle import resize_pil as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_bilinear_golden.npz")


def _cases():
    g = np.load(GOLD)
    i = 0
    while "in_%d" % i in g:
        yield i, g["in_%d" % i], g["out_%d" % i]
        i += 1


def test_oracle_resize_bit_exact_against_pil_golden():
    n = 0
    for i, src, ref in _cases():
        got = R.imresize_bilinear(src, ref.shape[0], ref.shape[1])
        assert got.dtype == np.uint8 and np.array_equal(got, ref), "case %d" % i
        n += 1
    assert n >= 6


def test_oracle_resize_batched_and_identity():
    _, src, ref = next(_cases())
    batch = np.stack([src, src[::-1].copy()])
    got = R.imresize_bilinear(batch, ref.shape[0], ref.shape[1])
    assert np.array_equal(got[0], ref)
    assert np.array_equal(got[1], R.imresize_bilinear(src[::-1].copy(), ref.shape[0], ref.shape[1]))
    assert R.imresize_bilinear(src, src.shape[0], src.shape[1]) is not None
    assert np.array_equal(R.imresize_bilinear(src, src.shape[0], src.shape[1]), src)


def test_product_coefficient_tables_equal_the_oracle():
    """video-learning-tf_b200/resize.py computes the tables the CUDA kernel consumes; same integers as the oracle's."""
    P = importlib.import_module("video-learning-tf_b200.resize")
    for (a, b) in ((53, 31), (20, 47), (160, 227), (320, 227), (30, 7), (227, 227), (1, 5), (5, 1)):
        b0, c0 = R.precompute_coeffs(a, b)
        b1, c1 = P.pil_bilinear_coeffs(a, b)
        assert np.array_equal(b0, b1) and np.array_equal(c0, c1), (a, b)
        # every row of coefficients sums to 2^22 up to the rounding of its taps, and stays inside the input
        assert np.all(np.abs(c1.sum(axis=1) - (1 << 22)) <= c1.shape[1])
        assert np.all(b1[:, 0] >= 0) and np.all(b1[:, 0] + b1[:, 1] <= a)


def test_golden_records_its_pillow_version():
    g = np.load(GOLD)
    assert str(g["pillow_version"]).count(".") >= 1


def test_oracle_resize_against_live_pil_on_random_extents():
    """Where Pillow is importable (it is in the build container), the restatement is also checked against PIL directly on
    random extents, including 1-pixel axes, strong down-scaling and up-scaling."""
    Image = pytest.importorskip("PIL.Image")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 40), st.integers(1, 40), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
    def check(h, w, oh, ow, seed):
        a = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(a).resize((ow, oh), resample=Image.BILINEAR))
        got = R.imresize_bilinear(a, oh, ow)
        assert np.array_equal(got, ref), (h, w, oh, ow)

    check()
