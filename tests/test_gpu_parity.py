"""GPU parity: the CUDA path (through the C-ABI, via the Engine / launch wrappers) against the CPU oracle.

Tolerances (north_star): integer / index outputs bit-exact; bf16 compute path <= 2e-2 relative to the
tensor's max-abs; fp32-only kernels <= 1e-3 (stated per test)."""
import numpy as np
import pytest
import torch

from oracle import lrcn_numpy as O

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2


def rel(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)


def rel_l2(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.sqrt(((got - ref) ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-30)


def bf16_round(a):
    return torch.from_numpy(np.asarray(a, np.float32)).to(torch.bfloat16).float().numpy()


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.fixture(scope="module")
def vl():
    import vlb200
    from vlb200 import _native, engine, kernels
    return dict(nv=_native, K=kernels, E=engine)


# ----------------------------------------------------------------------------------------------------------
# kernels one by one
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n,h,cin,cout,k,groups", [
    ("conv3", 3, 13, 256, 384, 3, 1), ("conv4", 2, 13, 384, 384, 3, 2), ("conv5", 2, 13, 384, 256, 3, 2),
    ("conv2", 2, 28, 96, 256, 5, 2)])
def test_conv_fwd_dgrad_wgrad_vs_oracle(vl, name, n, h, cin, cout, k, groups):
    K = vl["K"]
    rng = np.random.default_rng(10)
    x = bf16_round(rng.standard_normal((n, h, h, cin)))
    w = bf16_round(rng.standard_normal((k, k, cin // groups, cout)) * 0.05)
    b = rng.standard_normal(cout).astype(np.float32)
    dy = bf16_round(rng.standard_normal((n, h, h, cout)))
    y_ref = O.relu(O.conv2d_same(x, w, b, 1, groups))
    dx_ref, dw_ref, db_ref = O.conv2d_same_backward(x, w, dy, 1, groups)
    spec = K.ConvSpec(h, h, cin, cout, k, k, 1, groups)
    xd, dyd = dev(x, torch.bfloat16), dev(dy, torch.bfloat16)
    packed = K.pack_conv_weight_host(spec, dev(w))
    dev_packed = torch.empty_like(packed)  # the device packing kernel must agree with the torch restatement
    vl["nv"].call("vl_pack_bf16_t", dev(w.reshape(-1, cout)), k * k * (cin // groups), cout, dev_packed, spec.k_packed,
                  cin // groups, spec.cchunks * 64)
    assert torch.equal(packed, dev_packed)
    out = torch.empty(n, h, h, cout, dtype=torch.bfloat16, device="cuda")
    K.conv_fwd(spec, xd, packed, dev(b), out, relu=True)
    assert rel(out.float().cpu().numpy(), y_ref) < BF16_TOL
    dx = torch.empty(n, h, h, cin, dtype=torch.bfloat16, device="cuda")
    w2d = dev(w, torch.bfloat16).reshape(-1, cout).contiguous()
    K.conv_dgrad(spec, dyd, w2d, dx)
    assert rel(dx.float().cpu().numpy(), dx_ref) < BF16_TOL
    dw = torch.zeros(k * k * (cin // groups), cout, dtype=torch.float32, device="cuda")
    K.conv_wgrad(spec, xd, dyd, dw, split_k=3)
    assert rel(dw.cpu().numpy().reshape(dw_ref.shape), dw_ref) < 1e-3  # fp32 accumulate of bf16 products
    for split in (1, 5):  # operands swapped (output channels on M, im2col^T on N): same gradient
        dwt = torch.zeros_like(dw)
        K.conv_wgrad_t(spec, xd, dyd, dwt, split_k=split)
        assert rel(dwt.cpu().numpy().reshape(dw_ref.shape), dw_ref) < 1e-3
    db = torch.zeros(cout, dtype=torch.float32, device="cuda")
    vl["nv"].call("vl_colsum", dyd, db, n * h * h, cout, cout)
    assert rel(db.cpu().numpy(), db_ref) < 1e-3


@pytest.mark.parametrize("name,n,h,cin,cout,k,groups", [
    ("conv2", 3, 28, 96, 256, 5, 2), ("conv5", 2, 13, 384, 256, 3, 2), ("conv3", 2, 13, 256, 384, 3, 1),
    ("odd", 2, 9, 64, 32, 3, 1)])
def test_tap_shifted_conv_fwd_dgrad_vs_oracle(vl, name, n, h, cin, cout, k, groups):
    """csrc/conv_flat.cu: input tile staged once, filter taps as shifted shared-memory descriptors, channels on M."""
    nv, K = vl["nv"], vl["K"]
    rng = np.random.default_rng(19)
    x = bf16_round(rng.standard_normal((n, h, h, cin)))
    w = bf16_round(rng.standard_normal((k, k, cin // groups, cout)) * 0.05)
    b = rng.standard_normal(cout).astype(np.float32)
    dy = bf16_round(rng.standard_normal((n, h, h, cout)))
    y_ref = O.relu(O.conv2d_same(x, w, b, 1, groups))
    dx_ref, _, _ = O.conv2d_same_backward(x, w, dy, 1, groups)
    spec = K.ConvSpec(h, h, cin, cout, k, k, 1, groups)
    out = torch.full((n, h, h, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    K.conv_fwd_flat(spec, dev(x, torch.bfloat16), K.pack_conv_weight_host(spec, dev(w)), dev(b), out, relu=True)
    assert rel(out.float().cpu().numpy(), y_ref) < BF16_TOL
    kpad = -(-spec.cout_g // 64) * 64
    wd = torch.empty(cin, spec.taps * kpad, dtype=torch.bfloat16, device="cuda")
    nv.call("vl_pack_dgrad_kmajor", dev(w.reshape(-1, cout)), wd, spec.taps, spec.cin_g, spec.cout_g, groups)
    dx = torch.full((n, h, h, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    K.conv_dgrad_flat(spec, dev(dy, torch.bfloat16), wd, dx)
    assert rel(dx.float().cpu().numpy(), dx_ref) < BF16_TOL


@pytest.mark.parametrize("name,n,h,cin,cout,k,groups,sh,sw", [
    ("conv2 2x2", 3, 28, 96, 256, 5, 2, 2, 2), ("conv2 1x2", 2, 28, 96, 256, 5, 2, 1, 2),
    ("odd extent", 2, 9, 32, 64, 3, 1, 2, 2), ("conv5 2x1", 2, 13, 384, 256, 3, 2, 2, 1)])
def test_depth_to_space_dgrad_vs_oracle(vl, name, n, h, cin, cout, k, groups, sh, sw):
    """Data gradient of a stride-1 convolution as a stride-(sh,sw) forward convolution over dy with a depth-to-space
    epilogue (vl_pack_dgrad_d2s + vl_gemm d2s_*): same result as conv2d_backprop_input, also where h % sh != 0."""
    nv, K = vl["nv"], vl["K"]
    rng = np.random.default_rng(23)
    x = bf16_round(rng.standard_normal((n, h, h, cin)))
    w = bf16_round(rng.standard_normal((k, k, cin // groups, cout)) * 0.05)
    dy = bf16_round(rng.standard_normal((n, h, h, cout)))
    dx_ref, _, _ = O.conv2d_same_backward(x, w, dy, 1, groups)
    spec = K.ConvSpec(h, h, cin, cout, k, k, 1, groups)
    rows, cols = K.d2s_filter_shape(spec, sh, sw)
    wd = torch.empty(rows, cols, dtype=torch.bfloat16, device="cuda")
    nv.call("vl_pack_dgrad_d2s", dev(w), wd, k, k, spec.cin_g, spec.cout_g, groups, sh, sw)
    # the packed operand against its definition (include/vlb200.h)
    kpad = -(-spec.cout_g // 64) * 64
    ref = np.zeros((groups, sh, sw, spec.cin_g, k + sh - 1, k + sw - 1, kpad), np.float32)
    for g in range(groups):
        for dyy in range(sh):
            for dxx in range(sw):
                for ty in range(k + sh - 1):
                    for tx in range(k + sw - 1):
                        r, q = k - 1 + dyy - ty, k - 1 + dxx - tx
                        if 0 <= r < k and 0 <= q < k:
                            ref[g, dyy, dxx, :, ty, tx, :spec.cout_g] = w[r, q, :, g * spec.cout_g:(g + 1) * spec.cout_g]
    assert np.array_equal(wd.float().cpu().numpy(), ref.reshape(rows, cols))
    dx = torch.full((n, h, h, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    K.conv_dgrad_d2s(spec, dev(dy, torch.bfloat16), wd, dx, sh=sh, sw=sw)
    got = dx.float().cpu().numpy()
    assert np.isfinite(got).all()  # every output element is written exactly once
    assert rel(got, dx_ref) < BF16_TOL


def test_conv1_space_to_depth_path_vs_oracle(vl):
    """conv1 (11x11 stride 4 SAME, alexnet.py:60-77) as space-to-depth + 3x3 VALID im2col-TMA convolution."""
    nv, K, E = vl["nv"], vl["K"], vl["E"]
    rng = np.random.default_rng(12)
    n = 2
    frames_u8 = rng.integers(0, 256, size=(n, 227, 227, 3), dtype=np.uint8)
    mean = np.array([99.197148, 105.293620, 109.503945], np.float32)
    x = frames_u8.astype(np.float32) - mean
    w = bf16_round(rng.standard_normal((11, 11, 3, 96)) * 0.05)
    b = np.full(96, 0.1, np.float32)
    sp = E.encoder_specs(227, 227)
    s1s = sp["conv1_s2d"]
    assert (s1s.h, s1s.w, s1s.cin, s1s.p, s1s.q, s1s.taps) == (59, 59, 48, 57, 57, 9)
    # staging kernel: bit-exact against a numpy restatement (integer indexing + one bf16 rounding)
    xs = torch.empty(n, 59, 59, 48, dtype=torch.bfloat16, device="cuda")
    nv.call("vl_frames_s2d", dev(frames_u8), 1, dev(mean), xs, n, 227, 227, 4, 4, 4, 59, 59)
    pad = np.zeros((n, 236, 236, 3), np.float32)
    pad[:, 4:231, 4:231] = x
    ref = pad.reshape(n, 59, 4, 59, 4, 3).transpose(0, 1, 3, 2, 4, 5).reshape(n, 59, 59, 48)
    assert np.array_equal(xs.float().cpu().numpy(), bf16_round(ref))
    xs2 = torch.empty_like(xs)
    nv.call("vl_frames_s2d", dev(x), 0, None, xs2, n, 227, 227, 4, 4, 4, 59, 59)
    assert torch.equal(xs, xs2)  # fp32 feed (feeder.py:97-100) == uint8 + mean feed
    # filter packing + forward
    wp = torch.empty(96, s1s.k_packed, dtype=torch.bfloat16, device="cuda")  # K-major
    nv.call("vl_s2d_pack_filter", dev(w), wp, 11, 11, 3, 96, 4, 64, 1)
    out = torch.empty(n, 57, 57, 96, dtype=torch.bfloat16, device="cuda")
    K.conv_fwd(s1s, xs, wp, dev(b), out, relu=True)
    y_ref = O.relu(O.conv2d_same(bf16_round(x), w, b, 4, 1))
    assert rel(out.float().cpu().numpy(), y_ref) < BF16_TOL
    out2 = torch.full_like(out, float("nan"))
    K.conv_fwd_flat(s1s, xs, wp, dev(b), out2, relu=True)  # tap-shifted kernel: same result up to summation order
    assert rel(out2.float().cpu().numpy(), y_ref) < BF16_TOL
    # filter gradient through the space-to-depth form, scattered back to HWIO
    dy = bf16_round(rng.standard_normal((n, 57, 57, 96)))
    _, dw_ref, _ = O.conv2d_same_backward(bf16_round(x), w, dy, 4, 1, need_dx=False)
    dws = torch.zeros(9 * 48, 96, device="cuda")
    K.conv_wgrad(s1s, xs, dev(dy, torch.bfloat16), dws, split_k=4)
    dw = torch.empty(11, 11, 3, 96, device="cuda")
    nv.call("vl_s2d_unpack_grad", dws, dw, 11, 11, 3, 96, 4)
    assert rel(dw.cpu().numpy(), dw_ref) < 1e-3
    dws2 = torch.zeros_like(dws)
    K.conv_wgrad_t(s1s, xs, dev(dy, torch.bfloat16), dws2)
    assert rel(dws2.cpu().numpy(), dws.cpu().numpy()) < 1e-3
    # row-shift form: one k-block per output row, the three taps of a filter row contract against ONE staged input row
    # through an N-major descriptor with overlapping atoms; split-K over rows (also an uneven split)
    for split in (0, 5):
        dws3 = torch.zeros_like(dws)
        K.conv_wgrad_t(s1s, xs, dev(dy, torch.bfloat16), dws3, split_k=split, row_shift=True)
        assert rel(dws3.cpu().numpy(), dws.cpu().numpy()) < 1e-3
    dw3 = torch.empty(11, 11, 3, 96, device="cuda")
    nv.call("vl_s2d_unpack_grad", dws3, dw3, 11, 11, 3, 96, 4)
    assert rel(dw3.cpu().numpy(), dw_ref) < 1e-3


def test_staging_with_crop_and_mirror_bit_exact(vl):
    """vl_frames_s2d_crop == numpy crop (dataset_.py:444-461) + mirror (:498-500) + mean (:494-495) + space-to-depth."""
    nv = vl["nv"]
    rng = np.random.default_rng(31)
    n, hr, wr = 5, 240, 320
    raw = rng.integers(0, 256, size=(n, hr, wr, 3), dtype=np.uint8)
    crops = np.stack([rng.integers(0, hr - 227, n), rng.integers(0, wr - 227, n), rng.integers(0, 2, n)], 1).astype(np.int32)
    crops[0] = (0, 0, 0)
    crops[1] = (hr - 227, wr - 227, 1)
    mean = np.array([99.197148, 105.293620, 109.503945], np.float32)
    xs = torch.empty(n, 59, 59, 48, dtype=torch.bfloat16, device="cuda")
    nv.call("vl_frames_s2d_crop", dev(raw), 1, dev(mean), xs, n, hr, wr, dev(crops), 227, 227, 4, 4, 4, 59, 59)
    ref = np.zeros((n, 59, 59, 48), np.float32)
    for i, (y0, x0, mir) in enumerate(crops):
        img = raw[i, y0:y0 + 227, x0:x0 + 227, :].astype(np.float32)
        if mir:
            img = img[:, ::-1, :]
        pad = np.zeros((236, 236, 3), np.float32)
        pad[4:231, 4:231] = img - mean
        ref[i] = pad.reshape(59, 4, 59, 4, 3).transpose(0, 2, 1, 3, 4).reshape(59, 59, 48)
    assert np.array_equal(xs.float().cpu().numpy(), bf16_round(ref))


def test_resize_bilinear_bit_exact_against_pil_golden(vl):
    """vl_resize_bilinear_u8 (csrc/resize.cu) == PIL.Image.resize(BILINEAR) on the golden vectors, also batched."""
    import importlib
    import os
    P = importlib.import_module("video-learning-tf_b200.resize")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_bilinear_golden.npz"))
    rz = P.DeviceResizer("cuda")
    i = 0
    while "in_%d" % i in g:
        src, ref = g["in_%d" % i], g["out_%d" % i]
        batch = np.stack([src, src[::-1].copy(), src])
        out = rz.resize(dev(batch), ref.shape[0], ref.shape[1]).cpu().numpy()
        assert out.dtype == np.uint8 and np.array_equal(out[0], ref) and np.array_equal(out[2], ref), i
        assert np.array_equal(out[1], O_resize(src[::-1].copy(), ref.shape[0], ref.shape[1])), i
        i += 1
    assert i >= 6
    # one axis only (the other pass is skipped), and the identity
    src = g["in_0"]
    for (oh, ow) in ((src.shape[0], 29), (19, src.shape[1]), src.shape[:2]):
        out = rz.resize(dev(src[None]), oh, ow).cpu().numpy()[0]
        assert np.array_equal(out, O_resize(src, oh, ow)), (oh, ow)


def O_resize(img, oh, ow):
    from oracle import resize_pil as R
    return R.imresize_bilinear(img, oh, ow)


@pytest.mark.parametrize("c", [96, 256])
def test_lrn_fwd_bwd_vs_oracle(vl, c):
    nv = vl["nv"]
    rng = np.random.default_rng(12)
    x = bf16_round(np.maximum(rng.standard_normal((2, 5, 7, c)) * 40, 0))
    dy = bf16_round(rng.standard_normal((2, 5, 7, c)))
    y_ref = O.lrn(x)
    dx_ref = O.lrn_backward(x, dy) * (x > 0)
    xd, dyd = dev(x, torch.bfloat16), dev(dy, torch.bfloat16)
    y = torch.empty_like(xd)
    nv.call("vl_lrn_fwd", xd, y, 2 * 5 * 7, c, 2, 2e-05, 0.75, 1.0)
    assert rel(y.float().cpu().numpy(), y_ref) < 1e-2  # bf16 output rounding only
    dx = torch.empty_like(xd)
    nv.call("vl_lrn_bwd", xd, dyd, dx, 2 * 5 * 7, c, 2, 2e-05, 0.75, 1.0, 1)
    assert rel(dx.float().cpu().numpy(), dx_ref) < 1e-2


@pytest.mark.parametrize("h,c", [(57, 96), (28, 256), (13, 256)])
def test_maxpool_fwd_bwd_bit_exact(vl, h, c):
    nv = vl["nv"]
    rng = np.random.default_rng(13)
    x = bf16_round(np.maximum(rng.standard_normal((2, h, h, c)), 0))  # ReLU-like: many exact ties at 0
    y_ref, arg_ref = O.maxpool_3x3s2(x)
    p = (h - 3) // 2 + 1
    dy = bf16_round(rng.standard_normal((2, p, p, c)))
    xd = dev(x, torch.bfloat16)
    y = torch.empty(2, p, p, c, dtype=torch.bfloat16, device="cuda")
    arg = torch.empty(2, p, p, c, dtype=torch.uint8, device="cuda")
    nv.call("vl_maxpool_fwd", xd, y, arg, 2, h, h, c)
    assert np.array_equal(y.float().cpu().numpy(), y_ref)          # max of bf16 values: exact
    assert np.array_equal(arg.cpu().numpy(), arg_ref.astype(np.uint8))  # index work: bit-exact, first max wins
    dx = torch.empty_like(xd)
    nv.call("vl_maxpool_bwd", dev(dy, torch.bfloat16), arg, dx, xd, 2, h, h, c)
    dx_ref = O.maxpool_3x3s2_backward(x.shape, arg_ref, dy) * (x > 0)
    assert rel(dx.float().cpu().numpy(), dx_ref) < 1e-2  # sums of <= 4 bf16 values, rounded to bf16


@pytest.mark.parametrize("batch,t_len,d_in,hidden,cluster", [(3, 4, 64, 32, False), (5, 16, 128, 256, False),
                                                             (5, 16, 128, 256, True), (13, 3, 64, 256, True),
                                                             (64, 16, 64, 256, True)])
def test_lstm_fwd_bwd_vs_oracle(vl, batch, t_len, d_in, hidden, cluster):
    nv = vl["nv"]
    rng = np.random.default_rng(14)
    kern = (rng.uniform(-1, 1, size=(d_in + hidden, 4 * hidden)) * 0.2).astype(np.float32)
    bias = (rng.standard_normal(4 * hidden) * 0.1).astype(np.float32)
    x = rng.standard_normal((batch, t_len, d_in)).astype(np.float32)
    out_ref, caches = O.lstm_forward(x, [kern], [bias])
    dout = rng.standard_normal(out_ref.shape).astype(np.float32)
    dx_ref, dks, dbs = O.lstm_backward(caches, dout)
    n = batch * t_len
    gx = dev((x.reshape(n, d_in) @ kern[:d_in] + bias).astype(np.float32))
    wh = dev(kern[d_in:])
    acts = torch.empty(n, 4 * hidden, device="cuda")
    cs = torch.empty(n, hidden, device="cuda")
    hseq = torch.empty(n, hidden, device="cuda")
    hprev = torch.empty(n, hidden, dtype=torch.bfloat16, device="cuda")
    nv.call("vl_lstm_fwd_cluster" if cluster else "vl_lstm_fwd", gx, wh, acts, cs, hseq, None, hprev, batch, t_len,
            hidden, 1.0)
    assert rel(hseq.cpu().numpy().reshape(out_ref.shape), out_ref) < 1e-4  # fp32 kernel
    dg = torch.empty(n, 4 * hidden, dtype=torch.bfloat16, device="cuda")
    if cluster:
        nv.call("vl_lstm_bwd_cluster", dev(dout.reshape(n, hidden)), acts, cs, wh, dg, batch, t_len, hidden)
    else:
        wht = dev(np.ascontiguousarray(kern[d_in:].T))
        nv.call("vl_lstm_bwd", dev(dout.reshape(n, hidden)), acts, cs, wht, dg, batch, t_len, hidden)
    dgf = dg.float().cpu().numpy()
    # d(kernel) = [x, h_prev]^T dg ; compare against the oracle through the gate gradients
    xh = np.concatenate([x.reshape(n, d_in), hprev.float().cpu().numpy()], axis=1)
    assert rel(xh.T @ dgf, dks[0]) < BF16_TOL
    assert rel(dgf.sum(axis=0), dbs[0]) < BF16_TOL
    assert rel(dgf @ kern[:d_in].T, dx_ref.reshape(n, d_in)) < BF16_TOL


@pytest.mark.parametrize("h,c", [(57, 96), (28, 256)])
def test_fused_lrn_pool_matches_unfused_and_oracle(vl, h, c):
    nv = vl["nv"]
    rng = np.random.default_rng(18)
    n = 2
    x = bf16_round(np.maximum(rng.standard_normal((n, h, h, c)) * 30, 0))
    p = (h - 3) // 2 + 1
    dy = bf16_round(rng.standard_normal((n, p, p, c)))
    xd, dyd = dev(x, torch.bfloat16), dev(dy, torch.bfloat16)
    # forward: fused == lrn_fwd -> maxpool_fwd up to one bf16 rounding of lrn(x) (fp32 contraction differences)
    nrm = torch.empty_like(xd)
    nv.call("vl_lrn_fwd", xd, nrm, n * h * h, c, 2, 2e-05, 0.75, 1.0)
    y0 = torch.empty(n, p, p, c, dtype=torch.bfloat16, device="cuda")
    a0 = torch.empty(n, p, p, c, dtype=torch.uint8, device="cuda")
    nv.call("vl_maxpool_fwd", nrm, y0, a0, n, h, h, c)
    y1, a1 = torch.empty_like(y0), torch.empty_like(a0)
    nv.call("vl_lrn_pool_fwd", xd, y1, a1, n, h, h, c, 2, 2e-05, 0.75, 1.0)
    assert rel(y1.float().cpu().numpy(), y0.float().cpu().numpy()) < 1e-2
    assert (a0 == a1).float().mean().item() > 0.995  # argmax may move only between (near-)tied window entries
    y_ref, arg_ref = O.maxpool_3x3s2(O.bf16_round(O.lrn(x)))
    assert rel(y1.float().cpu().numpy(), y_ref) < 1e-2
    # backward: fused == maxpool_bwd -> lrn_bwd(+relu) within rounding, and matches the oracle; bias gradient too
    dn = torch.empty_like(xd)
    nv.call("vl_maxpool_bwd", dyd, a0, dn, None, n, h, h, c)
    dx0 = torch.empty_like(xd)
    nv.call("vl_lrn_bwd", xd, dn, dx0, n * h * h, c, 2, 2e-05, 0.75, 1.0, 1)
    dx1 = torch.empty_like(xd)
    db = torch.zeros(c, device="cuda")
    nv.call("vl_pool_lrn_bwd", xd, dyd, a0, dx1, db, n, h, h, c, 2, 2e-05, 0.75, 1.0)
    assert rel(dx1.float().cpu().numpy(), dx0.float().cpu().numpy()) < 1e-2
    dn_ref = O.bf16_round(O.maxpool_3x3s2_backward(x.shape, a0.cpu().numpy().astype(np.int64), dy))
    dx_ref = O.lrn_backward(x, dn_ref) * (x > 0)
    assert rel(dx1.float().cpu().numpy(), dx_ref) < 1e-2
    # the bias gradient sums the fp32 values before their bf16 rounding (closer to the reference's fp32 BiasAddGrad than
    # the sum of the stored bf16 values): the two differ by the rounding noise of n*h*w terms
    assert rel(db.cpu().numpy(), dx1.float().cpu().numpy().reshape(-1, c).sum(0)) < 5e-3
    assert rel(db.cpu().numpy(), dx_ref.reshape(-1, c).sum(0)) < 5e-3


def test_segment_pool_bit_exact_vs_numpy(vl):
    """Clip->video fusion (val.py:158-167): np.mean(axis=0) / last row, variable clips per video -> bit exact."""
    nv = vl["nv"]
    rng = np.random.default_rng(15)
    cpv = [3, 1, 7, 25, 2]
    seg = np.concatenate([[0], np.cumsum(cpv)]).astype(np.int32)
    x = (rng.standard_normal((seg[-1], 101)) * 30).astype(np.float32)
    for mode, fn in ((0, lambda r: np.mean(r, axis=0)), (1, lambda r: r[-1]), (2, lambda r: r.max(axis=0))):
        y = torch.empty(len(cpv), 101, device="cuda")
        nv.call("vl_segment_pool_fwd", dev(x), dev(seg), 0, len(cpv), 101, mode, y, None)
        ref = np.stack([fn(x[seg[i]:seg[i + 1]]) for i in range(len(cpv))])
        assert np.array_equal(y.cpu().numpy(), ref.astype(np.float32)), mode
        assert np.array_equal(y.cpu().numpy().argmax(1), ref.argmax(1))
    # fixed-length segments (temporal fusion over fpc, tf_util.py:4-30) and its gradient
    xf = rng.standard_normal((6 * 16, 256)).astype(np.float32)
    y = torch.empty(6, 256, device="cuda")
    nv.call("vl_segment_pool_fwd", dev(xf), None, 16, 6, 256, 0, y, None)
    assert rel(y.cpu().numpy(), O.temporal_fusion(xf.reshape(6, 16, 256), "avg")) < 1e-6
    dy = rng.standard_normal((6, 256)).astype(np.float32)
    for mode, name in ((0, "avg"), (1, "last")):
        dx = torch.empty(6 * 16, 256, device="cuda")
        nv.call("vl_segment_pool_bwd", dev(dy), None, 16, 6, 256, mode, dx)
        ref = O.temporal_fusion_backward((6, 16, 256), name, dy).reshape(96, 256)
        assert rel(dx.cpu().numpy(), ref) < 1e-6


def test_softmax_ce_vs_oracle(vl):
    nv = vl["nv"]
    rng = np.random.default_rng(16)
    rows, c = 37, 101
    logits = (rng.standard_normal((rows, c)) * 5).astype(np.float32)
    logits[3, 7] = logits[3, 2] = logits[3].max() + 1  # argmax tie: lowest index wins
    labels = np.zeros((rows, c), np.int32)
    labels[np.arange(rows), rng.integers(0, c, rows)] = 1
    labels[3] = 0
    labels[3, 2] = 1
    loss_ref, dl_ref, per_ref = O.softmax_ce(logits, labels)
    row_loss = torch.empty(2 * rows, device="cuda")
    scal = torch.zeros(2, device="cuda")
    dl = torch.empty(rows, 104, device="cuda")
    dlb = torch.empty(rows, 104, dtype=torch.bfloat16, device="cuda")
    nv.call("vl_softmax_ce", dev(logits), dev(labels), rows, c, 1.0 / rows, row_loss, scal, dl, dlb, 104)
    assert abs(scal[0].item() - loss_ref) < 1e-5 * abs(loss_ref)
    assert rel(row_loss[:rows].cpu().numpy(), per_ref) < 1e-5
    assert rel(dl.cpu().numpy()[:, :c], dl_ref) < 1e-5
    assert np.all(dl.cpu().numpy()[:, c:] == 0)
    correct_ref = (logits.argmax(1) == labels.argmax(1))
    assert np.array_equal(row_loss[rows:].cpu().numpy().astype(bool), correct_ref)  # integer output: bit-exact
    assert scal[1].item() == correct_ref.sum()


def test_optimizer_kernels_vs_oracle(vl):
    nv = vl["nv"]
    rng = np.random.default_rng(17)
    sizes = [1000, 64, 70000, 333]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + (-(-s // 64) * 64))
    n = offs[-1]
    g = np.zeros(n, np.float32)
    w = np.zeros(n, np.float32)
    grads, params = {}, {}
    for i, s in enumerate(sizes):
        grads[str(i)] = (rng.standard_normal(s) * 3).astype(np.float32)
        params[str(i)] = rng.standard_normal(s).astype(np.float32)
        g[offs[i]:offs[i] + s] = grads[str(i)]
        w[offs[i]:offs[i] + s] = params[str(i)]
    clipped, gn = O.clip_by_global_norm(grads, 10.0)
    gd, wd = dev(g), dev(w)
    sq = torch.zeros(len(sizes), device="cuda")
    scal = torch.zeros(8, device="cuda")
    ws = torch.empty(int(nv.lib().vl_grad_sqnorms_workspace(n, len(sizes))), device="cuda")
    nv.call("vl_grad_sqnorms", gd, n, dev(np.array(offs, np.int64)), len(sizes), sq, ws, ws.numel())
    sq_again = torch.empty_like(sq)
    nv.call("vl_grad_sqnorms", gd, n, dev(np.array(offs, np.int64)), len(sizes), sq_again, ws, ws.numel())
    assert torch.equal(sq, sq_again)  # deterministic reduction: data-parallel ranks derive the same clip scale
    nv.call("vl_clip_scalars", sq, len(sizes), 10.0, 1.0, scal)
    assert abs(scal[0].item() - gn) < 1e-4 * gn
    assert abs(scal[2].item() - O.mean_grad_norm(clipped)) < 1e-4 * gn
    p_sgd = {k: v.copy() for k, v in params.items()}
    O.sgd_update(p_sgd, clipped, 0.01)
    nv.call("vl_sgd_update", wd, gd, n, 0.01, scal, 1.0)
    got = wd.cpu().numpy()
    for i, s in enumerate(sizes):
        assert rel(got[offs[i]:offs[i] + s], p_sgd[str(i)]) < 1e-5
    # the fused variant: identical weights, and the bf16 operand copies of two ranges written in the same pass
    import ctypes
    wd2 = dev(w)
    segs = [(offs[1], offs[1] + (sizes[1] // 4) * 4), (offs[3], offs[3] + (sizes[3] // 4) * 4)]
    shadows = [torch.full((e - b,), float("nan"), dtype=torch.bfloat16, device="cuda") for b, e in segs]
    nv.call("vl_sgd_update_shadow", wd2, gd, n, 0.01, scal, 1.0, 2, (ctypes.c_int64 * 2)(*[b for b, _ in segs]),
            (ctypes.c_int64 * 2)(*[e for _, e in segs]), (ctypes.c_void_p * 2)(*[t.data_ptr() for t in shadows]))
    assert torch.equal(wd2, wd)
    for (b, e), t in zip(segs, shadows):
        assert torch.equal(t, wd[b:e].to(torch.bfloat16))
    # Adam, three steps
    wd = dev(w)
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    p_adam = {k: v_.copy() for k, v_ in params.items()}
    st = {}
    for t in range(1, 4):
        O.adam_update(p_adam, clipped, st, 0.01)
        nv.call("vl_adam_update", wd, gd, m, v, n, 0.01, 0.9, 0.999, 1e-8, t, scal, 1.0)
    got = wd.cpu().numpy()
    for i, s in enumerate(sizes):
        assert rel(got[offs[i]:offs[i] + s], p_adam[str(i)]) < 1e-4


def test_dropout_mask_statistics(vl):
    nv = vl["nv"]
    mask = torch.empty(64 * 256, device="cuda")
    nv.call("vl_dropout_mask", mask, mask.numel(), 0.5, 1234, 0)
    vals = set(np.unique(mask.cpu().numpy()).tolist())
    assert vals == {0.0, 2.0}
    assert abs((mask > 0).float().mean().item() - 0.5) < 0.03
    mask2 = torch.empty_like(mask)
    nv.call("vl_dropout_mask", mask2, mask.numel(), 0.5, 1234, 0)
    assert torch.equal(mask, mask2)  # counter based: reproducible


@pytest.mark.parametrize("workflow", ["lrcn", "singleframe"])
def test_shadow_gather_equals_the_pack_kernels(vl, workflow):
    """Engine.refresh_shadows refreshes every permuted / padded bf16 operand copy with ONE vl_gather_bf16 launch over a
    precomputed index table (shadow_table.py); each copy must be bit-identical to the stand-alone packing kernel whose
    layout it restates (include/vlb200.h)."""
    E, nv, K = vl["E"], vl["nv"], vl["K"]
    cfg = E.EngineConfig(workflow=workflow, fusion="avg", fpc=2, num_classes=101, lstm_hidden=256)
    eng = E.Engine(cfg, max_clips=1, params=E.init_variables(cfg, seed=17))
    sp, sh = eng.sp, eng.sh
    s1, s1s = sp["conv1"], sp["conv1_s2d"]
    ref = torch.empty_like(sh["conv1_fwd"])
    nv.call("vl_s2d_pack_filter", eng.var("dcnn/conv1W"), ref, s1.kh, s1.kw, 3, 96, s1.stride, s1s.cchunks * 64, 1)
    assert torch.equal(ref, sh["conv1_fwd"])
    for name in ("conv2", "conv3", "conv4", "conv5"):
        s = sp[name]
        w = eng.var2d("dcnn/%sW" % name)
        ref = torch.empty_like(sh[name])
        nv.call("vl_cast_f32_to_bf16", w, ref, w.numel())
        assert torch.equal(ref, sh[name]), name
        ref = torch.empty_like(sh[name + "_fwd"])
        nv.call("vl_pack_bf16_t", w, s.taps * s.cin_g, s.cout, ref, s.k_packed, s.cin_g, s.cchunks * 64)
        assert torch.equal(ref, sh[name + "_fwd"]), name
    s2 = sp["conv2"]
    ref = torch.empty_like(sh["conv2_d2s"])
    nv.call("vl_pack_dgrad_d2s", eng.var("dcnn/conv2W"), ref, s2.kh, s2.kw, s2.cin_g, s2.cout_g, s2.groups, 2, 2)
    assert torch.equal(ref, sh["conv2_d2s"])
    for key, wname in (("fc8", "dcnn/fc8W"), ("output_fc", "output_fc_w")):
        if key in sh:
            w = eng.var(wname)
            ref = torch.empty_like(sh[key])
            nv.call("vl_pack_bf16", w, w.shape[0], w.shape[1], ref, w.shape[0], eng.c_pad, w.shape[0], w.shape[0])
            assert torch.equal(ref, sh[key]), key
    assert ("fc8" in sh) == (workflow == "singleframe") and ("output_fc" in sh) == (workflow == "lrcn")
    # same-layout copies
    for key, wname in (("fc6", "dcnn/fc6W"), ("fc7", "dcnn/fc7W")):
        assert torch.equal(eng.var(wname).to(torch.bfloat16), sh[key]), key


# ----------------------------------------------------------------------------------------------------------
# whole path
# ----------------------------------------------------------------------------------------------------------
def _problem(cfg_kwargs, clips, seed=21):
    from vlb200 import engine as E
    cfg = E.EngineConfig(**cfg_kwargs)
    params = E.init_variables(cfg, seed=seed)
    rng = np.random.default_rng(seed + 1)
    # unit-range pixels keep the sigma=0.05 random init out of gate saturation (see tests/test_oracle.py)
    frames = rng.uniform(-1, 1, size=(clips * cfg.fpc, cfg.height, cfg.width, 3)).astype(np.float32)
    labels = rng.integers(0, cfg.num_classes, clips)
    onehot = np.zeros((clips, cfg.num_classes), np.int32)
    onehot[np.arange(clips), labels] = 1
    return cfg, params, frames, onehot


@pytest.mark.parametrize("kw", [
    dict(workflow="lrcn", fusion="avg", fpc=4, num_classes=101, lstm_hidden=256, lstm_layers=1),
    dict(workflow="lrcn", fusion="last", fpc=3, num_classes=11, lstm_hidden=64, lstm_layers=2),
    dict(workflow="singleframe", fusion="avg", fpc=4, num_classes=101),
])
def test_forward_logits_and_labels_vs_oracle(vl, kw):
    E = vl["E"]
    clips = 3
    cfg, params, frames, onehot = _problem(kw, clips)
    eng = E.Engine(cfg, max_clips=clips, params=params)
    logits = eng.forward(frames)
    if kw["workflow"] == "lrcn":
        ref, _ = O.lrcn_forward(params, frames, cfg.fpc, cfg.fusion, cfg.frame_encoding_layer)
        ref_q, _ = O.lrcn_forward(params, frames, cfg.fpc, cfg.fusion, cfg.frame_encoding_layer, q=O.bf16_round)
    else:
        ref, _ = O.singleframe_forward(params, frames, cfg.fpc, cfg.fusion)
        ref_q, _ = O.singleframe_forward(params, frames, cfg.fpc, cfg.fusion, q=O.bf16_round)
    assert logits.shape == ref.shape and logits.dtype == np.float32
    # north_star: logits within 2e-2 (bf16 path) of the fp32 reference semantics ...
    assert rel(logits, ref) < BF16_TOL
    # ... and much tighter against the oracle evaluated with the device's storage precision
    assert rel(logits, ref_q) < 1e-2
    # predicted labels: bit-exact wherever the oracle's top-2 margin exceeds the bf16 tolerance
    srt = np.sort(ref, axis=1)
    margin_ok = (srt[:, -1] - srt[:, -2]) > 2 * BF16_TOL * np.abs(ref).max()
    assert np.array_equal(logits.argmax(1)[margin_ok], ref.argmax(1)[margin_ok])
    # ... and against the bf16-storage oracle (same hard decisions) wherever ITS margin exceeds the measured distance
    srt_q = np.sort(ref_q, axis=1)
    margin_q = (srt_q[:, -1] - srt_q[:, -2]) > 2 * rel(logits, ref_q) * np.abs(ref_q).max()
    assert np.array_equal(logits.argmax(1)[margin_q], ref_q.argmax(1)[margin_q])


@pytest.mark.parametrize("kw,opt", [
    (dict(workflow="lrcn", fusion="avg", fpc=4, num_classes=101, lstm_hidden=256, clip_norm=10), "sgd"),
    (dict(workflow="lrcn", fusion="last", fpc=2, num_classes=11, lstm_hidden=64, lstm_layers=2, clip_norm=None), "adam"),
    (dict(workflow="singleframe", fusion="avg", fpc=2, num_classes=101, clip_norm=10), "sgd"),
    # a non-AlexNet-default input size: the generic (runtime-geometry) LRN / pooling / convolution paths
    (dict(workflow="lrcn", fusion="avg", fpc=2, num_classes=23, lstm_hidden=256, clip_norm=10, height=131, width=147), "sgd"),
])
def test_train_step_vs_oracle(vl, kw, opt):
    E = vl["E"]
    clips = 2
    kw = dict(kw, optimizer=opt, dropout_keep_prob=0.5 if kw["workflow"] == "lrcn" else 0.0)
    cfg, params, frames, onehot = _problem(kw, clips)
    mask = None
    if cfg.workflow == "lrcn":
        mask = (np.random.default_rng(5).uniform(size=(clips, cfg.lstm_hidden)) < 0.5).astype(np.float32) * 2.0
    eng = E.Engine(cfg, max_clips=clips, params=params)
    lr = 1e-3
    loss, lr_out, gstep, acc, gnorm = eng.train_step(frames, onehot, lr, dropout_mask=mask)
    # reference 1: fp32 semantics of the TF graph (north_star tolerance on the loss)
    res32 = O.train_step({k: v.copy() for k, v in params.items()}, frames, onehot, cfg.fpc, lr, workflow=cfg.workflow,
                         fusion=cfg.fusion, frame_encoding_layer=cfg.frame_encoding_layer, clip_norm=cfg.clip_norm,
                         optimizer=opt, opt_state={}, dropout_mask=mask)
    assert abs(loss - res32["loss"]) < BF16_TOL * max(1.0, abs(res32["loss"]))
    # reference 2: same graph with the device's storage precision (bf16 operands/activations, fp32 accumulate)
    p_ref = {k: v.copy() for k, v in params.items()}
    st = {}
    res = O.train_step(p_ref, frames, onehot, cfg.fpc, lr, workflow=cfg.workflow, fusion=cfg.fusion,
                       frame_encoding_layer=cfg.frame_encoding_layer, clip_norm=cfg.clip_norm, optimizer=opt,
                       opt_state=st, dropout_mask=mask, q=O.bf16_round)
    assert gstep == 1 and lr_out == lr
    assert abs(loss - res["loss"]) < 2e-3 * max(1.0, abs(res["loss"]))
    assert acc == float(res["accuracy"])
    # Gradients.  ReLU masks and pool argmaxes are hard decisions: a 1-ulp bf16 difference in a forward activation
    # flips a few of them, and with a 4-frame batch one flipped element changes a whole filter-gradient column.
    # Backward parity is therefore checked CONDITIONED on the device's forward state: the oracle's backward
    # restatement is run on the activations the device produced (forward parity is asserted separately above and
    # in test_forward_logits_and_labels_vs_oracle).
    g_ref, gn_ref = _oracle_backward_on_device_state(eng, cfg, params, frames, onehot, mask)
    got = eng.gradient_dict()
    errs = {k: (rel(got[k], g), rel_l2(got[k], g)) for k, g in g_ref.items()}
    print("gradient errors (max-abs rel, l2 rel):", {k: "%.1e/%.1e" % e for k, e in errs.items()})
    assert set(got) == set(g_ref)
    assert max(e[0] for e in errs.values()) < BF16_TOL, errs
    assert max(e[1] for e in errs.values()) < BF16_TOL, errs
    scale = cfg.clip_norm / max(gn_ref, cfg.clip_norm) if cfg.clip_norm else 1.0
    clipped = {k: v * scale for k, v in g_ref.items()}
    assert abs(gnorm - O.mean_grad_norm(clipped)) < BF16_TOL * O.mean_grad_norm(clipped)
    # updated variables: apply the oracle's optimiser to the conditioned gradients
    p_ref = {k: v.copy() for k, v in params.items()}
    if opt == "sgd":
        O.sgd_update(p_ref, clipped, lr)
    else:
        O.adam_update(p_ref, clipped, {}, lr)
    sd = eng.state_dict()
    for k, v in p_ref.items():
        step = np.abs(v - params[k]).max()
        bad = np.abs(sd[k] - v) > 5e-2 * step + 1e-7
        if opt == "adam":
            # Adam's first step is lr * sign(g): elements whose tiny gradient flips sign under rounding move the
            # other way; allow a small fraction of such elements
            assert bad.mean() < 0.02, (k, bad.mean())
        else:
            assert not bad.any(), (k, np.abs(sd[k] - v).max(), step)
    assert int(sd["global_step"]) == 1


@pytest.mark.parametrize("kw", [
    # lstm fusion `state` (lstm.py:81,91-93; model.py:137-143): logits = fc_convert(final h of the top layer), no dropout
    dict(workflow="lrcn", fusion="state", fpc=3, num_classes=11, lstm_hidden=64, lstm_layers=2, dropout_keep_prob=0.5),
    dict(workflow="lrcn", fusion="state", fpc=2, num_classes=64, lstm_hidden=64),  # dims agree: no fc at all
    # classifier fc with EARLY fusion of the fc7 features (model.py:103-108), then convert_dim_fc
    dict(workflow="fc", frame_encoding_layer="fc7", fusion="avg", early_fusion=True, fpc=3, num_classes=11),
    dict(workflow="fc", frame_encoding_layer="fc6", fusion="last", early_fusion=True, fpc=2, num_classes=101),
    # classifier fc on fc7 features per frame, LATE fusion of the logits (model.py:149-151)
    dict(workflow="fc", frame_encoding_layer="fc7", fusion="avg", early_fusion=False, fpc=2, num_classes=101),
])
def test_fusion_variants_forward_and_train_step_vs_oracle(vl, kw):
    """The pipeline variants of Model.build_pipeline beyond the two BASELINE workflows, against oracle/lrcn_torch.py:
    forward logits (fp32 oracle: north-star tolerance; bf16-storage oracle: tighter) and one SGD step (loss, and the
    applied update of every variable against the bf16-storage oracle's update)."""
    from oracle import lrcn_torch as T
    E = vl["E"]
    clips = 3
    kw = dict(kw, optimizer="sgd", clip_norm=10)
    cfg, params, frames, onehot = _problem(kw, clips)
    names = [n for n, _ in E.variable_shapes(cfg)]
    if kw["fusion"] == "state":
        assert ("fc_convert_w" in names) == (cfg.lstm_hidden != cfg.num_classes) and "output_fc_w" not in names
    if kw["workflow"] == "fc":
        assert "fc_convert_w" in names and not any(n.startswith("rnn/") for n in names)
        assert ("dcnn/fc7W" in names) == (kw["frame_encoding_layer"] == "fc7") and "dcnn/fc8W" not in names
    fusion = cfg.fusion
    if cfg.workflow == "fc":
        fusion = (cfg.fusion, None) if cfg.early_fusion else (None, cfg.fusion)
    x = torch.tensor(frames)
    with torch.no_grad():
        ref = T.logits_fn(T.to_torch(params), x, cfg.fpc, cfg.workflow, fusion, cfg.frame_encoding_layer).numpy()
        ref_q = T.logits_fn(T.to_torch(params), x, cfg.fpc, cfg.workflow, fusion, cfg.frame_encoding_layer, q=True).numpy()
    eng = E.Engine(cfg, max_clips=clips, params=params)
    logits = eng.forward(frames)
    assert logits.shape == ref.shape == (clips, cfg.num_classes)
    if "fc_convert_w" not in names:
        # the logits ARE the final hidden state: tanh(c) * sigmoid(o) of gates that the 4096-wide random projection
        # saturates, i.e. every unit whose pre-activation lies within rounding distance of zero may flip on its own and
        # nothing downstream averages the flips out.  Element-wise agreement instead of a max-norm bound.
        close = np.abs(logits - ref_q) < BF16_TOL
        print("state fusion without fc: %.1f %% of the units within %.0e of the bf16-storage oracle" % (
            100 * close.mean(), BF16_TOL))
        assert close.mean() > 0.9
        loss, _, gstep, acc, gnorm = eng.train_step(frames, onehot, 1e-2)
        assert gstep == 1 and np.isfinite(loss) and np.isfinite(gnorm)
        return
    assert rel(logits, ref) < BF16_TOL and rel(logits, ref_q) < 1e-2
    lr = 1e-2
    loss, _, gstep, acc, gnorm = eng.train_step(frames, onehot, lr)  # dropout_keep_prob > 0 must be ignored for `state`
    P = T.to_torch(params, requires_grad=True)
    res = T.train_step(P, x, torch.tensor(onehot), cfg.fpc, lr, cfg.workflow, fusion, cfg.frame_encoding_layer,
                       clip_norm=10, q=True)
    assert gstep == 1 and abs(loss - res["loss"]) < 2e-3 * max(1.0, abs(res["loss"]))
    assert abs(acc - res["accuracy"]) < 1e-6
    sd = eng.state_dict()
    errs = {}
    for name in names:
        upd_ref = P[name].detach().numpy() - params[name]
        if np.abs(upd_ref).max() == 0:
            assert np.array_equal(sd[name], params[name]), name
            continue
        errs[name] = rel_l2(sd[name] - params[name], upd_ref)
    print("update errors (l2 rel vs bf16-storage oracle):", {k: "%.1e" % v for k, v in errs.items()})
    # What this test adds is the HEAD wiring of the variant (the encoder backward is the code path of every workflow and
    # is held to 2e-2 conditioned on the device's forward state in test_train_step_vs_oracle).  With 6-9 frames in the
    # batch a single flipped ReLU / arg-max decision moves a whole filter-gradient column, and the torch oracle models
    # the storage points of the forward pass only: head variables tight, everything below by direction (l2 < 0.2 is a
    # cosine above 0.98).
    head = [n for n in errs if not n.startswith("dcnn/conv") and not n.startswith("dcnn/fc6")]
    assert max(errs[n] for n in head) < 8e-2, errs
    assert max(errs.values()) < 0.2, errs


def _oracle_backward_on_device_state(eng, cfg, params, frames, onehot, mask):
    """Run the oracle's backward restatement on the forward activations the device produced."""
    q = O.bf16_round
    A = eng.A
    n = frames.shape[0]
    b = n // cfg.fpc

    def f(t):
        return t.float().cpu().numpy()

    ac = {k: f(A[k][:n]) for k in ("a1", "p1", "a2", "p2", "a3", "a4", "a5", "p5", "f6", "f7")}
    ac["n1"], ac["n2"] = q(O.lrn(ac["a1"])), q(O.lrn(ac["a2"]))  # never materialised on the device (fused LRN+pool)
    for k in ("arg1", "arg2", "arg5"):
        ac[k] = A[k][:n].cpu().numpy().astype(np.int64)
    ac["x0"] = q(frames)
    ac["flat"] = ac["p5"].reshape(n, -1)
    ac["final_layer"] = cfg.frame_encoding_layer
    logits = f(A["logits"][:b])
    _, dlogits, _ = O.softmax_ce(logits, onehot)
    if cfg.workflow == "singleframe":
        cache = dict(ac=ac, seq_shape=(b, cfg.fpc, cfg.num_classes), fusion=cfg.fusion)
        grads = O.singleframe_backward(params, cache, dlogits, q)
    else:
        hd, t_len = cfg.lstm_hidden, cfg.fpc
        feat = ac["f7"] if cfg.frame_encoding_layer == "fc7" else ac["f6"]
        lc = []
        inp = feat.reshape(b, t_len, -1)
        for layer in range(cfg.lstm_layers):
            acts = f(A["acts%d" % layer][:n]).reshape(b, t_len, 4 * hd)
            cs = f(A["cs%d" % layer][:n]).reshape(b, t_len, hd)
            hs = f(A["hseq%d" % layer][:n]).reshape(b, t_len, hd)
            steps = []
            for t in range(t_len):
                si, tj, sf, so = np.split(acts[:, t], 4, axis=1)
                h_prev = hs[:, t - 1] if t > 0 else np.zeros((b, hd), np.float32)
                c_prev = cs[:, t - 1] if t > 0 else np.zeros((b, hd), np.float32)
                steps.append((q(inp[:, t]), h_prev, si, sf, so, tj, c_prev, np.tanh(cs[:, t])))
            kern = params["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer]
            lc.append((inp.shape, steps, kern))
            inp = hs
        dropped = f(A["dropped"][:b]) if mask is not None else f(A["fused"][:b])
        cache = dict(ac=ac, lc=lc, outs_shape=(b, t_len, hd), dropped=dropped, fusion=cfg.fusion, mask=mask,
                     feat_shape=feat.shape)
        grads = O.lrcn_backward(params, cache, dlogits, q)
    return grads, float(O.global_norm(grads))


def test_half_batch_forward_chains_equal_the_single_chain(vl, monkeypatch):
    """The forward convolution stack runs as two half-batch chains on two streams from 256 frames on (engine._encoder_fwd);
    logits and one train step must be bit-identical to the single chain (VL_FWD_HALVES=0): same kernels, same per-frame
    tile arithmetic, only the stream schedule differs."""
    E = vl["E"]
    clips, fpc, c = 16, 16, 101  # 256 frames: the smallest batch that takes the two-chain path
    cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=fpc, num_classes=c, lstm_hidden=256, clip_norm=10,
                         dropout_keep_prob=0.0, optimizer="sgd", mean=(99.197148, 105.293620, 109.503945))
    params = E.init_variables(cfg, seed=21)
    rng = np.random.default_rng(4)
    frames = rng.integers(0, 256, size=(clips * fpc, 227, 227, 3), dtype=np.uint8)
    onehot = np.zeros((clips, c), np.int32)
    onehot[np.arange(clips), rng.integers(0, c, clips)] = 1
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("VL_FWD_HALVES", mode)
        eng = E.Engine(cfg, max_clips=clips, params=params)
        logits = eng.forward(frames)
        step = eng.train_step(frames, onehot, 1e-3)
        out[mode] = (logits, step, eng.state_dict()["dcnn/conv1W"])
        del eng
    assert np.array_equal(out["1"][0], out["0"][0])
    assert out["1"][1][0] == out["0"][1][0] and out["1"][1][3] == out["0"][1][3]  # loss, accuracy
    # conv1's filter gradient is a split-K red.add sum (order varies run to run): compare within rounding
    assert rel(out["1"][2], out["0"][2]) < 1e-5


def test_full_size_properties_config2(vl):
    """BASELINE.json configs[1] at its FULL size (64 clips x 16 frames per GPU), where the oracle would take minutes:
    size-independent properties of the path.
      (1) batch independence: the first 8 clips of the 64-clip forward == the 8-clip forward, bit for bit
          (every frame / clip is processed by the same tile arithmetic whatever its neighbours are);
      (2) permutation: reversing the clip order reverses the logits rows bit for bit, and the mean loss of the train
          step is unchanged up to the order of the final sums;
      (3) clip -> video fusion of the 64 rows on the device == numpy on the host (val.py:158-167), labels bit exact;
      (4) loss / accuracy read back from the step equal the softmax cross-entropy of the device logits."""
    E, nv = vl["E"], vl["nv"]
    clips, fpc, c = 64, 16, 101
    cfg = E.EngineConfig(workflow="lrcn", fusion="avg", fpc=fpc, num_classes=c, lstm_hidden=256, clip_norm=10,
                         dropout_keep_prob=0.0, optimizer="sgd", mean=(99.197148, 105.293620, 109.503945))
    eng = E.Engine(cfg, max_clips=clips, params=E.init_variables(cfg, seed=3))
    g = torch.Generator(device="cuda").manual_seed(5)
    frames = torch.randint(0, 256, (clips * fpc, 227, 227, 3), dtype=torch.uint8, device="cuda", generator=g)
    labels = np.random.default_rng(6).integers(0, c, clips)
    onehot = np.zeros((clips, c), np.int32)
    onehot[np.arange(clips), labels] = 1
    full = eng.forward(frames)
    assert full.shape == (clips, c) and np.isfinite(full).all()
    # (1)
    part = eng.forward(frames[:8 * fpc].contiguous())
    assert np.array_equal(part, full[:8])
    # (2) forward
    rev_frames = frames.view(clips, fpc, 227, 227, 3).flip(0).reshape(clips * fpc, 227, 227, 3).contiguous()
    rev = eng.forward(rev_frames)
    assert np.array_equal(rev[::-1], full)
    # (3) videos of 1..7 clips
    cpv = [3, 1, 7, 5, 2, 6, 4, 1, 7, 3, 5, 2, 6, 4, 7, 1]
    assert sum(cpv) == clips
    seg = np.concatenate([[0], np.cumsum(cpv)]).astype(np.int32)
    y = torch.empty(len(cpv), c, device="cuda")
    nv.call("vl_segment_pool_fwd", dev(full), dev(seg), 0, len(cpv), c, 0, y, None)
    ref = np.stack([np.mean(full[seg[i]:seg[i + 1]], axis=0) for i in range(len(cpv))]).astype(np.float32)
    assert np.array_equal(y.cpu().numpy(), ref) and np.array_equal(y.cpu().numpy().argmax(1), ref.argmax(1))
    # (4) + (2) train step (no update: the weights stay those of the forward passes above)
    loss, _, _, acc, gnorm = eng.train_step(frames, onehot, 1e-3, apply_update=False)
    z = full.astype(np.float64)
    lse = np.log(np.exp(z - z.max(1, keepdims=True)).sum(1)) + z.max(1)
    ce = float(np.mean(lse - z[np.arange(clips), labels]))
    assert abs(loss - ce) < 1e-4 * max(1.0, abs(ce))
    assert acc == float(np.mean(full.argmax(1) == labels))
    loss_rev, _, _, acc_rev, gnorm_rev = eng.train_step(rev_frames, onehot[::-1].copy(), 1e-3, apply_update=False)
    assert abs(loss_rev - loss) < 1e-5 * max(1.0, abs(loss)) and acc_rev == acc
    assert abs(gnorm_rev - gnorm) < 2e-2 * gnorm  # split-K atomics / bf16 partial sums: order-dependent rounding only
    # (5) fc6 / fc7 filter gradients at M = 1024 with the engine's own split-K choice and its zero-skip of the gradient
    # arena (Engine._split_k / _zero_grads): numpy on the operands the device holds (bf16 values, fp32 accumulate)
    A, G = eng.A, eng.G
    n = clips * fpc
    for wname, xbuf, dybuf in (("dcnn/fc7W", A["f6"][:n], G["df7"][:n]),
                               ("dcnn/fc6W", A["p5"][:n].view(n, -1), G["df6"][:n])):
        ref_dw = xbuf.float().cpu().numpy().T @ dybuf.float().cpu().numpy()
        got_dw = eng.var2d(wname, eng.grads).cpu().numpy()
        assert rel(got_dw, ref_dw) < 1e-3, wname
    before = eng.grads.clone()
    eng.train_step(rev_frames, onehot[::-1].copy(), 1e-3, apply_update=False)  # same step again: nothing stale adds up
    for wname in ("dcnn/fc7W", "dcnn/fc6W"):
        o = eng.var_off[wname]
        cnt = eng.var2d(wname).numel()
        assert torch.equal(before[o:o + cnt], eng.grads[o:o + cnt]), wname  # unsplit plain stores: bit-identical
    assert rel(eng.grads.cpu().numpy(), before.cpu().numpy()) < 1e-3
