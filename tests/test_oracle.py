"""The numpy oracle against (a) hand-computed micro cases of the TF op semantics and (b) an
independent torch-CPU implementation with autograd (oracle/lrcn_torch.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import lrcn_numpy as O
from oracle import lrcn_torch as T


def test_same_padding_geometry():
    # alexnet.py:76 conv1 11x11/4 SAME on 227 -> 57 with 4/4; conv2 5x5 -> 2/2; 3x3 -> 1/1
    assert O.same_pad(227, 11, 4) == (57, 4, 4)
    assert O.same_pad(28, 5, 1) == (28, 2, 2)
    assert O.same_pad(13, 3, 1) == (13, 1, 1)
    assert O.same_pad(8, 3, 2) == (4, 0, 1)  # TF puts the odd pixel AFTER


def test_conv_micro_hand_computed():
    x = np.arange(9, dtype=np.float32).reshape(1, 3, 3, 1)
    w = np.zeros((3, 3, 1, 1), np.float32)
    w[0, 0, 0, 0] = 1.0  # picks the upper-left neighbour (cross-correlation, no flip)
    w[1, 1, 0, 0] = 10.0
    y = O.conv2d_same(x, w, np.zeros(1, np.float32), 1, 1)[0, :, :, 0]
    exp = 10 * x[0, :, :, 0]
    exp[1:, 1:] += x[0, :2, :2, 0]
    assert np.array_equal(y, exp)


def test_lrn_micro():
    x = np.array([1, 2, 3, 4, 5, 6, 7], np.float32).reshape(1, 1, 1, 7)
    y = O.lrn(x)[0, 0, 0]
    sq = x[0, 0, 0] ** 2
    for d in range(7):
        s = 1.0 + 2e-05 * sq[max(0, d - 2):min(7, d + 3)].sum()  # alpha NOT divided by the window
        assert abs(y[d] - x[0, 0, 0, d] / s ** 0.75) < 1e-6


def test_maxpool_micro_first_max_wins():
    x = np.zeros((1, 5, 5, 1), np.float32)
    x[0, 1, 1, 0] = 3.0
    x[0, 2, 2, 0] = 3.0  # tie inside window (0,0): first in scan order gets the gradient
    y, arg = O.maxpool_3x3s2(x)
    assert y.shape == (1, 2, 2, 1) and y[0, 0, 0, 0] == 3.0
    dx = O.maxpool_3x3s2_backward(x.shape, arg, np.ones_like(y))
    assert dx[0, 1, 1, 0] == 1.0
    assert dx[0, 2, 2, 0] == 3.0  # max of the three other windows


def test_lstm_one_step_hand():
    rng = np.random.default_rng(0)
    kern = rng.standard_normal((3 + 2, 8)).astype(np.float32)
    bias = rng.standard_normal(8).astype(np.float32)
    x = rng.standard_normal((1, 1, 3)).astype(np.float32)
    out, _ = O.lstm_forward(x, [kern], [bias])
    g = x[0, 0] @ kern[:3] + bias
    i, j, f, o = g[0:2], g[2:4], g[4:6], g[6:8]  # BasicLSTMCell order i, j, f, o
    sig = lambda v: 1 / (1 + np.exp(-v))
    c = sig(i) * np.tanh(j)  # c0 = 0
    h = np.tanh(c) * sig(o)
    assert np.allclose(out[0, 0], h, atol=1e-6)


def test_softmax_ce_and_accuracy():
    z = np.array([[1.0, 2.0, 3.0], [1.0, 1.0, 0.0]], np.float32)
    y = np.array([[0, 0, 1], [0, 1, 0]], np.int32)
    loss, dz, per = O.softmax_ce(z, y)
    exp0 = -np.log(np.exp(3) / np.exp([1, 2, 3]).sum())
    exp1 = -np.log(np.exp(1) / np.exp([1, 1, 0]).sum())
    assert abs(loss - (exp0 + exp1) / 2) < 1e-6
    assert O.accuracy(z, y) == 0.5  # row 1: argmax ties resolve to index 0 != 1


def test_truncated_normal_bounds():
    w = O.truncated_normal(np.random.default_rng(1), (20000,), 0.05)
    assert np.abs(w).max() <= 0.1 + 1e-7 and abs(w.std() - 0.05 * 0.8796) < 2e-3


def _small_problem(workflow, fusion, frame_layer="fc7", clips=1, fpc=2, classes=7, hidden=16, layers=1, seed=3):
    params = O.init_params(seed, classes, frame_layer if workflow == "lrcn" else "fc8", hidden, layers,
                           with_lstm=(workflow == "lrcn"))
    rng = np.random.default_rng(seed + 1)
    # unit-range pixels: with the reference's sigma=0.05 init raw +-128 pixels saturate every LSTM gate (SURVEY 7,
    # 'Random-init magnitude'), which makes gradient comparisons vacuous
    frames = rng.uniform(-1, 1, size=(clips * fpc, 227, 227, 3)).astype(np.float32)
    labels = rng.integers(0, classes, size=clips)
    onehot = np.zeros((clips, classes), np.int32)
    onehot[np.arange(clips), labels] = 1
    return params, frames, onehot


@pytest.mark.parametrize("workflow,fusion,layers", [("lrcn", "avg", 1), ("lrcn", "last", 2), ("singleframe", "avg", 1)])
def test_numpy_oracle_matches_torch_autograd(workflow, fusion, layers):
    clips, fpc = 2, 2
    params, frames, onehot = _small_problem(workflow, fusion, clips=clips, fpc=fpc, layers=layers)
    mask = None
    if workflow == "lrcn":
        mask = (np.random.default_rng(9).uniform(size=(clips, 16)) < 0.5).astype(np.float32) * 2.0
    p_np = {k: v.copy() for k, v in params.items()}
    res = O.train_step(p_np, frames, onehot, fpc, lr=1e-3, workflow=workflow, fusion=fusion, clip_norm=10,
                       dropout_mask=mask)
    p_t = T.to_torch(params, requires_grad=True, dtype=torch.float64)
    rt = T.train_step(p_t, torch.tensor(frames, dtype=torch.float64), torch.tensor(onehot), fpc, 1e-3, workflow,
                      fusion, clip_norm=10, dropout_mask=None if mask is None else torch.tensor(mask, dtype=torch.float64))
    lt = rt["logits"].numpy()
    scale = np.abs(lt).max()
    assert np.abs(res["logits"] - lt).max() <= 1e-4 * scale
    assert abs(res["loss"] - rt["loss"]) <= 1e-4 * max(1.0, abs(rt["loss"]))
    assert abs(res["global_norm"] - rt["global_norm"]) <= 2e-3 * rt["global_norm"]
    # clipped gradients (oracle) vs raw torch gradients * scale, per variable, relative to the variable's max
    for k, g in res["grads"].items():
        gt = rt["grads"][k].numpy() * rt["scale"]
        denom = max(np.abs(gt).max(), 1e-12)
        assert np.abs(g - gt).max() / denom < 5e-3, k
    # updated parameters agree
    for k in p_np:
        assert np.abs(p_np[k] - p_t[k].detach().numpy()).max() < 1e-5, k


def test_adam_matches_torch_formula():
    rng = np.random.default_rng(0)
    p = {"w": rng.standard_normal(50).astype(np.float32)}
    st = {}
    ref = p["w"].astype(np.float64).copy()
    m = np.zeros(50)
    v = np.zeros(50)
    for t in range(1, 4):
        g = rng.standard_normal(50).astype(np.float32)
        O.adam_update(p, {"w": g}, st, 0.01)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g.astype(np.float64) ** 2
        ref -= 0.01 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-8)
    assert np.allclose(p["w"], ref, atol=1e-5)


@pytest.mark.parametrize("h,w,k,sh,sw,groups", [(6, 7, 5, 2, 2, 2), (5, 5, 3, 2, 1, 1), (4, 6, 3, 1, 2, 1)])
def test_depth_to_space_formulation_of_the_data_gradient(h, w, k, sh, sw, groups):
    """The algebra behind `vl_pack_dgrad_d2s` / `kernels.conv_dgrad_d2s` (include/vlb200.h): the data gradient of a
    stride-1 SAME convolution equals a stride-(sh, sw) correlation over dy with a (k+sh-1) x (k+sw-1) filter that holds,
    for every sub-position (dy, dx) of an output block, the rotated filter shifted by (dy, dx) - also where h % sh != 0."""
    rng = np.random.default_rng(7)
    cin_g, cout_g = 4, 3
    cin, cout = cin_g * groups, cout_g * groups
    x = rng.standard_normal((2, h, w, cin)).astype(np.float32)
    wt = rng.standard_normal((k, k, cin_g, cout)).astype(np.float32)
    dy = rng.standard_normal((2, h, w, cout)).astype(np.float32)
    dx_ref, _, _ = O.conv2d_same_backward(x, wt, dy, 1, groups)
    pad = (k - 1) // 2            # SAME, stride 1, odd k: pad_top = pad_left = (k-1)/2
    pt = k - 1 - pad              # top / left padding of the walk over dy
    kh2, kw2 = k + sh - 1, k + sw - 1
    py, qx = -(-h // sh), -(-w // sw)
    dyp = np.zeros((2, h + 2 * kh2, w + 2 * kw2, cout), np.float32)   # generous zero border
    dyp[:, kh2:kh2 + h, kw2:kw2 + w] = dy
    got = np.zeros_like(dx_ref)
    for g in range(groups):
        wp = np.zeros((kh2, kw2, cout_g, sh, sw, cin_g), np.float32)  # [ty][tx][k][(dy, dx, c)]
        for dyy in range(sh):
            for dxx in range(sw):
                for ty in range(kh2):
                    for tx in range(kw2):
                        r, q = k - 1 + dyy - ty, k - 1 + dxx - tx
                        if 0 <= r < k and 0 <= q < k:
                            wp[ty, tx, :, dyy, dxx, :] = wt[r, q, :, g * cout_g:(g + 1) * cout_g].T
        for Y in range(py):
            for X in range(qx):
                y0, x0 = kh2 + sh * Y - pt, kw2 + sw * X - pt
                patch = dyp[:, y0:y0 + kh2, x0:x0 + kw2, g * cout_g:(g + 1) * cout_g]  # [n][ty][tx][k]
                blk = np.einsum("nabk,abkyxc->nyxc", patch, wp)
                for dyy in range(sh):
                    for dxx in range(sw):
                        yy, xx = sh * Y + dyy, sw * X + dxx
                        if yy < h and xx < w:
                            got[:, yy, xx, g * cin_g:(g + 1) * cin_g] = blk[:, dyy, dxx]
    assert np.abs(got - dx_ref).max() < 1e-4 * max(1.0, np.abs(dx_ref).max())


def test_caption_oracle_masking_and_feedback_semantics():
    """oracle/caption_numpy.py against the semantics it restates (lstm.py:22-42,102-143,145-265): state tuple (v, v) on
    every layer, dynamic_rnn masking (zero output, state carried through), feedback order of the three visual modes."""
    from oracle import caption_numpy as C
    rng = np.random.default_rng(3)
    b, t_len, d, hd = 3, 4, 5, 8
    kernels = [rng.standard_normal((d + hd, 4 * hd)).astype(np.float32) * 0.3,
               rng.standard_normal((hd + hd, 4 * hd)).astype(np.float32) * 0.3]
    biases = [rng.standard_normal(4 * hd).astype(np.float32) * 0.1 for _ in range(2)]
    x = rng.standard_normal((b, t_len, d)).astype(np.float32)
    init = rng.standard_normal((b, hd)).astype(np.float32)
    out, st = C.evaluate_sequence(x, kernels, biases, [4, 2, 0], init)
    assert not out[1, 2:].any() and not out[2].any() and out[0].all()
    # a zero-length sequence keeps LSTMStateTuple(v, v) on both layers
    assert np.array_equal(st[0][0][2], init[2]) and np.array_equal(st[1][1][2], init[2])
    # a length-2 sequence equals the full evaluation of its first two steps
    out2, st2 = C.evaluate_sequence(x[1:2, :2], kernels, biases, None, init[1:2])
    assert np.allclose(out[1, :2], out2[0]) and np.allclose(st[1][0][1], st2[1][0][0])
    # one step by hand: gates i, j, f, o with forget_bias 1 and c = h = v
    g = x[0, 0] @ kernels[0][:d] + init[0] @ kernels[0][d:] + biases[0]
    i, j, f, o = np.split(g, 4)
    sig = lambda v: 1 / (1 + np.exp(-v))
    c1 = init[0] * sig(f + 1.0) + sig(i) * np.tanh(j)
    h1 = np.tanh(c1) * sig(o)
    out1, _ = C.evaluate_sequence(x[:1, :1], kernels[:1], biases[:1], None, init[:1])
    assert np.allclose(out1[0, 0], h1, atol=1e-6)
    # feedback decode: input_bias consumes the visual vector as step 0 and stores seq_len - 1 words
    vocab, e = 11, d
    emb = rng.standard_normal((vocab, e)).astype(np.float32)
    ow, ob = rng.standard_normal((hd, vocab)).astype(np.float32), rng.standard_normal(vocab).astype(np.float32)
    start = rng.standard_normal(e).astype(np.float32)
    vis = rng.standard_normal((b, e)).astype(np.float32)
    w_bias = C.generate_feedback_sequence(vis, kernels, biases, ow, ob, start, emb, 5, "input_bias")
    assert w_bias.shape == (b * 4,) and w_bias.dtype == np.int64
    kc = [rng.standard_normal((e + e + hd, 4 * hd)).astype(np.float32) * 0.3, kernels[1]]
    w_cat = C.generate_feedback_sequence(vis, kc, biases, ow, ob, start, emb, 5, "input_concat")
    assert w_cat.shape == (b * 5,)
    # items are independent and ordered item-major
    w_one = C.generate_feedback_sequence(vis[1:2], kc, biases, ow, ob, start, emb, 5, "input_concat")
    assert np.array_equal(w_cat[5:10], w_one)
    with pytest.raises(ValueError, match="Undefined rnn visual input mode"):
        C.generate_feedback_sequence(vis, kernels, biases, ow, ob, start, emb, 2, "nope")
