"""CPU oracle for the LRCN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-numpy restatement of the arithmetic the reference delegates to TensorFlow 1.x ops, forward and
backward.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product path (video-learning-tf_b200/) never does.

PARITY UNPINNED: the arithmetic lives in TensorFlow 1.x (un-vendored, un-pinned: the reference's
dependencies.txt:1-4 does not even list it) which cannot be installed here (Python 3.12, no network), and
the reference ships no tests or golden vectors.  The restatement follows the published TF-1.x op
semantics at the reference's call sites, is cross-checked against an independent torch-CPU
implementation (oracle/lrcn_torch.py, tests/test_oracle.py) and against hand-computed micro cases.

Reference call sites restated here (file:line in /root/reference):
  models/alexnet/alexnet.py:15-31   dcnn.conv      -> conv2d_same (+groups via split/concat on axis 3)
  models/alexnet/alexnet.py:40-46   make_w_b       -> truncated_normal(0.05), bias 0.1
  models/alexnet/alexnet.py:60-280  dcnn.create    -> alexnet_forward / alexnet_backward
  models/lstm/lstm.py:9-20,102-143  make_cell / evaluate_sequence -> lstm_forward / lstm_backward
  models/lstm/lstm.py:50-56,59-99   apply_dropout / forward_pass_sequence
  tf_util.py:4-30,32-60,126-133     temporal fusion, convert_dim_fc, aggregate_clip_vectors
  train.py:117-124,142-149,199-222  CE loss, accuracy, clip_by_global_norm + SGD/Adam
"""
import numpy as np

F32 = np.float32


def bf16_round(x):
    """Round-to-nearest-even to bfloat16 precision, returned as float32 (what the device path stores)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def _ident(x):
    return x


# Storage-precision model.  Every function below that takes `q` applies it exactly where the device path
# stores a tensor in bf16 (GEMM operands, activations and activation gradients; accumulation, biases, LSTM
# state and weight gradients stay fp32).  q=None is the plain fp32 restatement of the TensorFlow graph (the
# numbers the north_star tolerance for logits/losses refers to); q=bf16_round makes hard decisions (ReLU masks,
# max-pool argmax) coincide with the device path so gradients can be compared tightly.

LRN_RADIUS, LRN_ALPHA, LRN_BETA, LRN_BIAS = 2, 2e-05, 0.75, 1.0  # alexnet.py:80-89,121-130


# --------------------------------------------------------------------------------------------------
# initialisation (alexnet.py:40-46, tf_util.py:44-49, BasicLSTMCell defaults)
# --------------------------------------------------------------------------------------------------
def truncated_normal(rng, shape, stddev):
    """tf.truncated_normal: N(0, stddev^2) with values beyond 2 stddev re-drawn."""
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(F32)


ALEXNET_SHAPES = [  # name, filter shape HWIO (I = Cin/groups), stride, groups
    ("conv1", (11, 11, 3, 96), 4, 1),
    ("conv2", (5, 5, 48, 256), 1, 2),
    ("conv3", (3, 3, 256, 384), 1, 1),
    ("conv4", (3, 3, 192, 384), 1, 2),
    ("conv5", (3, 3, 192, 256), 1, 2),
]


def init_params(seed=1234, num_classes=101, frame_encoding_layer="fc7", lstm_hidden=256, lstm_layers=1,
                with_lstm=True):
    """Random-init parameters keyed by the TF variable names of the reference."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, shp, _, _ in ALEXNET_SHAPES:
        p["dcnn/%sW" % name] = truncated_normal(rng, shp, 0.05)
        p["dcnn/%sb" % name] = np.full((shp[3],), 0.1, F32)
    p["dcnn/fc6W"] = truncated_normal(rng, (9216, 4096), 0.05)
    p["dcnn/fc6b"] = np.full((4096,), 0.1, F32)
    feat = 4096
    if frame_encoding_layer != "fc6":
        p["dcnn/fc7W"] = truncated_normal(rng, (4096, 4096), 0.05)
        p["dcnn/fc7b"] = np.full((4096,), 0.1, F32)
        if frame_encoding_layer != "fc7":  # any other value -> fc8 logits (alexnet.py:273-275)
            p["dcnn/fc8W"] = truncated_normal(rng, (4096, num_classes), 0.05)
            p["dcnn/fc8b"] = np.full((num_classes,), 0.1, F32)
            feat = num_classes
    if with_lstm:
        d_in = feat
        for layer in range(lstm_layers):
            fan_in, fan_out = d_in + lstm_hidden, 4 * lstm_hidden
            lim = np.sqrt(6.0 / (fan_in + fan_out))  # glorot_uniform, the tf.get_variable default
            p["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer] = rng.uniform(
                -lim, lim, size=(fan_in, fan_out)).astype(F32)
            p["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % layer] = np.zeros((fan_out,), F32)
            d_in = lstm_hidden
        if lstm_hidden != num_classes:  # convert_dim_fc is the identity when dims agree (tf_util.py:39-40)
            p["output_fc_w"] = truncated_normal(rng, (lstm_hidden, num_classes), 0.05)
            p["output_fc_b"] = np.full((num_classes,), 0.1, F32)
    return p


# --------------------------------------------------------------------------------------------------
# TF op restatements
# --------------------------------------------------------------------------------------------------
def same_pad(in_size, k, s):
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return out, total // 2, total - total // 2


def _im2col(x, kh, kw, s, pads):
    """x [N,H,W,C] -> col [N,P,Q,kh*kw*C] (tap-major, channel-minor), zero padded."""
    (pt, pb), (pl, pr) = pads
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    n, hp, wp, c = xp.shape
    p = (hp - kh) // s + 1
    q = (wp - kw) // s + 1
    st = xp.strides
    win = np.lib.stride_tricks.as_strided(
        xp, shape=(n, p, q, kh, kw, c), strides=(st[0], st[1] * s, st[2] * s, st[1], st[2], st[3]), writeable=False)
    return np.ascontiguousarray(win).reshape(n, p, q, kh * kw * c), (n, hp, wp, c)


def conv2d_same(x, w, b, stride, groups):
    """tf.nn.conv2d(NHWC, HWIO, SAME) per group + concat + bias_add (alexnet.py:21-31)."""
    kh, kw, cig, co = w.shape
    n, h, wd, c = x.shape
    assert c == cig * groups
    p, pt, pb = same_pad(h, kh, stride)
    q, pl, pr = same_pad(wd, kw, stride)
    cog = co // groups
    outs = []
    for g in range(groups):
        col, _ = _im2col(x[..., g * cig:(g + 1) * cig], kh, kw, stride, ((pt, pb), (pl, pr)))
        wg = w[..., g * cog:(g + 1) * cog].reshape(kh * kw * cig, cog)
        outs.append(col.reshape(-1, kh * kw * cig) @ wg)
    y = np.concatenate(outs, axis=1).reshape(n, p, q, co) + b
    return y.astype(F32)


def conv2d_same_backward(x, w, dy, stride, groups, need_dx=True):
    kh, kw, cig, co = w.shape
    n, h, wd, c = x.shape
    p, pt, pb = same_pad(h, kh, stride)
    q, pl, pr = same_pad(wd, kw, stride)
    cog = co // groups
    dw = np.zeros_like(w)
    dx = np.zeros_like(x) if need_dx else None
    for g in range(groups):
        col, (n_, hp, wp, _) = _im2col(x[..., g * cig:(g + 1) * cig], kh, kw, stride, ((pt, pb), (pl, pr)))
        col2 = col.reshape(-1, kh * kw * cig)
        dyg = dy[..., g * cog:(g + 1) * cog].reshape(-1, cog)
        dw[..., g * cog:(g + 1) * cog] = (col2.T @ dyg).reshape(kh, kw, cig, cog)
        if need_dx:
            wg = w[..., g * cog:(g + 1) * cog].reshape(kh * kw * cig, cog)
            dcol = (dyg @ wg.T).reshape(n, p, q, kh, kw, cig)
            dxp = np.zeros((n, hp, wp, cig), F32)
            for r in range(kh):
                for s_ in range(kw):
                    dxp[:, r:r + (p - 1) * stride + 1:stride, s_:s_ + (q - 1) * stride + 1:stride, :] += dcol[:, :, :, r, s_, :]
            dx[..., g * cig:(g + 1) * cig] = dxp[:, pt:pt + h, pl:pl + wd, :]
    db = dy.reshape(-1, co).sum(axis=0)
    return dx, dw.astype(F32), db.astype(F32)


def relu(x):
    return np.maximum(x, 0).astype(F32)


def lrn_scale(x):
    """bias + alpha * sum_{j in [d-r, d+r]} x_j^2 over channels (tf.nn.local_response_normalization)."""
    sq = (x * x).astype(F32)
    d = x.shape[-1]
    acc = np.zeros_like(sq)
    for off in range(-LRN_RADIUS, LRN_RADIUS + 1):
        lo, hi = max(0, -off), min(d, d - off)
        acc[..., lo:hi] += sq[..., lo + off:hi + off]
    return (LRN_BIAS + LRN_ALPHA * acc).astype(F32)


def lrn(x):
    return (x * np.power(lrn_scale(x), -LRN_BETA)).astype(F32)


def lrn_backward(x, dy):
    s = lrn_scale(x)
    y = x * np.power(s, -LRN_BETA)
    t = (dy * y / s).astype(F32)  # dy_d * x_d * s_d^(-beta-1)
    d = x.shape[-1]
    acc = np.zeros_like(t)
    for off in range(-LRN_RADIUS, LRN_RADIUS + 1):
        lo, hi = max(0, -off), min(d, d - off)
        acc[..., lo:hi] += t[..., lo + off:hi + off]
    return (dy * np.power(s, -LRN_BETA) - 2.0 * LRN_ALPHA * LRN_BETA * x * acc).astype(F32)


def maxpool_3x3s2(x):
    """tf.nn.max_pool(ksize 3, stride 2, VALID); also returns the window-local argmax (first max wins)."""
    n, h, w, c = x.shape
    p, q = (h - 3) // 2 + 1, (w - 3) // 2 + 1
    st = x.strides
    win = np.lib.stride_tricks.as_strided(x, shape=(n, p, q, 3, 3, c),
                                          strides=(st[0], st[1] * 2, st[2] * 2, st[1], st[2], st[3]), writeable=False)
    flat = win.reshape(n, p, q, 9, c)
    arg = flat.argmax(axis=3)  # first occurrence on ties, (h, w) scan order like TF's MaxPoolGrad
    y = np.take_along_axis(flat, arg[:, :, :, None, :], axis=3)[:, :, :, 0, :]
    return y.astype(F32), arg


def maxpool_3x3s2_backward(x_shape, arg, dy):
    n, h, w, c = x_shape
    p, q = dy.shape[1], dy.shape[2]
    dx = np.zeros(x_shape, F32)
    for t in range(9):
        r, s = divmod(t, 3)
        dx[:, r:r + 2 * (p - 1) + 1:2, s:s + 2 * (q - 1) + 1:2, :] += dy * (arg == t)
    return dx


def sigmoid(x):
    # overflow-free form (the benchmark's input range saturates the gates: |x| reaches 1e3 and beyond)
    x = np.asarray(x, dtype=F32)
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e)).astype(F32)


def lstm_forward(x, kernels, biases, forget_bias=1.0, q=None):
    """MultiRNNCell([BasicLSTMCell]) under dynamic_rnn, zero initial state, full-length sequences.

    x [B,T,D]; gate order i, j, f, o (BasicLSTMCell); returns top-layer outputs [B,T,H] and the cache.
    [x_t, h] @ kernel is evaluated as x_t @ kernel[:D] + h @ kernel[D:] (same sum; the device path runs the first
    term as one tensor-core GEMM over all t with bf16 operands and keeps the recurrent term in fp32)."""
    q = q or _ident
    b, t_len, _ = x.shape
    caches = []
    inp = x
    for kern, bias in zip(kernels, biases):
        d_in = inp.shape[2]
        hdim = kern.shape[1] // 4
        xin = q(inp)
        wx, wh = q(kern[:d_in]), kern[d_in:]
        h = np.zeros((b, hdim), F32)
        c = np.zeros((b, hdim), F32)
        outs = np.zeros((b, t_len, hdim), F32)
        steps = []
        for t in range(t_len):
            g = (xin[:, t, :] @ wx + h @ wh + bias).astype(F32)
            i, j, f, o = np.split(g, 4, axis=1)
            si, sf, so, tj = sigmoid(i), sigmoid(f + forget_bias), sigmoid(o), np.tanh(j).astype(F32)
            c_new = (c * sf + si * tj).astype(F32)
            tc = np.tanh(c_new).astype(F32)
            h_new = (tc * so).astype(F32)
            steps.append((xin[:, t, :], h, si, sf, so, tj, c, tc))
            h, c = h_new, c_new
            outs[:, t, :] = h
        caches.append((inp.shape, steps, kern))
        inp = outs
    return inp, caches


def lstm_backward(caches, dout, q=None):
    """BPTT through lstm_forward; dout [B,T,H] is the gradient w.r.t. the top-layer outputs."""
    q = q or _ident
    dkernels, dbiases = [], []
    for inp_shape, steps, kern in reversed(caches):
        b, t_len, d_in = inp_shape
        hdim = kern.shape[1] // 4
        wx, wh = q(kern[:d_in]), kern[d_in:]
        dk = np.zeros_like(kern)
        db = np.zeros((4 * hdim,), F32)
        dinp = np.zeros(inp_shape, F32)
        dh_next = np.zeros((b, hdim), F32)
        dc_next = np.zeros((b, hdim), F32)
        for t in reversed(range(t_len)):
            x_t, h_prev, si, sf, so, tj, c_prev, tc = steps[t]
            dh = dout[:, t, :] + dh_next
            do = dh * tc * so * (1 - so)
            dc = dh * so * (1 - tc * tc) + dc_next
            di = dc * tj * si * (1 - si)
            dj = dc * si * (1 - tj * tj)
            df = dc * c_prev * sf * (1 - sf)
            dc_next = dc * sf
            dg = np.concatenate([di, dj, df, do], axis=1).astype(F32)
            dgq = q(dg)
            dk[:d_in] += x_t.T @ dgq
            dk[d_in:] += q(h_prev).T @ dgq
            db += dgq.sum(axis=0)
            dinp[:, t, :] = dgq @ wx.T
            dh_next = dg @ wh.T
        dkernels.append(dk)
        dbiases.append(db)
        dout = dinp
    return dout, list(reversed(dkernels)), list(reversed(dbiases))


def temporal_fusion(x, method):
    """apply_temporal_fusion (tf_util.py:4-30): x [B,T,D] -> [B,D]."""
    if method == "last":
        return x[:, -1, :].copy()
    if method == "avg":
        return x.mean(axis=1, dtype=F32).astype(F32)
    raise ValueError("Undefined frame fusion type : %s" % method)


def temporal_fusion_backward(shape, method, dy):
    b, t, d = shape
    dx = np.zeros(shape, F32)
    if method == "last":
        dx[:, -1, :] = dy
    else:
        dx[:] = (dy / F32(t))[:, None, :]
    return dx


def softmax_ce(logits, onehot):
    """mean_n( tf.nn.softmax_cross_entropy_with_logits ) and its gradient w.r.t. logits (train.py:121-123)."""
    z = logits - logits.max(axis=1, keepdims=True)
    lse = np.log(np.exp(z).sum(axis=1, keepdims=True))
    logp = z - lse
    y = onehot.astype(F32)
    per = -(y * logp).sum(axis=1)
    loss = per.mean(dtype=F32)
    n = logits.shape[0]
    dlogits = (np.exp(logp) * y.sum(axis=1, keepdims=True) - y) / F32(n)
    return F32(loss), dlogits.astype(F32), per.astype(F32)


def accuracy(logits, onehot):
    """mean(argmax(logits,1) == argmax(labels,1)); lowest index wins ties (train.py:145-147)."""
    return F32((logits.argmax(axis=1) == onehot.argmax(axis=1)).mean())


# --------------------------------------------------------------------------------------------------
# AlexNet encoder (alexnet.py:49-280)
# --------------------------------------------------------------------------------------------------
def alexnet_forward(params, frames, final_layer="fc7", keep_cache=False, q=None):
    """frames [N,227,227,3] fp32 (BGR, mean subtracted) -> features [N,4096] (fc6/fc7) or fc8 logits [N,C]."""
    q = q or _ident
    c = {}
    P = params
    x0 = q(frames.astype(F32))
    a1 = q(relu(conv2d_same(x0, q(P["dcnn/conv1W"]), P["dcnn/conv1b"], 4, 1)))
    n1 = q(lrn(a1))
    p1, arg1 = maxpool_3x3s2(n1)
    a2 = q(relu(conv2d_same(p1, q(P["dcnn/conv2W"]), P["dcnn/conv2b"], 1, 2)))
    n2 = q(lrn(a2))
    p2, arg2 = maxpool_3x3s2(n2)
    a3 = q(relu(conv2d_same(p2, q(P["dcnn/conv3W"]), P["dcnn/conv3b"], 1, 1)))
    a4 = q(relu(conv2d_same(a3, q(P["dcnn/conv4W"]), P["dcnn/conv4b"], 1, 2)))
    a5 = q(relu(conv2d_same(a4, q(P["dcnn/conv5W"]), P["dcnn/conv5b"], 1, 2)))
    p5, arg5 = maxpool_3x3s2(a5)
    flat = p5.reshape(p5.shape[0], -1)  # HWC-major flatten (alexnet.py:228)
    f6 = q(relu(flat @ q(P["dcnn/fc6W"]) + P["dcnn/fc6b"]))
    out = f6
    f7 = None
    if final_layer != "fc6":
        f7 = q(relu(f6 @ q(P["dcnn/fc7W"]) + P["dcnn/fc7b"]))
        out = f7
        if final_layer != "fc7":
            out = (f7 @ q(P["dcnn/fc8W"]) + P["dcnn/fc8b"]).astype(F32)
    if keep_cache:
        c.update(x0=x0, a1=a1, n1=n1, p1=p1, arg1=arg1, a2=a2, n2=n2, p2=p2, arg2=arg2, a3=a3, a4=a4, a5=a5, p5=p5,
                 arg5=arg5, flat=flat, f6=f6, f7=f7, final_layer=final_layer)
    return out.astype(F32), c


def alexnet_backward(params, c, dout, q=None):
    """Gradients of every dcnn/* variable given d(loss)/d(output of alexnet_forward) (post-ReLU for fc6/fc7)."""
    q = q or _ident
    P = params
    g = {}
    fl = c["final_layer"]
    if fl == "fc6":
        df6 = q(dout * (c["f6"] > 0))
    else:
        if fl != "fc7":
            dout = q(dout)
            g["dcnn/fc8W"] = c["f7"].T @ dout
            g["dcnn/fc8b"] = dout.sum(axis=0)
            df7 = q((dout @ q(P["dcnn/fc8W"]).T) * (c["f7"] > 0))
        else:
            df7 = q(dout * (c["f7"] > 0))
        g["dcnn/fc7W"] = c["f6"].T @ df7
        g["dcnn/fc7b"] = df7.sum(axis=0)
        df6 = q((df7 @ q(P["dcnn/fc7W"]).T) * (c["f6"] > 0))
    g["dcnn/fc6W"] = c["flat"].T @ df6
    g["dcnn/fc6b"] = df6.sum(axis=0)
    dp5 = q(df6 @ q(P["dcnn/fc6W"]).T).reshape(c["p5"].shape)
    da5 = q(maxpool_3x3s2_backward(c["a5"].shape, c["arg5"], dp5) * (c["a5"] > 0))
    da4, g["dcnn/conv5W"], g["dcnn/conv5b"] = conv2d_same_backward(c["a4"], q(P["dcnn/conv5W"]), da5, 1, 2)
    da4 = q(da4 * (c["a4"] > 0))
    da3, g["dcnn/conv4W"], g["dcnn/conv4b"] = conv2d_same_backward(c["a3"], q(P["dcnn/conv4W"]), da4, 1, 2)
    da3 = q(da3 * (c["a3"] > 0))
    dp2, g["dcnn/conv3W"], g["dcnn/conv3b"] = conv2d_same_backward(c["p2"], q(P["dcnn/conv3W"]), da3, 1, 1)
    dn2 = q(maxpool_3x3s2_backward(c["n2"].shape, c["arg2"], q(dp2)))
    da2 = q(lrn_backward(c["a2"], dn2) * (c["a2"] > 0))
    dp1, g["dcnn/conv2W"], g["dcnn/conv2b"] = conv2d_same_backward(c["p1"], q(P["dcnn/conv2W"]), da2, 1, 2)
    dn1 = q(maxpool_3x3s2_backward(c["n1"].shape, c["arg1"], q(dp1)))
    da1 = q(lrn_backward(c["a1"], dn1) * (c["a1"] > 0))
    _, g["dcnn/conv1W"], g["dcnn/conv1b"] = conv2d_same_backward(c["x0"], q(P["dcnn/conv1W"]), da1, 4, 1, need_dx=False)
    return {k: v.astype(F32) for k, v in g.items()}


# --------------------------------------------------------------------------------------------------
# Whole pipelines (models/model.py:84-155)
# --------------------------------------------------------------------------------------------------
def _lstm_vars(params):
    kernels, biases = [], []
    layer = 0
    while "rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer in params:
        kernels.append(params["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer])
        biases.append(params["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % layer])
        layer += 1
    return kernels, biases


def lrcn_forward(params, frames, fpc, fusion="avg", frame_encoding_layer="fc7", dropout_mask=None,
                 keep_cache=False, q=None):
    """LRCN: dcnn(fc6|fc7) -> LSTM -> temporal fusion -> dropout -> output_fc  => logits [clips, C].

    dropout_mask: None (validation / keep_prob <= 0) or an array [clips,H] already scaled by 1/keep_prob."""
    qq = q or _ident
    feats, ac = alexnet_forward(params, frames, frame_encoding_layer, keep_cache, q)
    seq = feats.reshape(-1, fpc, feats.shape[1])  # lstm.py:120
    kernels, biases = _lstm_vars(params)
    outs, lc = lstm_forward(seq, kernels, biases, q=q)
    fused = temporal_fusion(outs, fusion)
    dropped = fused if dropout_mask is None else (fused * dropout_mask).astype(F32)
    if "output_fc_w" in params:
        logits = (qq(dropped) @ qq(params["output_fc_w"]) + params["output_fc_b"]).astype(F32)
    else:
        logits = dropped
    cache = dict(ac=ac, lc=lc, outs_shape=outs.shape, dropped=dropped, fusion=fusion, mask=dropout_mask,
                 feat_shape=feats.shape) if keep_cache else None
    return logits, cache


def lrcn_backward(params, cache, dlogits, q=None):
    qq = q or _ident
    g = {}
    if "output_fc_w" in params:
        dl = qq(dlogits)
        g["output_fc_w"] = qq(cache["dropped"]).T @ dl
        g["output_fc_b"] = dl.sum(axis=0)
        dd = dl @ qq(params["output_fc_w"]).T
    else:
        dd = dlogits
    if cache["mask"] is not None:
        dd = dd * cache["mask"]
    douts = temporal_fusion_backward(cache["outs_shape"], cache["fusion"], dd.astype(F32))
    dseq, dks, dbs = lstm_backward(cache["lc"], douts, q=q)
    for layer, (dk, db) in enumerate(zip(dks, dbs)):
        g["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer] = dk
        g["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % layer] = db
    dfeat = dseq.reshape(cache["feat_shape"])
    g.update(alexnet_backward(params, cache["ac"], dfeat, q))
    return {k: v.astype(F32) for k, v in g.items()}


def singleframe_forward(params, frames, fpc, fusion="avg", keep_cache=False, q=None):
    """Single-frame workflow: dcnn -> fc8 logits per frame -> late fusion over fpc (model.py:149-151)."""
    fl, ac = alexnet_forward(params, frames, "fc8", keep_cache, q)
    seq = fl.reshape(-1, fpc, fl.shape[1])
    logits = temporal_fusion(seq, fusion)
    cache = dict(ac=ac, seq_shape=seq.shape, fusion=fusion) if keep_cache else None
    return logits, cache


def singleframe_backward(params, cache, dlogits, q=None):
    dseq = temporal_fusion_backward(cache["seq_shape"], cache["fusion"], dlogits)
    return alexnet_backward(params, cache["ac"], dseq.reshape(-1, dseq.shape[2]), q)


# --------------------------------------------------------------------------------------------------
# optimiser (train.py:199-222)
# --------------------------------------------------------------------------------------------------
def global_norm(grads):
    return F32(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values())))


def clip_by_global_norm(grads, clip_norm):
    gn = global_norm(grads)
    scale = F32(clip_norm) / max(gn, F32(clip_norm))
    return {k: (v * scale).astype(F32) for k, v in grads.items()}, gn


def mean_grad_norm(grads):
    """grads_norm summary: mean over variables of ||g_i||_2 (train.py:219-222)."""
    return F32(np.mean([np.sqrt(float((g.astype(np.float64) ** 2).sum())) for g in grads.values()]))


def sgd_update(params, grads, lr):
    for k, g in grads.items():
        params[k] = (params[k] - F32(lr) * g).astype(F32)


def adam_update(params, grads, state, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t * m / (sqrt(v) + eps)."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    for k, g in grads.items():
        m = state.setdefault("m/" + k, np.zeros_like(g))
        v = state.setdefault("v/" + k, np.zeros_like(g))
        m[:] = beta1 * m + (1 - beta1) * g
        v[:] = beta2 * v + (1 - beta2) * g * g
        params[k] = (params[k] - F32(lr_t) * m / (np.sqrt(v) + F32(eps))).astype(F32)


def train_step(params, frames, onehot, fpc, lr, workflow="lrcn", fusion="avg", frame_encoding_layer="fc7",
               clip_norm=None, optimizer="sgd", opt_state=None, dropout_mask=None, q=None):
    """One `sess.run([loss, lr, global_step, optimizer])` (run_task.py:44) on the CPU.

    Returns dict(loss, accuracy, grads_norm, global_norm, logits, grads); updates `params` in place."""
    if workflow == "lrcn":
        logits, cache = lrcn_forward(params, frames, fpc, fusion, frame_encoding_layer, dropout_mask, True, q)
    else:
        logits, cache = singleframe_forward(params, frames, fpc, fusion, True, q)
    loss, dlogits, _ = softmax_ce(logits, onehot)
    acc = accuracy(logits, onehot)
    if workflow == "lrcn":
        grads = lrcn_backward(params, cache, dlogits, q)
    else:
        grads = singleframe_backward(params, cache, dlogits, q)
    gn = global_norm(grads)
    if clip_norm:
        grads, gn = clip_by_global_norm(grads, clip_norm)
    if optimizer == "sgd":
        sgd_update(params, grads, lr)
    elif optimizer == "adam":
        adam_update(params, grads, opt_state if opt_state is not None else {}, lr)
    else:
        raise ValueError("Undefined optimizer %s" % optimizer)
    return dict(loss=loss, accuracy=acc, grads_norm=mean_grad_norm(grads), global_norm=gn, logits=logits, grads=grads)
