"""Independent torch-CPU implementation of the LRCN path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Second opinion for oracle/lrcn_numpy.py (tests/test_oracle.py cross-checks the two) and the multi-threaded
CPU arm timed by `bench.py` (`cpu_baseline`, `--impl reference`): the same graph the reference builds with
TensorFlow ops (models/alexnet/alexnet.py:49-280, models/lstm/lstm.py:59-143, tf_util.py:4-60, train.py:117-222)
expressed with torch.nn.functional ops (oneDNN/MKL kernels on the host cores) and torch autograd.
TensorFlow itself cannot be installed here (see oracle/lrcn_numpy.py header): PARITY UNPINNED.
"""
import numpy as np
import torch
import torch.nn.functional as F


class _RoundBF16(torch.autograd.Function):
    """Storage-precision model (same role as `q=bf16_round` in oracle/lrcn_numpy.py): the value is rounded to
    bfloat16 where the device path stores a tensor in bf16, and so is the gradient that flows back through that
    storage point (the device keeps activation gradients in bf16 as well)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _ident(x):
    return x


def _q(flag):
    return _RoundBF16.apply if flag else _ident


def _same_pad(x_nchw, k, s):
    h, w = x_nchw.shape[2], x_nchw.shape[3]

    def pads(n):
        out = -(-n // s)
        total = max((out - 1) * s + k - n, 0)
        return total // 2, total - total // 2

    (pt, pb), (pl, pr) = pads(h), pads(w)
    return F.pad(x_nchw, (pl, pr, pt, pb))


def _conv(x, w_hwio, b, stride, groups):
    w = w_hwio.permute(3, 2, 0, 1)  # HWIO -> OIHW
    return F.conv2d(_same_pad(x, w_hwio.shape[0], stride), w, b, stride=stride, groups=groups)


def _lrn(x):
    # tf LRN: x / (1 + alpha * sum_{5 ch} x^2)^beta ; torch divides alpha by the window size
    return F.local_response_norm(x, size=5, alpha=2e-05 * 5, beta=0.75, k=1.0)


def alexnet(params, frames_nhwc, final_layer="fc7", q=False):
    """q=True rounds to bf16 at the device path's storage points (operands, activations), like the numpy oracle."""
    P = params
    r = _q(q)
    x = r(frames_nhwc.permute(0, 3, 1, 2))
    x = F.max_pool2d(r(_lrn(r(F.relu(_conv(x, r(P["dcnn/conv1W"]), P["dcnn/conv1b"], 4, 1))))), 3, 2)
    x = F.max_pool2d(r(_lrn(r(F.relu(_conv(x, r(P["dcnn/conv2W"]), P["dcnn/conv2b"], 1, 2))))), 3, 2)
    x = r(F.relu(_conv(x, r(P["dcnn/conv3W"]), P["dcnn/conv3b"], 1, 1)))
    x = r(F.relu(_conv(x, r(P["dcnn/conv4W"]), P["dcnn/conv4b"], 1, 2)))
    x = r(F.relu(_conv(x, r(P["dcnn/conv5W"]), P["dcnn/conv5b"], 1, 2)))
    x = F.max_pool2d(x, 3, 2)
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)  # HWC-major flatten
    out = r(F.relu(flat @ r(P["dcnn/fc6W"]) + P["dcnn/fc6b"]))
    if final_layer != "fc6":
        out = r(F.relu(out @ r(P["dcnn/fc7W"]) + P["dcnn/fc7b"]))
        if final_layer != "fc7":
            out = out @ r(P["dcnn/fc8W"]) + P["dcnn/fc8b"]
    return out


def lstm(params, seq, forget_bias=1.0, q=False, initial_state=None, return_state=False):
    """seq [B,T,D]; BasicLSTMCell gate order i, j, f, o.  initial_state: None or a list of (c, h) per layer
    (lstm.py:127-130); return_state: also return the final (c, h) of every layer (dynamic_rnn's second output)."""
    r = _q(q)
    layer = 0
    inp = seq
    states = []
    while "rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer in params:
        kern = params["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer]
        bias = params["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % layer]
        hd = kern.shape[1] // 4
        if initial_state is not None:
            c, h = initial_state[layer]
        else:
            h = torch.zeros(inp.shape[0], hd, dtype=inp.dtype)
            c = torch.zeros_like(h)
        outs = []
        # input projection for all timesteps at once (same math as the per-step concat matmul)
        gx = r(inp) @ r(kern[:inp.shape[2]]) + bias
        wh = kern[inp.shape[2]:]
        for t in range(inp.shape[1]):
            g = gx[:, t] + h @ wh
            i, j, f, o = g.chunk(4, dim=1)
            c = c * torch.sigmoid(f + forget_bias) + torch.sigmoid(i) * torch.tanh(j)
            h = torch.tanh(c) * torch.sigmoid(o)
            outs.append(h)
        inp = torch.stack(outs, dim=1)
        states.append((c, h))
        layer += 1
    if return_state:
        return inp, states
    return inp


def fuse(x, method):
    if method == "last":
        return x[:, -1]
    if method == "avg":
        return x.mean(dim=1)
    raise ValueError("Undefined frame fusion type : %s" % method)


def logits_fn(params, frames, fpc, workflow="lrcn", fusion="avg", frame_encoding_layer="fc7", dropout_mask=None,
              q=False):
    r = _q(q)
    if workflow == "lrcn":
        feats = alexnet(params, frames, frame_encoding_layer, q)
        seq = feats.reshape(-1, fpc, feats.shape[1])
        if fusion == "state":
            # lstm.py:81,91-93 + model.py:137-143: no temporal fusion / dropout / output_fc; the logits are the final
            # hidden state h of the top layer, mapped by convert_dim_fc under its default name "fc_convert"
            _, states = lstm(params, seq, q=q, return_state=True)
            out = states[-1][1]
            if "fc_convert_w" in params:
                out = r(out) @ r(params["fc_convert_w"]) + params["fc_convert_b"]
            return out
        out = fuse(lstm(params, seq, q=q), fusion)
        if dropout_mask is not None:
            out = out * dropout_mask
        if "output_fc_w" in params:
            out = r(out) @ r(params["output_fc_w"]) + params["output_fc_b"]
        return out
    if workflow == "fc":
        # classifier fc (model.py:103-125,149-151): optional early fusion of the frame features, convert_dim_fc
        # ("fc_convert") when the feature width differs from the class count, optional late fusion of the logits
        early, late = fusion if isinstance(fusion, tuple) else (None, fusion)
        feats = alexnet(params, frames, frame_encoding_layer, q)
        if early is not None and fpc > 1:
            feats = fuse(feats.reshape(-1, fpc, feats.shape[1]), early)
        if "fc_convert_w" in params:
            feats = r(feats) @ r(params["fc_convert_w"]) + params["fc_convert_b"]
        if late is not None and early is None and fpc > 1:
            feats = fuse(feats.reshape(-1, fpc, feats.shape[1]), late)
        return feats
    fl = alexnet(params, frames, "fc8", q)
    return fuse(fl.reshape(-1, fpc, fl.shape[1]), fusion)


def loss_fn(logits, onehot):
    y = onehot.to(logits.dtype)
    return -(y * F.log_softmax(logits, dim=1)).sum(dim=1).mean()


def to_torch(params_np, requires_grad=False, dtype=torch.float32):
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad) for k, v in params_np.items()}


def train_step(params, frames, onehot, fpc, lr, workflow="lrcn", fusion="avg", frame_encoding_layer="fc7",
               clip_norm=None, dropout_mask=None, q=False):
    """SGD step with torch autograd; `params` is a dict of leaf tensors (requires_grad) updated in place."""
    for p in params.values():
        p.grad = None
    logits = logits_fn(params, frames, fpc, workflow, fusion, frame_encoding_layer, dropout_mask, q)
    loss = loss_fn(logits, onehot)
    loss.backward()
    with torch.no_grad():
        grads = {k: p.grad for k, p in params.items() if p.grad is not None}
        gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
        scale = 1.0
        if clip_norm:
            scale = clip_norm / max(gn, clip_norm)
        for k, g in grads.items():
            params[k] -= lr * scale * g
    acc = (logits.argmax(dim=1) == onehot.argmax(dim=1)).float().mean().item()
    return dict(loss=loss.item(), accuracy=acc, global_norm=gn, logits=logits.detach(), grads=grads, scale=scale)
