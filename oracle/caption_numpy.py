"""CPU oracle of the captioning LSTM variants -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in plain numpy, what the reference builds with TensorFlow 1.x ops in
  models/lstm/lstm.py:22-42     get_zero_state / get_state_tuple: LSTMStateTuple(v, v) for EVERY layer
  models/lstm/lstm.py:102-143   evaluate_sequence: dynamic_rnn with sequence_length = nonzero_per_sequence and an
                                optional initial state (beyond its length a sequence emits zeros and keeps its state)
  models/lstm/lstm.py:145-265   generate_feedback_sequence: greedy decode, the i-th word's embedding is the (i+1)-th
                                input; visual input as initial state (state_bias), concatenated to every input
                                (input_concat) or as a first extra input (input_bias); get_embedding_from_logits
PARITY UNPINNED like oracle/lrcn_numpy.py (TensorFlow cannot run here; the captioning path is not even wired into the
reference's current Model / Validation: val.py:32 "Not implemented").  Only tests/ may import this module.
"""
import numpy as np

from .lrcn_numpy import F32, _ident, sigmoid

# `q` (None or lrcn_numpy.bf16_round) marks the device path's bf16 storage points: the operands of the tensor-core
# products (cell inputs x and kernel[:d], the output fc, the state fc); the recurrent product h @ kernel[d:] stays fp32.


def _cell(x, c, h, kern, bias, forget_bias=1.0, q=_ident):
    """BasicLSTMCell: gates i, j, f, o = split([x, h] @ kernel + bias)."""
    d = x.shape[1]
    g = (q(x) @ q(kern[:d]) + h @ kern[d:] + bias).astype(F32)
    i, j, f, o = np.split(g, 4, axis=1)
    c_new = (c * sigmoid(f + forget_bias) + sigmoid(i) * np.tanh(j)).astype(F32)
    h_new = (np.tanh(c_new) * sigmoid(o)).astype(F32)
    return c_new, h_new


def evaluate_sequence(x, kernels, biases, lengths=None, init_vec=None, q=None):
    """x [B,T,D] -> (outputs [B,T,H] of the top layer, [(c, h)] final state per layer)."""
    q = q or _ident
    b, t_len, _ = x.shape
    hid = kernels[0].shape[1] // 4
    lengths = np.full(b, t_len, np.int64) if lengths is None else np.asarray(lengths)
    state = []
    for _ in kernels:
        v = np.zeros((b, hid), F32) if init_vec is None else np.asarray(init_vec, F32).copy()
        state.append((v.copy(), v.copy()))  # LSTMStateTuple(c = v, h = v)
    out = np.zeros((b, t_len, hid), F32)
    for t in range(t_len):
        inp = x[:, t, :]
        valid = (t < lengths)[:, None]
        for layer, (kern, bias) in enumerate(zip(kernels, biases)):
            c, h = state[layer]
            c_new, h_new = _cell(inp, c, h, kern, bias, q=q)
            state[layer] = (np.where(valid, c_new, c), np.where(valid, h_new, h))
            inp = np.where(valid, h_new, 0.0).astype(F32)  # what dynamic_rnn hands to the next layer / emits
        out[:, t, :] = inp
    return out, state


def generate_feedback_sequence(visual, kernels, biases, out_w, out_b, start_vector, embedding, seq_len,
                               mode="state_bias", state_fc=None, q=None):
    """Greedy decode (lstm.py:145-265).  Returns the word indices, item-major like index_accumulation:
    [item 0 step 0, item 0 step 1, ..., item 1 step 0, ...]."""
    q = q or _ident
    visual = np.asarray(visual, F32)
    hid = kernels[0].shape[1] // 4
    if mode == "state_bias" and state_fc is not None:
        visual = (q(visual) @ q(state_fc[0]) + state_fc[1]).astype(F32)  # convert_dim_fc(..., "input_state_fc")
    indices = []
    for item in range(visual.shape[0]):
        vec = visual[item:item + 1]
        io = np.asarray(start_vector, F32)[None, :]
        state = [(np.zeros((1, hid), F32), np.zeros((1, hid), F32)) for _ in kernels]
        for i in range(seq_len):
            if mode == "state_bias":
                if i == 0:
                    state = [(vec.copy(), vec.copy()) for _ in kernels]
            elif mode == "input_concat":
                io = np.concatenate([io, vec], axis=1)
            elif mode == "input_bias":
                if i == 0:
                    io = vec
                elif i == 1:
                    io = np.asarray(start_vector, F32)[None, :]
            else:
                raise ValueError("Undefined rnn visual input mode [%s]" % mode)
            inp = io
            for layer, (kern, bias) in enumerate(zip(kernels, biases)):
                c, h = _cell(inp, state[layer][0], state[layer][1], kern, bias, q=q)
                state[layer] = (c, h)
                inp = h
            logits = inp if out_w is None else (q(inp) @ q(out_w) + out_b).astype(F32)
            word = int(np.argmax(logits, axis=1)[0])  # tf.arg_max: lowest index on ties
            io = embedding[word][None, :].astype(F32)
            if not (mode == "input_bias" and i == 0):
                indices.append(word)
    return np.asarray(indices, np.int64)
