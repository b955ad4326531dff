"""Captioning LSTM variants on the device (BASELINE configs[2]/[3]).

Reference: models/lstm/lstm.py:102-143 (`evaluate_sequence` with `nonzero_per_sequence` lengths and an initial state),
:145-265 (`generate_feedback_sequence`: greedy decode with the visual vector as initial state / concatenated input /
first input, `get_embedding_from_logits`).  The reference's current Model / Validation do not wire this path
(val.py:32 "Not implemented"); what is built here are the device operators with the reference's semantics, on the
kernels of the activity-recognition path:

  input projection of a whole sequence, vocabulary projection   tcgen05 GEMM (vl_gemm), bf16 operands, fp32 out
  recurrence with initial state + per-sequence lengths          vl_lstm_fwd_ex (csrc/lstm.cu)
  argmax over the vocabulary + embedding gather                 vl_argmax_gather (csrc/head.cu)

Variables carry the reference's TF names: `rnn/multi_rnn_cell/cell_<l>/basic_lstm_cell/{kernel,bias}`,
`output_fc_{w,b}` (hidden -> vocabulary, convert_dim_fc), `input_state_fc_{w,b}` (visual -> hidden, state_bias mode).
"""
import math

import numpy as np
import torch

from . import _native as nv
from . import kernels as K

BF16, F32 = torch.bfloat16, torch.float32


def _pad8(n):
    return -(-n // 8) * 8


def caption_variable_shapes(input_dim, hidden, layers, output_dim, visual_dim=None, mode="state_bias"):
    d_in = input_dim + (visual_dim if mode == "input_concat" else 0)
    out = []
    for layer in range(layers):
        out.append(("rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer, (d_in + hidden, 4 * hidden)))
        out.append(("rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % layer, (4 * hidden,)))
        d_in = hidden
    if hidden != output_dim:
        out.append(("output_fc_w", (hidden, output_dim)))
        out.append(("output_fc_b", (output_dim,)))
    if mode == "state_bias" and visual_dim is not None and visual_dim != hidden:
        out.append(("input_state_fc_w", (visual_dim, hidden)))
        out.append(("input_state_fc_b", (hidden,)))
    return out


def init_caption_variables(shapes, seed=1234):
    """BasicLSTMCell defaults (glorot-uniform kernel, zero bias) and convert_dim_fc (truncated normal 0.05, bias 0.1)."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, shape in shapes:
        if name.endswith("basic_lstm_cell/kernel"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            p[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif name.endswith("basic_lstm_cell/bias"):
            p[name] = np.zeros(shape, np.float32)
        elif len(shape) == 1:
            p[name] = np.full(shape, 0.1, np.float32)
        else:
            v = rng.standard_normal(size=shape)
            bad = np.abs(v) > 2.0
            while bad.any():
                v[bad] = rng.standard_normal(size=int(bad.sum()))
                bad = np.abs(v) > 2.0
            p[name] = (v * 0.05).astype(np.float32)
    return p


class CaptionLSTM(object):
    """MultiRNNCell([BasicLSTMCell(hidden)] * layers) + output fc, evaluated on the device."""

    def __init__(self, params, hidden, layers, device="cuda:0"):
        nv.lib()
        if not torch.cuda.is_available():
            raise nv.NativeError("vlb200.captioning needs a CUDA device (sm_100a); no CPU fallback exists")
        self.dev = torch.device(device)
        self.hidden, self.layers = int(hidden), int(layers)
        self.p = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(self.dev) for k, v in params.items()}
        self.wx, self.wh, self.bias = [], [], []
        for layer in range(layers):
            kern = self.p["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % layer]
            d_in = kern.shape[0] - hidden
            wx = torch.zeros(_pad8(d_in), 4 * hidden, dtype=BF16, device=self.dev)  # rows padded: TMA pitch % 8
            wx[:d_in] = kern[:d_in].to(BF16)
            self.wx.append(wx)
            self.wh.append(kern[d_in:].contiguous())
            self.bias.append(self.p["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % layer])
        self.out_w = self.out_b = None
        if "output_fc_w" in self.p:
            w = self.p["output_fc_w"]
            self.vocab = int(w.shape[1])
            self.out_w = torch.zeros(hidden, _pad8(self.vocab), dtype=BF16, device=self.dev)
            self.out_w[:, :self.vocab] = w.to(BF16)
            self.out_b = self.p["output_fc_b"]
        else:
            self.vocab = hidden
        self.state_w = None
        if "input_state_fc_w" in self.p:
            w = self.p["input_state_fc_w"]
            self.state_w = torch.zeros(_pad8(w.shape[0]), hidden, dtype=BF16, device=self.dev)
            self.state_w[:w.shape[0]] = w.to(BF16)

    # -- helpers ------------------------------------------------------------------------------------
    def _bf16_rows(self, x):
        """fp32 / bf16 [rows, d] -> bf16 [rows, pad8(d)] (zero padded columns: TMA row pitch)."""
        rows, d = x.shape
        if x.dtype == BF16 and d % 8 == 0:
            return x
        x = x.contiguous() if x.dtype == F32 else x.float().contiguous()
        out = torch.empty(rows, _pad8(d), dtype=BF16, device=self.dev)
        nv.call("vl_pack_bf16", x, rows, d, out, rows, _pad8(d), rows, rows)
        return out

    def _project(self, h_bf16):
        """convert_dim_fc(h, output_dim, "output_fc"): logits fp32 [rows, vocab]."""
        rows = h_bf16.shape[0]
        if self.out_w is None:
            return h_bf16.float()
        logits = torch.empty(rows, self.out_w.shape[1], dtype=F32, device=self.dev)
        K.linear_fwd(h_bf16, self.out_w, self.out_b, logits, relu=False, n=self.vocab)
        return logits[:, :self.vocab]

    def initial_state(self, visual):
        """state_bias: the visual vector, mapped to the state width by input_state_fc when needed (lstm.py:74-77,
        171-173), becomes c AND h of every layer (get_state_tuple, lstm.py:34-42)."""
        v = torch.as_tensor(visual, dtype=F32, device=self.dev)
        if self.state_w is not None:
            out = torch.empty(v.shape[0], self.hidden, dtype=F32, device=self.dev)
            K.linear_fwd(self._bf16_rows(v), self.state_w, self.p["input_state_fc_b"], out, relu=False)
            v = out
        return v.contiguous()

    # -- evaluate_sequence (lstm.py:102-143) ---------------------------------------------------------
    def evaluate_sequence(self, x, lengths=None, init_vec=None):
        """x [B, T, D] -> (top-layer outputs fp32 [B, T, H] (zero beyond a sequence's length), [(c, h)] per layer)."""
        x = torch.as_tensor(x, dtype=F32, device=self.dev)
        b, t_len, _ = x.shape
        hd = self.hidden
        len_dev = None if lengths is None else torch.as_tensor(np.asarray(lengths, np.int32)).to(self.dev)
        init = None if init_vec is None else torch.as_tensor(init_vec, dtype=F32, device=self.dev).contiguous()
        inp = self._bf16_rows(x.reshape(b * t_len, -1))
        states = []
        hseq = None
        for layer in range(self.layers):
            gx = torch.empty(b * t_len, 4 * hd, dtype=F32, device=self.dev)
            K.linear_fwd(inp, self.wx[layer], self.bias[layer], gx, relu=False)
            hseq = torch.empty(b * t_len, hd, dtype=F32, device=self.dev)
            hseq_bf = torch.empty(b * t_len, hd, dtype=BF16, device=self.dev)
            h_last = torch.empty(b, hd, dtype=F32, device=self.dev)
            c_last = torch.empty(b, hd, dtype=F32, device=self.dev)
            nv.call("vl_lstm_fwd_ex", gx, self.wh[layer], init, init, len_dev, None, None, hseq, hseq_bf, None, h_last,
                    c_last, b, t_len, hd, 1.0)
            states.append((c_last, h_last))
            inp = hseq_bf
        self._top_bf16 = inp
        return hseq.view(b, t_len, hd), states

    def sequence_logits(self, x, lengths=None, init_vec=None):
        """Teacher-forced logits of every timestep: output_fc over the (masked) outputs, fp32 [B, T, vocab]."""
        out, _ = self.evaluate_sequence(x, lengths, init_vec)
        b, t_len, _ = out.shape
        return self._project(self._top_bf16).reshape(b, t_len, self.vocab)

    # -- generate_feedback_sequence (lstm.py:145-265) -------------------------------------------------
    def generate_feedback_sequence(self, visual, start_vector, embedding, seq_len, mode="state_bias"):
        """Greedy decode of `seq_len` steps for every item of the batch AT ONCE (the reference unrolls a Python loop per
        item; items are independent).  Returns int64 word indices in the reference's order (item-major)."""
        emb = torch.as_tensor(embedding, dtype=F32, device=self.dev).contiguous()
        start = torch.as_tensor(start_vector, dtype=F32, device=self.dev)
        vis = torch.as_tensor(visual, dtype=F32, device=self.dev)
        b, hd, e = vis.shape[0], self.hidden, emb.shape[1]
        io = start[None, :].expand(b, -1).contiguous()
        zeros = torch.zeros(b, hd, dtype=F32, device=self.dev)
        if mode == "state_bias":
            v = self.initial_state(vis)
            state = [(v.clone(), v.clone()) for _ in range(self.layers)]
        elif mode in ("input_concat", "input_bias"):
            state = [(zeros.clone(), zeros.clone()) for _ in range(self.layers)]
        else:
            raise ValueError("Undefined rnn visual input mode [%s]" % mode)
        words = []
        idx = torch.empty(b, dtype=torch.int64, device=self.dev)
        nxt = torch.empty(b, e, dtype=F32, device=self.dev)
        for i in range(seq_len):
            if mode == "input_concat":
                step_in = torch.cat([io, vis], dim=1)
            elif mode == "input_bias" and i == 0:
                step_in = vis
            elif mode == "input_bias" and i == 1:
                step_in = start[None, :].expand(b, -1)
            else:
                step_in = io
            inp = self._bf16_rows(step_in.contiguous())
            for layer in range(self.layers):
                gx = torch.empty(b, 4 * hd, dtype=F32, device=self.dev)
                K.linear_fwd(inp, self.wx[layer], self.bias[layer], gx, relu=False)
                c0, h0 = state[layer]
                h_bf = torch.empty(b, hd, dtype=BF16, device=self.dev)
                h_last = torch.empty(b, hd, dtype=F32, device=self.dev)
                c_last = torch.empty(b, hd, dtype=F32, device=self.dev)
                nv.call("vl_lstm_fwd_ex", gx, self.wh[layer], h0, c0, None, None, None, None, h_bf, None, h_last,
                        c_last, b, 1, hd, 1.0)
                state[layer] = (c_last, h_last)
                inp = h_bf
            logits = self._project(inp)
            nv.call("vl_argmax_gather", logits, b, self.vocab, logits.stride(0), emb, e, idx, nxt, None)
            io = nxt.clone()
            if not (mode == "input_bias" and i == 0):
                words.append(idx.clone())
        return torch.stack(words, dim=1).reshape(-1).cpu().numpy()
