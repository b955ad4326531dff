"""Multi-input pipelines on the device: the tensor-list fusions of `Model.build_pipeline` and the LSTM
whose initial state comes from an auxiliary input.

Reference: models/model.py:41-76 (several inputs per pipeline: dataset tags or earlier pipelines), :128-135 (the second
input of an LSTM classifier is its state vector), tf_util.py:99-124 (`vec_seq_concat`), :136-192
(`apply_tensor_list_fusion`: avg / maximum / concat / ibias), :195-206 (`replicate_auxilliary_tensor`),
models/lstm/lstm.py:74-77,127-130 (`input_state_fc`, `get_state_tuple`).

Element-wise fusions run in `vl_fuse_list`; concat / ibias / replication are pure data movement over device buffers
(torch is the allocator / copy shell here, no arithmetic); the state-biased LSTM reuses `captioning.CaptionLSTM`
(vl_gemm + vl_lstm_fwd_ex) and the segmented pooling of the activity-recognition head.
"""
import ctypes

import torch

from . import _native as nv
from .utils import error

F32 = torch.float32
POOL = {"avg": 0, "last": 1, "max": 2}


def replicate_auxilliary_tensor(aux, cpv_main, cpv_aux):
    """tf_util.py:195-206: `reshape(aux, [1, -1])`, `tile([tile_num, 1])`, `reshape([-1, dim_aux])` with tile_num =
    int(cpv_main / cpv_aux) - i.e. the WHOLE aux block is repeated tile_num times (rows a0..am, a0..am, ...), which is
    what the code does (its comment says "in place")."""
    tile_num = int(cpv_main / cpv_aux)
    if tile_num > 1:
        aux = aux.reshape(1, -1).repeat(tile_num, 1).reshape(-1, aux.shape[-1])
    return aux.contiguous()


def vec_seq_concat(seq, vec, sequence_length, order="vecfirst"):
    """tf_util.py:99-124: every row of `vec` is repeated `sequence_length` times (tile along the columns, reshape back to
    one vector per row) and concatenated column-wise with `seq`."""
    vec_dim = vec.shape[-1]
    rep = vec.repeat(1, sequence_length).reshape(-1, vec_dim)
    if rep.shape[0] != seq.shape[0]:
        error("vec_seq_concat: %d sequence rows against %d replicated vector rows" % (seq.shape[0], rep.shape[0]))
    return torch.cat([rep, seq] if order == "vecfirst" else [seq, rep], dim=1).contiguous()


def apply_tensor_list_fusion(inputs, fusion_method, dims, fpcs, cpvs):
    """tf_util.py:136-192.  inputs: list of fp32 device tensors [rows_i, dims_i].  Returns (tensor, dim, fpc, cpv)."""
    cpv_ratio = int(cpvs[0] / cpvs[1]) if len(inputs) == 2 else None
    if fusion_method in ("avg", "maximum"):
        n = inputs[0].numel()
        if any(t.shape != inputs[0].shape for t in inputs):
            error("Input fusion [%s] needs equally shaped tensors, got %s" % (fusion_method, [tuple(t.shape) for t in inputs]))
        ptrs = (ctypes.c_void_p * len(inputs))(*[t.contiguous().data_ptr() for t in inputs])
        out = torch.empty_like(inputs[0])
        nv.call("vl_fuse_list", ptrs, len(inputs), n, POOL["avg" if fusion_method == "avg" else "max"], out, None)
        return out, dims[0], fpcs[0], cpvs[0]
    if fusion_method == "concat":
        if cpv_ratio == 1:
            return torch.cat(list(inputs), dim=1).contiguous(), sum(dims), fpcs[0], cpvs[0]
        aux = replicate_auxilliary_tensor(inputs[1], cpvs[0], cpvs[1])
        return vec_seq_concat(inputs[0], aux, fpcs[0]), sum(dims), fpcs[0], cpvs[0]
    if fusion_method == "ibias":
        main, aux = inputs
        if cpv_ratio != 1:
            aux = replicate_auxilliary_tensor(aux, cpvs[0], cpvs[1])
        mdim, adim = dims
        mfpc = fpcs[0]
        main3 = main.reshape(-1, mfpc, mdim)
        aux3 = aux.reshape(-1, 1, adim)
        combo = torch.cat([aux3, main3], dim=1)  # the aux vector becomes the first element of every sequence
        return combo.reshape(-1, mdim).contiguous(), mdim, mfpc + 1, cpvs[0]
    error("Unknown fusion method: [%s]" % fusion_method)


class StateBiasedLSTMHead(object):
    """LSTM classifier whose second input is its state vector (model.py:128-135, lstm.py:59-99): the aux vector of every
    clip goes through `input_state_fc` (when its width differs from the hidden size) and becomes c AND h of every layer;
    then dynamic_rnn over the clip's frame features, temporal fusion, output fc.  Forward only (validation)."""

    def __init__(self, params, hidden, layers, fusion="avg", device="cuda:0"):
        from .captioning import CaptionLSTM
        if fusion not in ("avg", "last"):
            error("Undefined frame fusion type : %s" % fusion)
        self.net = CaptionLSTM(params, hidden, layers, device)
        self.fusion = fusion

    def forward(self, features, aux, fpc):
        """features fp32 / bf16 [clips * fpc, d] (e.g. Engine.forward_features), aux fp32 [clips, d_aux] ->
        logits fp32 [clips, num_classes]."""
        net = self.net
        clips = features.shape[0] // fpc
        init = net.initial_state(aux)
        out, _ = net.evaluate_sequence(features.reshape(clips, fpc, -1), None, init)
        hd = net.hidden
        fused = torch.empty(clips, hd, dtype=F32, device=net.dev)
        fused_bf = torch.empty(clips, hd, dtype=torch.bfloat16, device=net.dev)
        nv.call("vl_segment_pool_fwd", out.reshape(clips * fpc, hd), None, fpc, clips, hd, POOL[self.fusion], fused, fused_bf)
        return net._project(fused_bf)
