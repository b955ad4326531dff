"""Multi-input pipelines: tensor-list fusions and the state-biased LSTM head against numpy restatements
of tf_util.py:99-124,136-206 and lstm.py:59-99 (tile / reshape / concat have identical semantics in numpy)."""
import numpy as np
import pytest
import torch

from oracle import caption_numpy as C
from oracle import lrcn_numpy as O

pytestmark = pytest.mark.gpu


def _np_replicate(aux, cpv_main, cpv_aux):  # tf_util.py:195-206
    t = int(cpv_main / cpv_aux)
    return np.tile(aux.reshape(1, -1), (t, 1)).reshape(-1, aux.shape[-1]) if t > 1 else aux


def _np_vec_seq_concat(seq, vec, sl):  # tf_util.py:99-124
    return np.concatenate([np.tile(vec, (1, sl)).reshape(-1, vec.shape[-1]), seq], axis=1)


def test_tensor_list_fusions_bit_exact():
    import vlb200  # noqa: F401
    from vlb200 import fusion as F
    rng = np.random.default_rng(61)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    a, b, c = (rng.standard_normal((12, 40)).astype(np.float32) for _ in range(3))
    out, dim, fpc, cpv = F.apply_tensor_list_fusion([d(a), d(b), d(c)], "avg", [40, 40, 40], [3, 3, 3], [2, 2, 2])
    assert np.array_equal(out.cpu().numpy(), np.mean(np.stack([a, b, c]), axis=0, dtype=np.float32)) and (dim, fpc, cpv) == (40, 3, 2)
    out, *_ = F.apply_tensor_list_fusion([d(a), d(b)], "maximum", [40, 40], [3, 3], [2, 2])
    assert np.array_equal(out.cpu().numpy(), np.maximum(a, b))
    # concat, equal clips per video: column concatenation
    out, dim, *_ = F.apply_tensor_list_fusion([d(a), d(b[:, :8])], "concat", [40, 8], [3, 3], [2, 2])
    assert dim == 48 and np.array_equal(out.cpu().numpy(), np.concatenate([a, b[:, :8]], axis=1))
    # concat with an auxiliary input of fewer clips per video: replicate (block tiling, as the code does), then repeat
    # every vector over the frames of its clip
    main = rng.standard_normal((2 * 2 * 3, 40)).astype(np.float32)     # 2 videos x cpv 2 x fpc 3
    aux = rng.standard_normal((2, 8)).astype(np.float32)               # 2 videos x cpv 1, vectors
    out, dim, fpc, cpv = F.apply_tensor_list_fusion([d(main), d(aux)], "concat", [40, 8], [3, 1], [2, 1])
    ref = _np_vec_seq_concat(main, _np_replicate(aux, 2, 1), 3)
    assert (dim, fpc, cpv) == (48, 3, 2) and np.array_equal(out.cpu().numpy(), ref)
    # ibias: the aux vector becomes an extra first timestep of every clip
    auxv = rng.standard_normal((4, 40)).astype(np.float32)
    out, dim, fpc, cpv = F.apply_tensor_list_fusion([d(main), d(auxv)], "ibias", [40, 40], [3, 1], [2, 2])
    ref = np.concatenate([auxv.reshape(-1, 1, 40), main.reshape(-1, 3, 40)], axis=1).reshape(-1, 40)
    assert (dim, fpc, cpv) == (40, 4, 2) and np.array_equal(out.cpu().numpy(), ref)
    with pytest.raises(Exception, match="Unknown fusion method"):
        F.apply_tensor_list_fusion([d(a), d(b)], "last", [40, 40], [3, 3], [2, 2])


@pytest.mark.parametrize("d_aux,fusion", [(40, "avg"), (64, "last")])
def test_state_biased_lstm_head_vs_oracle(d_aux, fusion):
    """model.py:128-135 + lstm.py:59-99: aux vector -> input_state_fc -> (c, h) of every layer -> dynamic_rnn over the
    clip's features -> temporal fusion -> output fc."""
    import vlb200  # noqa: F401
    from vlb200 import captioning as M, fusion as F
    rng = np.random.default_rng(67)
    clips, fpc, d, hd, layers, classes = 5, 4, 96, 64, 2, 11
    shapes = M.caption_variable_shapes(d, hd, layers, classes, d_aux, "state_bias")
    p = M.init_caption_variables(shapes, seed=5)
    feats = O.bf16_round(rng.standard_normal((clips * fpc, d)).astype(np.float32))
    aux = O.bf16_round(rng.standard_normal((clips, d_aux)).astype(np.float32))
    head = F.StateBiasedLSTMHead(p, hd, layers, fusion)
    logits = head.forward(torch.from_numpy(feats).cuda(), torch.from_numpy(aux).cuda(), fpc).cpu().numpy()
    q = O.bf16_round
    init = aux if d_aux == hd else (q(aux) @ q(p["input_state_fc_w"]) + p["input_state_fc_b"]).astype(np.float32)
    kernels = [p["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % l] for l in range(layers)]
    biases = [p["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % l] for l in range(layers)]
    out, _ = C.evaluate_sequence(feats.reshape(clips, fpc, d), kernels, biases, None, init, q=q)
    fused = O.temporal_fusion(out, fusion)
    ref = q(fused) @ q(p["output_fc_w"]) + p["output_fc_b"]
    assert logits.shape == (clips, classes)
    err = np.abs(logits - ref).max() / np.abs(ref).max()
    assert err < 1e-4, err
    assert np.array_equal(logits.argmax(1), ref.argmax(1))
