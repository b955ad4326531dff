"""GPU parity of the captioning LSTM variants against oracle/caption_numpy.py: masked sequences with an
initial state (lstm.py:102-143), their BPTT at kernel level, and the greedy feedback decode in its three visual modes
(lstm.py:145-265).  Word indices are an integer output: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import caption_numpy as C
from oracle import lrcn_numpy as O

pytestmark = pytest.mark.gpu


def rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)


@pytest.fixture(scope="module")
def cap():
    import vlb200  # noqa: F401
    from vlb200 import _native, captioning
    return dict(nv=_native, M=captioning)


def _lstm_params(M, rng, input_dim, hidden, layers, vocab, visual_dim=None, mode="state_bias"):
    shapes = M.caption_variable_shapes(input_dim, hidden, layers, vocab, visual_dim, mode)
    p = M.init_caption_variables(shapes, seed=int(rng.integers(1 << 30)))
    for k in p:  # biases away from their constant initialisation so that every term matters
        if p[k].ndim == 1:
            p[k] = (rng.standard_normal(p[k].shape) * 0.1).astype(np.float32)
    return p


def _kb(p, layers):
    return ([p["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % l] for l in range(layers)],
            [p["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % l] for l in range(layers)])


def test_masked_sequences_with_initial_state_vs_oracle(cap):
    """evaluate_sequence with nonzero_per_sequence and an initial state: outputs are zero beyond a sequence's length, the
    state is carried through; a zero-length sequence returns its initial state."""
    M = cap["M"]
    rng = np.random.default_rng(41)
    b, t_len, d, hd, layers, vocab = 6, 7, 24, 64, 2, 200
    p = _lstm_params(M, rng, d, hd, layers, vocab)
    x = rng.standard_normal((b, t_len, d)).astype(np.float32)
    lengths = np.array([7, 3, 1, 5, 0, 7], np.int32)
    init = (rng.standard_normal((b, hd)) * 0.5).astype(np.float32)
    net = M.CaptionLSTM(p, hd, layers)
    kernels, biases = _kb(p, layers)
    for lens, iv in ((lengths, init), (None, None), (lengths, None)):
        out, states = net.evaluate_sequence(x, lens, iv)
        ref_out, ref_states = C.evaluate_sequence(x, kernels, biases, lens, iv, q=O.bf16_round)
        assert rel(out.cpu().numpy(), ref_out) < 1e-4
        for (c, h), (rc, rh) in zip(states, ref_states):
            assert rel(c.cpu().numpy(), rc) < 1e-4 and rel(h.cpu().numpy(), rh) < 1e-4
        if lens is not None:
            o = out.cpu().numpy()
            for i, n in enumerate(lens):
                assert not o[i, n:].any()  # exact zeros beyond the length (dynamic_rnn)
        # against the plain fp32 semantics: the north-star bf16 tolerance
        ref32, _ = C.evaluate_sequence(x, kernels, biases, lens, iv)
        assert rel(out.cpu().numpy(), ref32) < 2e-2
    logits = net.sequence_logits(x, lengths, init)
    ref_out, _ = C.evaluate_sequence(x, kernels, biases, lengths, init, q=O.bf16_round)
    ref_logits = O.bf16_round(ref_out.reshape(b * t_len, hd)) @ O.bf16_round(p["output_fc_w"]) + p["output_fc_b"]
    assert logits.shape == (b, t_len, vocab) and rel(logits.cpu().numpy().reshape(b * t_len, vocab), ref_logits) < 1e-4


def test_masked_bptt_with_initial_and_final_state_gradients(cap):
    """vl_lstm_bwd_ex against torch autograd of the same masked recurrence: gate gradients dg (bf16 on the device),
    gradient w.r.t. the initial state; padded steps emit zero gate gradients and pass the state gradient through."""
    nv = cap["nv"]
    rng = np.random.default_rng(43)
    b, t_len, hd = 5, 6, 64
    lengths = np.array([6, 2, 4, 0, 5], np.int32)
    gx = torch.tensor(rng.standard_normal((b, t_len, 4 * hd)).astype(np.float32), requires_grad=True)
    wh = torch.tensor((rng.standard_normal((hd, 4 * hd)) * 0.2).astype(np.float32))
    h0 = torch.tensor((rng.standard_normal((b, hd)) * 0.5).astype(np.float32), requires_grad=True)
    c0 = torch.tensor((rng.standard_normal((b, hd)) * 0.5).astype(np.float32), requires_grad=True)
    g_out = torch.tensor(rng.standard_normal((b, t_len, hd)).astype(np.float32))
    g_hl = torch.tensor(rng.standard_normal((b, hd)).astype(np.float32))
    g_cl = torch.tensor(rng.standard_normal((b, hd)).astype(np.float32))
    # torch reference
    h, c = h0, c0
    outs = []
    lens_t = torch.tensor(lengths.astype(np.int64))
    for t in range(t_len):
        g = gx[:, t] + h @ wh
        i, j, f, o = g.chunk(4, dim=1)
        c_new = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
        h_new = torch.tanh(c_new) * torch.sigmoid(o)
        valid = (t < lens_t)[:, None]
        c = torch.where(valid, c_new, c)
        h = torch.where(valid, h_new, h)
        outs.append(torch.where(valid, h_new, torch.zeros_like(h_new)))
    loss = (torch.stack(outs, 1) * g_out).sum() + (h * g_hl).sum() + (c * g_cl).sum()
    loss.backward()
    # device
    d = lambda t: t.detach().cuda().contiguous()
    acts = torch.empty(b * t_len, 4 * hd, device="cuda")
    cs = torch.empty(b * t_len, hd, device="cuda")
    hseq = torch.empty(b * t_len, hd, device="cuda")
    hl, cl = torch.empty(b, hd, device="cuda"), torch.empty(b, hd, device="cuda")
    lens_d = torch.from_numpy(lengths).cuda()
    nv.call("vl_lstm_fwd_ex", d(gx).view(b * t_len, -1), d(wh), d(h0), d(c0), lens_d, acts, cs, hseq, None, None, hl, cl,
            b, t_len, hd, 1.0)
    assert rel(hl.cpu().numpy(), h.detach().numpy()) < 1e-5 and rel(cl.cpu().numpy(), c.detach().numpy()) < 1e-5
    wht = d(wh).t().contiguous()
    dg = torch.empty(b * t_len, 4 * hd, dtype=torch.bfloat16, device="cuda")
    dh0, dc0 = torch.empty(b, hd, device="cuda"), torch.empty(b, hd, device="cuda")
    nv.call("vl_lstm_bwd_ex", d(g_out).view(b * t_len, hd), d(g_hl), d(g_cl), acts, cs, d(c0), wht, lens_d, dg, dh0, dc0,
            b, t_len, hd)
    assert rel(dg.float().cpu().numpy(), gx.grad.numpy().reshape(b * t_len, -1)) < 1e-2  # bf16 storage of dg
    assert rel(dh0.cpu().numpy(), h0.grad.numpy()) < 1e-4 and rel(dc0.cpu().numpy(), c0.grad.numpy()) < 1e-4
    pad = np.arange(t_len)[None, :] >= lengths[:, None]
    assert not dg.float().cpu().numpy().reshape(b, t_len, -1)[pad].any()


@pytest.mark.parametrize("mode,dv,e", [("state_bias", 48, 32), ("state_bias", 64, 32), ("input_concat", 48, 32),
                                       ("input_bias", 32, 32)])
def test_greedy_feedback_decode_word_indices_bit_exact(cap, mode, dv, e):
    """generate_feedback_sequence: the word indices (argmax over the vocabulary, lowest index on ties, fed back through
    the embedding matrix) are an integer output and must equal the oracle's, item by item and step by step."""
    M = cap["M"]
    rng = np.random.default_rng(47)
    b, hd, layers, vocab, seq = 5, 64, 2, 1000, 9
    p = _lstm_params(M, rng, e, hd, layers, vocab, dv, mode)
    # the default initialisation decodes one constant word (the output bias wins); stronger projections make the
    # chosen word depend on the state, i.e. on every earlier step of the feedback loop
    p["output_fc_w"] = (p["output_fc_w"] * 40).astype(np.float32)
    for l in range(layers):
        k = "rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % l
        p[k] = (p[k] * 4).astype(np.float32)
    emb = O.bf16_round((rng.standard_normal((vocab, e)) * 0.7).astype(np.float32))
    start = O.bf16_round(rng.standard_normal(e).astype(np.float32))
    visual = O.bf16_round(rng.standard_normal((b, dv)).astype(np.float32))
    net = M.CaptionLSTM(p, hd, layers)
    got = net.generate_feedback_sequence(visual, start, emb, seq, mode)
    kernels, biases = _kb(p, layers)
    state_fc = (p["input_state_fc_w"], p["input_state_fc_b"]) if "input_state_fc_w" in p else None
    assert (state_fc is not None) == (mode == "state_bias" and dv != hd)
    ref = C.generate_feedback_sequence(visual, kernels, biases, p["output_fc_w"], p["output_fc_b"], start, emb, seq, mode,
                                       state_fc, q=O.bf16_round)
    steps = seq - 1 if mode == "input_bias" else seq  # the first output of input_bias is not stored (lstm.py:243-245)
    assert got.dtype == np.int64 and got.shape == ref.shape == (b * steps,)
    assert np.array_equal(got, ref)
    assert len(set(got.tolist())) >= 4  # a decode that depends on its inputs, not one constant word


def test_argmax_gather_lowest_index_wins_ties(cap):
    nv = cap["nv"]
    rng = np.random.default_rng(49)
    rows, v, e = 7, 10000, 40
    logits = rng.standard_normal((rows, v)).astype(np.float32)
    logits[0, [17, 4000, 9999]] = 9.0  # three-way tie: index 17
    logits[1, 9999] = 9.0
    logits[2, 0] = 9.0
    emb = rng.standard_normal((v, e)).astype(np.float32)
    idx = torch.empty(rows, dtype=torch.int64, device="cuda")
    out = torch.empty(rows, e, device="cuda")
    out_bf = torch.empty(rows, e, dtype=torch.bfloat16, device="cuda")
    lg = torch.from_numpy(logits).cuda()
    nv.call("vl_argmax_gather", lg, rows, v, v, torch.from_numpy(emb).cuda(), e, idx, out, out_bf)
    ref = logits.argmax(1)
    assert np.array_equal(idx.cpu().numpy(), ref) and ref[0] == 17
    assert np.array_equal(out.cpu().numpy(), emb[ref])
    assert torch.equal(out_bf.cpu(), torch.from_numpy(emb[ref]).to(torch.bfloat16))
