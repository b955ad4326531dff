"""BASELINE configs[2] / [3] shapes on the captioning operators: 2-layer LSTM(512), vocabulary 10 000, batch 256.
  teacher-forced logits of masked sequences (max caption length 20 + 1) with the fc7 vector as initial state
  greedy feedback decode of 20 words for the whole batch
(the reference never wires these into its Model / Validation; the times are for the device operators alone)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import torch
import vlb200  # noqa
from vlb200 import captioning as M

b, t_len, e, hd, layers, vocab, dv = 256, 21, 304, 512, 2, 10000, 4096
rng = np.random.default_rng(0)
p = M.init_caption_variables(M.caption_variable_shapes(e, hd, layers, vocab, dv, "state_bias"), seed=1)
net = M.CaptionLSTM(p, hd, layers)
x = torch.randn(b, t_len, e, device="cuda")
lengths = rng.integers(5, t_len + 1, b).astype(np.int32)
visual = torch.randn(b, dv, device="cuda")
emb = torch.randn(vocab, e, device="cuda") * 0.5
start = torch.randn(e, device="cuda")

def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / it * 1e3

init = net.initial_state(visual)
ms = t(lambda: net.sequence_logits(x, lengths, init))
flop = 2.0 * b * t_len * (e * 4 * hd + hd * 4 * hd + 2 * hd * 4 * hd + hd * vocab)
print("teacher-forced logits  [%d x %d] -> vocab %d : %7.3f ms  (%.0f captions/s, %.1f TFLOP/s incl. the fp32 recurrence)" % (
    b, t_len, vocab, ms, b / ms * 1e3, flop / ms / 1e9))
ms = t(lambda: net.generate_feedback_sequence(visual, start, emb, 20, "state_bias"), it=3)
print("greedy decode, 20 words, batch %d          : %7.3f ms  (%.0f captions/s, %.3f ms per word step)" % (b, ms, b / ms * 1e3, ms / 20))
