"""LSTM input projection (m = frames, n = 4H = 1024, k = 4096) against the tile shape: block_n x msub."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import vlb200  # noqa
from vlb200 import kernels as K

def t(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

for (m, n, k) in ((1024, 1024, 4096), (1024, 4096, 4096), (1024, 4096, 9216), (256, 1024, 4096)):
    x = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(k, n, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda")
    for bn, ms in ((0, 0), (256, 1), (128, 1), (128, 2), (64, 1), (64, 2), (192, 1)):
        try:
            us = t(lambda: K.linear_fwd(x, w, b, out, relu=False, block_n=bn, msub=ms))
            print("m=%d n=%d k=%d  block_n=%3d msub=%d : %7.1f us  %6.0f TFLOP/s" % (m, n, k, bn, ms, us, 2.0 * m * n * k / us / 1e6))
        except Exception as ex:
            print("m=%d n=%d k=%d  block_n=%3d msub=%d : %s" % (m, n, k, bn, ms, str(ex)[:80]))

print("---- split-K (fp32 red.add into the zeroed output) ----")
m, n, k = 1024, 1024, 4096
x = torch.randn(m, k, device="cuda").to(torch.bfloat16)
w = (torch.randn(k, n, device="cuda") * 0.05).to(torch.bfloat16)
b = torch.randn(n, device="cuda")
out = torch.empty(m, n, device="cuda")
K.linear_fwd(x, w, b, out, relu=False)
ref = out.clone()
for bn, ms, sk in ((256, 1, 2), (256, 1, 4), (256, 1, 8), (128, 1, 2), (128, 1, 4), (128, 2, 4), (128, 2, 8), (64, 1, 2), (64, 1, 4)):
    us = t(lambda: K.linear_fwd(x, w, b, out, relu=False, block_n=bn, msub=ms, split_k=sk))
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    print("m=%d n=%d k=%d block_n=%3d msub=%d split_k=%d : %7.1f us (incl. zero fill)  rel diff %.1e" % (m, n, k, bn, ms, sk, us, err))
