"""Profiling target: conv_flat_kernel on conv1 (space-to-depth form) and conv2 forward at 1024 frames, each twice."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import kernels as K
n, dev, bf = 1024, "cuda", torch.bfloat16
specs = [("conv1", K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID")), ("conv2", K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2))]
work = []
for name, s in specs:
    x = torch.randn(n, s.h, s.w, s.cin, device=dev).to(bf)
    wp = K.pack_conv_weight_host(s, torch.randn(s.kh, s.kw, s.cin_g, s.cout, device=dev) * 0.05)
    b = torch.full((s.cout,), 0.1, device=dev)
    y = torch.empty(n, s.p, s.q, s.cout, device=dev, dtype=bf)
    flops = 2.0 * n * s.p * s.q * s.taps * s.cin_g * s.cout
    work.append((name, flops, lambda s=s, x=x, wp=wp, b=b, y=y: K.conv_fwd_flat(s, x, wp, b, y, relu=True)))
for rep in range(2):
    for name, flops, fn in work:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if rep == 1:
            ms = e0.elapsed_time(e1)
            print("%-8s %8.1f us  %7.1f TFLOP/s" % (name, ms * 1e3, flops / ms / 1e9))
