"""Profiling target: the heaviest contraction launches of the train step at config-2 size (conv1 fwd, conv2 fwd,
conv2 dgrad, conv2 wgrad, conv3 fwd), each launched twice (first = warm-up).  `ncu --set full -k regex:umma_gemm`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
specs = [("conv1", K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID")), ("conv2", K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2)),
         ("conv3", K.ConvSpec(13, 13, 256, 384, 3, 3, 1, 1))]
work = []
for name, s in specs:
    x = torch.randn(n, s.h, s.w, s.cin, device=dev).to(bf)
    w = torch.randn(s.kh, s.kw, s.cin_g, s.cout, device=dev) * 0.05
    wp = K.pack_conv_weight_host(s, w)
    wd = w.to(bf).reshape(s.taps * s.cin_g, s.cout).contiguous()
    b = torch.full((s.cout,), 0.1, device=dev)
    y = torch.empty(n, s.p, s.q, s.cout, device=dev, dtype=bf)
    dy = torch.randn(n, s.p, s.q, s.cout, device=dev).to(bf)
    dw = torch.zeros(s.taps * s.cin_g, s.cout, device=dev)
    flops = 2.0 * n * s.p * s.q * s.taps * s.cin_g * s.cout
    work.append((name + " fwd", flops, lambda s=s, x=x, wp=wp, b=b, y=y: K.conv_fwd(s, x, wp, b, y, relu=True)))
    if name == "conv2":
        dx = torch.empty(n, s.h, s.w, s.cin, device=dev, dtype=bf)
        work.append((name + " dgrad", flops, lambda s=s, dy=dy, wd=wd, dx=dx: K.conv_dgrad(s, dy, wd, dx)))
        work.append((name + " wgrad", flops, lambda s=s, x=x, dy=dy, dw=dw: K.conv_wgrad(s, x, dy, dw)))
for rep in range(2):
    for name, flops, fn in work:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if rep == 1:
            ms = e0.elapsed_time(e1)
            print("%-12s %8.1f us  %7.1f TFLOP/s" % (name, ms * 1e3, flops / ms / 1e9))
