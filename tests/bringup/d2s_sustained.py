"""conv2 data gradient, depth-to-space variants under SUSTAINED load (1 s loops, power-capped clocks): the 2x2 form issues
1.44x the real MACs at N = 192, the 1x2 / 2x1 forms 1.2x at N = 96."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa
from vlb200 import _native as nv, kernels as K
n = 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
spec = K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2)
dy = torch.randn(n, 28, 28, 256, device=dev).to(bf)
w = (torch.randn(5, 5, 48, 256, device=dev) * 0.05)
w2d = w.to(bf).reshape(-1, 256).contiguous()
flops = K.conv_flops(spec, n)

def sustained(fn, seconds=1.0):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); fn(); fn(); e1.record(); torch.cuda.synchronize()
    us0 = e0.elapsed_time(e1) / 3 * 1e3
    iters = max(10, int(seconds * 1e6 / us0))
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    time.sleep(0.5)
    return us0, e0.elapsed_time(e1) / iters * 1e3

dx0 = torch.empty(n, 28, 28, 96, dtype=bf, device=dev)
for rep in range(2):
    a, b = sustained(lambda: K.conv_dgrad(spec, dy, w2d, dx0))
    print("im2col dgrad (N=48)            first %7.1f us  sustained %7.1f us" % (a, b), flush=True)
    for (sh, sw, bn, ms) in ((2, 2, 192, 1), (1, 2, 96, 2), (2, 1, 96, 2), (1, 2, 96, 1)):
        rows, cols = K.d2s_filter_shape(spec, sh, sw)
        wd = torch.empty(rows, cols, dtype=bf, device=dev)
        nv.call("vl_pack_dgrad_d2s", w, wd, 5, 5, 48, 128, 2, sh, sw)
        dx1 = torch.empty(n, 28, 28, 96, dtype=bf, device=dev)
        a, b = sustained(lambda: K.conv_dgrad_d2s(spec, dy, wd, dx1, sh=sh, sw=sw, block_n=bn, msub=ms))
        print("d2s %dx%d block_n=%3d msub=%d       first %7.1f us  sustained %7.1f us  (%.0f TFLOP/s sustained)" % (sh, sw, bn, ms, a, b, flops / b / 1e6), flush=True)
