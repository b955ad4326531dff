"""conv2 filter gradient (swapped orientation): result against the unswapped kernel, and timing."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa
from vlb200 import kernels as K
n = 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
s2 = K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2)
x2 = torch.randn(n, 28, 28, 96, device=dev).to(bf)
dy2 = torch.randn(n, 28, 28, 256, device=dev).to(bf)
ref = torch.zeros(25 * 48, 256, dtype=torch.float32, device=dev)
K.conv_wgrad(s2, x2, dy2.view(-1, 256), ref)
got = torch.zeros_like(ref)
K.conv_wgrad_t(s2, x2, dy2.view(-1, 256), got)
torch.cuda.synchronize()
print("swapped vs unswapped: max rel diff %.3e, nan %d" % (((got - ref).abs().max() / ref.abs().max()).item(), int(torch.isnan(got).sum())))
def timed(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
print("conv2 wgrad swapped: %.1f us" % timed(lambda: K.conv_wgrad_t(s2, x2, dy2.view(-1, 256), got)))
for bn in (192, 256):
    print("  block_n %d: %.1f us" % (bn, timed(lambda: K.conv_wgrad_t(s2, x2, dy2.view(-1, 256), got, block_n=bn))))
