"""Profiling target: the three worst GEMM launches of the train step at config-2 size (conv1 fwd, conv2 fwd, conv2
dgrad), each launched twice (first = warm-up).  Used under `ncu --set full -k regex:umma_gemm`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = "cuda"
bf = torch.bfloat16
torch.manual_seed(0)
# conv1 as dense GEMM on the patch matrix
m1 = n * 57 * 57
col = torch.randn(m1, 384, device=dev).to(bf)
w1 = (torch.randn(384, 96, device=dev) * 0.05).to(bf)
b1 = torch.full((96,), 0.1, device=dev)
a1 = torch.empty(m1, 96, device=dev, dtype=bf)
# conv2
s2 = K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2)
p1 = torch.randn(n, 28, 28, 96, device=dev).to(bf)
w2 = torch.randn(5, 5, 48, 256, device=dev) * 0.05
w2p = K.pack_conv_weight_host(s2, w2)
w2d = w2.to(bf).reshape(25 * 48, 256).contiguous()
b2 = torch.full((256,), 0.1, device=dev)
a2 = torch.empty(n, 28, 28, 256, device=dev, dtype=bf)
da2 = torch.randn(n, 28, 28, 256, device=dev).to(bf)
dp1 = torch.empty(n, 28, 28, 96, device=dev, dtype=bf)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(2):
    ev[0].record()
    K.linear_fwd(col, w1, b1, a1, relu=True)
    ev[1].record()
    K.conv_fwd(s2, p1, w2p, b2, a2, relu=True)
    ev[2].record()
    K.conv_dgrad(s2, da2, w2d, dp1)
    ev[3].record()
torch.cuda.synchronize()
print("conv1 fwd %.1f us, conv2 fwd %.1f us, conv2 dgrad %.1f us" % (
    ev[0].elapsed_time(ev[1]) * 1e3, ev[1].elapsed_time(ev[2]) * 1e3, ev[2].elapsed_time(ev[3]) * 1e3))
