"""Profiling target: umma_gemm_kernel on the backward / wide-layer launches of the train step at 1024 frames (conv1 filter
gradient, conv2 data / filter gradient, conv3 forward / data / filter gradient, fc6 forward / filter gradient), each
launched twice (first = warm-up).  `ncu --set full -k regex:umma_gemm`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa
from vlb200 import _native as nv, kernels as K
n, dev, bf = 1024, "cuda", torch.bfloat16
torch.manual_seed(0)
work = []
s1s = K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID")
x1 = torch.randn(n, 59, 59, 48, device=dev).to(bf); dy1 = torch.randn(n, 57, 57, 96, device=dev).to(bf)
dw1 = torch.zeros(9 * 48, 96, dtype=torch.float32, device=dev)
f1 = 2.0 * n * 57 * 57 * 96 * 363
work.append(("conv1 wgrad (row-shift)", f1, lambda: K.conv_wgrad_t(s1s, x1, dy1.view(-1, 96), dw1, row_shift=True, flops=f1)))
s2 = K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2)
x2 = torch.randn(n, 28, 28, 96, device=dev).to(bf); dy2 = torch.randn(n, 28, 28, 256, device=dev).to(bf)
dw2 = torch.zeros(25 * 48, 256, dtype=torch.float32, device=dev)
w = (torch.randn(5, 5, 48, 256, device=dev) * 0.05)
rows, cols = K.d2s_filter_shape(s2, 2, 2)
wd = torch.empty(rows, cols, dtype=bf, device=dev)
nv.call("vl_pack_dgrad_d2s", w, wd, 5, 5, 48, 128, 2, 2, 2)
dx2 = torch.empty(n, 28, 28, 96, dtype=bf, device=dev)
f2 = K.conv_flops(s2, n)
work.append(("conv2 dgrad (d2s)", f2, lambda: K.conv_dgrad_d2s(s2, dy2, wd, dx2, sh=2, sw=2)))
work.append(("conv2 wgrad (swapped)", f2, lambda: K.conv_wgrad_t(s2, x2, dy2.view(-1, 256), dw2)))
s3 = K.ConvSpec(13, 13, 256, 384, 3, 3, 1, 1)
x3 = torch.randn(n, 13, 13, 256, device=dev).to(bf)
w3k = K.pack_conv_weight_host(s3, torch.randn(3, 3, 256, 384, device=dev) * 0.05)
w3d = (torch.randn(9 * 256, 384, device=dev) * 0.05).to(bf)
b384 = torch.zeros(384, device=dev)
a3 = torch.empty(n, 13, 13, 384, device=dev, dtype=bf); dy3 = torch.randn(n, 13, 13, 384, device=dev).to(bf)
dx3 = torch.empty(n, 13, 13, 256, device=dev, dtype=bf); dw3 = torch.zeros(9 * 256, 384, dtype=torch.float32, device=dev)
f3 = K.conv_flops(s3, n)
work.append(("conv3 fwd", f3, lambda: K.conv_fwd(s3, x3, w3k, b384, a3)))
work.append(("conv3 dgrad", f3, lambda: K.conv_dgrad(s3, dy3, w3d, dx3, relu_mask=x3)))
work.append(("conv3 wgrad", f3, lambda: K.conv_wgrad(s3, x3, dy3, dw3)))
xf = torch.randn(n, 9216, device=dev).to(bf); wf = (torch.randn(9216, 4096, device=dev) * 0.02).to(bf)
b4096 = torch.zeros(4096, device=dev); of = torch.empty(n, 4096, dtype=bf, device=dev)
dyf = torch.randn(n, 4096, device=dev).to(bf); dwf = torch.zeros(9216, 4096, dtype=torch.float32, device=dev)
ff = 2.0 * n * 9216 * 4096
work.append(("fc6 fwd", ff, lambda: K.linear_fwd(xf, wf, b4096, of, relu=True)))
work.append(("fc6 wgrad", ff, lambda: K.linear_wgrad(xf, dyf, dwf, split_k=1)))
for rep in range(2):
    for name, flops, fn in work:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if rep == 1:
            ms = e0.elapsed_time(e1)
            print("%-26s %8.1f us  %7.1f TFLOP/s" % (name, ms * 1e3, flops / ms / 1e9))
