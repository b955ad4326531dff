"""conv2 data gradient at config-2 size (1024 frames): im2col kernel (N = 48 per group) against the depth-to-space
formulation (stride-(sh,sw) forward convolution over dy, N = sh*sw*48)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv, kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
spec = K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2)
dy = torch.randn(n, 28, 28, 256, device=dev).to(bf)
w = (torch.randn(5, 5, 48, 256, device=dev) * 0.05)
w2d = w.to(bf).reshape(-1, 256).contiguous()
flops = 2.0 * n * 28 * 28 * 25 * 48 * 256

def timed(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-34s %8.1f us  %7.1f TFLOP/s (layer FLOPs)" % (name, ms * 1e3, flops / ms / 1e9), flush=True)

dx0 = torch.empty(n, 28, 28, 96, dtype=bf, device=dev)
timed("im2col dgrad (N=48)", lambda: K.conv_dgrad(spec, dy, w2d, dx0))
for (sh, sw, bn, ms) in ((2, 2, 192, 1), (1, 2, 96, 2), (1, 2, 96, 1), (2, 1, 96, 2), (2, 2, 96, 2), (2, 2, 64, 2)):
    rows, cols = K.d2s_filter_shape(spec, sh, sw)
    wd = torch.empty(rows, cols, dtype=bf, device=dev)
    nv.call("vl_pack_dgrad_d2s", w, wd, 5, 5, 48, 128, 2, sh, sw)
    dx1 = torch.full((n, 28, 28, 96), float("nan"), dtype=bf, device=dev)
    try:
        timed("d2s %dx%d block_n=%d msub=%d" % (sh, sw, bn, ms), lambda: K.conv_dgrad_d2s(spec, dy, wd, dx1, sh=sh, sw=sw, block_n=bn, msub=ms))
        err = (dx1.float() - dx0.float()).abs().max().item() / dx0.float().abs().max().item()
        print("    max rel diff vs im2col dgrad: %.3e" % err)
    except Exception as e:
        print("    failed:", e)
