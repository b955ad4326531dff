"""Tile-shape probe for the feed-bound contractions (L2 -> SM ingest ~60 B/clk/SM): conv3 / conv4 forward and data
gradient with 128 x 192 tiles (msub 1) against 256 x 128 / 256 x 96 tiles (msub 2)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv, kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)

def timed(name, fn, flops):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-40s %8.1f us  %7.1f TFLOP/s" % (name, ms * 1e3, flops / ms / 1e9), flush=True)

for (name, cin, cout, groups) in (("conv3", 256, 384, 1), ("conv4", 384, 384, 2)):
    spec = K.ConvSpec(13, 13, cin, cout, 3, 3, 1, groups)
    x = torch.randn(n, 13, 13, cin, device=dev).to(bf)
    dy = torch.randn(n, 13, 13, cout, device=dev).to(bf)
    w = torch.randn(3, 3, cin // groups, cout, device=dev) * 0.05
    wk = K.pack_conv_weight_host(spec, w)
    w2d = w.to(bf).reshape(-1, cout).contiguous()
    b = torch.zeros(cout, device=dev)
    out = torch.empty(n, 13, 13, cout, dtype=bf, device=dev)
    dx = torch.empty(n, 13, 13, cin, dtype=bf, device=dev)
    flops = 2.0 * n * 169 * 9 * (cin // groups) * cout
    for (bn, ms) in ((0, 0), (128, 2), (96, 2), (192, 1), (256, 1), (64, 2)):
        if spec.cout_g % bn if bn else False:
            continue
        try:
            timed("%s fwd block_n=%d msub=%d" % (name, bn, ms), lambda: K.conv_fwd(spec, x, wk, b, out, block_n=bn, msub=ms), flops)
        except Exception as e:
            print("   fwd bn=%d msub=%d failed: %s" % (bn, ms, str(e)[:80]))
    for (bn, ms) in ((0, 0), (128, 2), (96, 2), (192, 1), (64, 2)):
        if spec.cin_g % bn if bn else False:
            continue
        try:
            timed("%s dgrad block_n=%d msub=%d" % (name, bn, ms), lambda: K.conv_dgrad(spec, dy, w2d, dx, relu_mask=x, block_n=bn, msub=ms), flops)
        except Exception as e:
            print("   dgrad bn=%d msub=%d failed: %s" % (bn, ms, str(e)[:80]))
