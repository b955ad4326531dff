"""Split-K / tile-shape sweep of the convolution filter gradients at config-2 size (1024 frames)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv, kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)

def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3 * 1e3

for (name, h, cin, cout, k, groups) in (("conv2", 28, 96, 256, 5, 2), ("conv3", 13, 256, 384, 3, 1), ("conv4", 13, 384, 384, 3, 2),
                                        ("conv5", 13, 384, 256, 3, 2)):
    spec = K.ConvSpec(h, h, cin, cout, k, k, 1, groups)
    x = torch.randn(n, h, h, cin, device=dev).to(bf)
    dy = torch.randn(n, h, h, cout, device=dev).to(bf)
    dw = torch.zeros(k * k * (cin // groups), cout, dtype=torch.float32, device=dev)
    flops = 2.0 * n * h * h * k * k * (cin // groups) * cout
    res = []
    for (fn_name, fn) in (("wgrad", K.conv_wgrad), ("wgrad_t", K.conv_wgrad_t)):
        for sk in (0, 4, 8, 16, 32, 64):
            try:
                us = timed(lambda: fn(spec, x, dy, dw, split_k=sk))
                res.append((us, "%s split_k=%d" % (fn_name, sk)))
            except Exception as e:
                res.append((1e9, "%s split_k=%d failed %s" % (fn_name, sk, str(e)[:60])))
    res.sort()
    print("%s: " % name + "; ".join("%s %.0f us (%.0f TF/s)" % (lab, us, flops / us / 1e6) for us, lab in res[:4]), flush=True)
    auto = [r for r in res if r[1].endswith("wgrad split_k=0")][0]
    print("    auto (used by the engine): %.0f us" % auto[0], flush=True)
