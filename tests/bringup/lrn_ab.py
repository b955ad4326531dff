"""A/B of the fused LRN + pool kernels (forward / backward, both AlexNet geometries) against env switches."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import vlb200  # noqa
from vlb200 import _native as nv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
LRN = (2, 2e-05, 0.75, 1.0)

def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

for (h, c) in ((57, 96), (28, 256)):
    p = (h - 3) // 2 + 1
    x = (torch.randn(n, h, h, c, device="cuda") * 60).clamp_(min=0).to(torch.bfloat16)
    dy = torch.randn(n, p, p, c, device="cuda").to(torch.bfloat16)
    y = torch.empty(n, p, p, c, device="cuda", dtype=torch.bfloat16)
    arg = torch.empty(n, p, p, c, device="cuda", dtype=torch.uint8)
    dx = torch.empty_like(x)
    db = torch.zeros(c, device="cuda")
    fwd_bytes = n * c * (h * h * 2 + p * p * 3)
    bwd_bytes = n * c * (h * h * 4 + p * p * 3)
    ref = None
    for rep in range(2):
        for mode in ("0", "1"):
            os.environ["VL_LRN_POW"] = mode
            us = t(lambda: nv.call("vl_lrn_pool_fwd", x, y, arg, n, h, h, c, *LRN))
            if ref is None:
                ref = y.float().clone()
            err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
            print("fwd %dx%dx%d VL_LRN_POW=%s : %7.1f us  %6.0f GB/s (%.2f of 6555)  max rel diff vs mode 0: %.2e" % (
                h, h, c, mode, us, fwd_bytes / us / 1e3, fwd_bytes / us / 1e3 / 6555, err))
        us = t(lambda: nv.call("vl_pool_lrn_bwd", x, dy, arg, dx, db, n, h, h, c, *LRN))
        print("bwd %dx%dx%d             : %7.1f us  %6.0f GB/s (%.2f of 6555)" % (h, h, c, us, bwd_bytes / us / 1e3, bwd_bytes / us / 1e3 / 6555))
