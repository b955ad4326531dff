"""Fused LRN + pool forward: register / shuffle kernel (fourth generation, default) against the shared-memory strip
kernel (VL_LRN_FWD_V2=1; the row-ring kernel of profiles/r02_lrn_fwd4.txt has been removed) on both AlexNet geometries: pooled values and argmax codes compared, then timed."""
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
    torch.manual_seed(1)
    x = (torch.randn(n, h, h, c, device="cuda") * 60).clamp_(min=0).to(torch.bfloat16)
    fwd_bytes = n * c * (h * h * 2 + p * p * 3)
    outs = {}
    for mode in ("v3", "v4"):  # "v3" = the reference kernel of the comparison (strip kernel)
        if mode == "v3":
            os.environ["VL_LRN_FWD_V2"] = "1"
        else:
            os.environ.pop("VL_LRN_FWD_V2", None)
        y = torch.full((n, p, p, c), float("nan"), device="cuda", dtype=torch.bfloat16)
        arg = torch.full((n, p, p, c), 255, device="cuda", dtype=torch.uint8)
        nv.call("vl_lrn_pool_fwd", x, y, arg, n, h, h, c, *LRN)
        torch.cuda.synchronize()
        outs[mode] = (y, arg)
    y3, a3 = outs["v3"]; y4, a4 = outs["v4"]
    neq = (y3.view(torch.int16) != y4.view(torch.int16))
    print("%dx%dx%d: pooled values differing: %d of %d (max rel diff %.2e), argmax codes equal: %.5f, nan in v4: %d, code max %d" % (
        h, h, c, int(neq.sum()), y3.numel(), ((y3.float() - y4.float()).abs().max() / y3.float().abs().max()).item(),
        (a3 == a4).float().mean().item(), int(torch.isnan(y4.float()).sum()), int(a4.max())), flush=True)
    y = torch.empty(n, p, p, c, device="cuda", dtype=torch.bfloat16)
    arg = torch.empty(n, p, p, c, device="cuda", dtype=torch.uint8)
    for rep in range(2):
        for mode, env in (("v2 strip kernel", {"VL_LRN_FWD_V2": "1"}), ("v4 grid 4/SM", {"VL_LRN_FWD_CTAS": "4"}),
                          ("v4 grid 8/SM", {"VL_LRN_FWD_CTAS": "8"}), ("v4 grid 16/SM", {}), ("v4 grid 32/SM", {"VL_LRN_FWD_CTAS": "32"})):
            for k in ("VL_LRN_FWD_V2", "VL_LRN_FWD_CTAS"):
                os.environ.pop(k, None)
            os.environ.update(env)
            us = t(lambda: nv.call("vl_lrn_pool_fwd", x, y, arg, n, h, h, c, *LRN))
            print("fwd %dx%dx%d %-14s: %7.1f us  %6.0f GB/s (%.3f of 6555)" % (h, h, c, mode, us, fwd_bytes / us / 1e3, fwd_bytes / us / 1e3 / 6555), flush=True)
