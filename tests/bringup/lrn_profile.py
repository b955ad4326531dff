"""Profiling target: the fused LRN + pool forward / backward kernels on both AlexNet geometries, each launched twice
(first = warm-up).  `ncu --set full -k regex:"lrn_pool_fwd_kernel4|pool_lrn_bwd_kernel4"`."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import vlb200  # noqa
from vlb200 import _native as nv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
LRN = (2, 2e-05, 0.75, 1.0)
for (h, c) in ((57, 96), (28, 256)):
    p = (h - 3) // 2 + 1
    x = (torch.randn(n, h, h, c, device="cuda") * 60).clamp_(min=0).to(torch.bfloat16)
    dy = torch.randn(n, p, p, c, device="cuda").to(torch.bfloat16)
    y = torch.empty(n, p, p, c, device="cuda", dtype=torch.bfloat16)
    arg = torch.empty(n, p, p, c, device="cuda", dtype=torch.uint8)
    dx = torch.empty_like(x)
    db = torch.zeros(c, device="cuda")
    for rep in range(2):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        nv.call("vl_lrn_pool_fwd", x, y, arg, n, h, h, c, *LRN)
        e1.record()
        nv.call("vl_pool_lrn_bwd", x, dy, arg, dx, db, n, h, h, c, *LRN)
        e2.record(); torch.cuda.synchronize()
        if rep:
            print("%dx%dx%d n=%d: fwd %.1f us, bwd %.1f us" % (h, h, c, n, e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3))
