"""Fused LRN / pool backward: L2 prefetch of the CTA's next unit (VL_LRN_BWD_PF=1, default) against none, same box."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import vlb200  # noqa
from vlb200 import _native as nv
n = 1024
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
    nv.call("vl_lrn_pool_fwd", x, y, arg, n, h, h, c, *LRN)
    dx = torch.empty_like(x); db = torch.zeros(c, device="cuda")
    ref = None
    for rep in range(2):
        for pf in ("0", "1", "2", "3"):
            os.environ["VL_LRN_BWD_PF"] = pf
            us = t(lambda: nv.call("vl_pool_lrn_bwd", x, dy, arg, dx, db, n, h, h, c, *LRN))
            if ref is None: ref = dx.clone()
            print("bwd %dx%dx%d prefetch=%s: %7.1f us  equal to first: %s" % (h, h, c, pf, us, bool(torch.equal(dx, ref))), flush=True)
