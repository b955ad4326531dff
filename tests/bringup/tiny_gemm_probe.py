"""Fixed cost of one contraction launch: back-to-back launches of tiny products (1 tile) with stages switched off."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa
from vlb200 import _native as nv, kernels as K

dev, bf = "cuda", torch.bfloat16
def timed(fn, it=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

z = torch.zeros(1024, device=dev)
print("vl_zero 4 KB (plain tiny kernel)      : %6.2f us per launch" % timed(lambda: nv.call("vl_zero", z, 4096)))
for (m, n, k) in ((64, 101, 256), (64, 256, 104), (128, 256, 64), (128, 256, 1024), (128, 256, 4096), (1024, 1024, 4096)):
    x = torch.randn(m, k, device=dev).to(bf)
    w = (torch.randn(k, max(n, 104), device=dev) * 0.05).to(bf)
    b = torch.zeros(max(n, 104), device=dev)
    out = torch.empty(m, n, device=dev)
    for dbg, label in ((0, "full"), (1, "no MMA"), (1 | 2 | 4, "hand-shakes only"), (1 | 2 | 4 | 8, "hand-shakes, no stores")):
        os.environ["VL_GEMM_DBG"] = str(dbg)
        print("dense fwd %4dx%4dx%4d %-22s: %6.2f us per launch" % (m, n, k, label, timed(lambda: K.linear_fwd(x, w, b, out, n=n))), flush=True)
os.environ.pop("VL_GEMM_DBG")
