"""Do an issue-bound LRN / pool gradient kernel and a persistent tensor-core kernel overlap when launched on two streams?
Measures conv2's filter gradient (swapped form) alone, the 57x57x96 LRN + pool gradient alone, and both together, for
several values of the shared-memory reserve (vl_set_smem_reserve)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import vlb200  # noqa
from vlb200 import _native as nv, kernels as K, engine as E

n = 1024
sp = E.encoder_specs(227, 227)
s2 = sp["conv2"]
LRN = (2, 2e-05, 0.75, 1.0)
p1 = torch.randn(n, 28, 28, 96, device="cuda").to(torch.bfloat16)
da2 = torch.randn(n, 28, 28, 256, device="cuda").to(torch.bfloat16)
dw = torch.zeros(s2.taps * s2.cin_g, s2.cout, device="cuda")
a1 = (torch.randn(n, 57, 57, 96, device="cuda") * 50).clamp_(min=0).to(torch.bfloat16)
dp1 = torch.randn(n, 28, 28, 96, device="cuda").to(torch.bfloat16)
arg1 = torch.randint(0, 9, (n, 28, 28, 96), device="cuda", dtype=torch.uint8)
da1 = torch.empty_like(a1)
db = torch.zeros(96, device="cuda")
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()

def wgrad():
    K.conv_wgrad_t(s2, p1, da2, dw)

def lrn():
    nv.call("vl_pool_lrn_bwd", a1, dp1, arg1, da1, db, n, 57, 57, 96, *LRN)

def timed(fa, fb, it=10):
    for _ in range(2):
        if fa:
            with torch.cuda.stream(sA): fa()
        if fb:
            with torch.cuda.stream(sB): fb()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sA.wait_event(e0); sB.wait_event(e0)
    for _ in range(it):
        if fa:
            with torch.cuda.stream(sA): fa()
        if fb:
            with torch.cuda.stream(sB): fb()
    torch.cuda.current_stream().wait_stream(sA); torch.cuda.current_stream().wait_stream(sB)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

for reserve in (0, 4096, 8192, 32768, 65536):
    nv.lib().vl_set_smem_reserve(reserve)
    a, b, ab = timed(wgrad, None), timed(None, lrn), timed(wgrad, lrn)
    print("reserve %6d B: conv2 wgrad %6.1f us, lrn1 bwd %6.1f us, both on two streams %6.1f us (sum %6.1f)" % (reserve, a, b, ab, a + b))
nv.lib().vl_set_smem_reserve(0)
