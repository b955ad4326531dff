"""conv_flat forward (conv1 s2d / conv2 / conv5 at N frames) against the number of epilogue groups."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import vlb200  # noqa
from vlb200 import kernels as K, engine as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sp = E.encoder_specs(227, 227)

def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

for name in ("conv1_s2d", "conv2", "conv5"):
    s = sp[name]
    x = torch.randn(n, s.h, s.w, s.cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(s.cout, s.k_packed, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(s.cout, device="cuda")
    out = torch.empty(n, s.p, s.q, s.cout, device="cuda", dtype=torch.bfloat16)
    ref = None
    for groups in ("2", "3", "4"):
        os.environ["VL_FLAT_GROUPS"] = groups
        us = t(lambda: K.conv_fwd_flat(s, x, w, b, out, relu=True))
        if ref is None:
            ref = out.clone()
        same = torch.equal(ref, out)
        flops = 2.0 * n * s.p * s.q * s.cout * s.taps * s.cin_g
        print("%-10s n=%d groups=%s : %7.1f us  %6.0f TFLOP/s (real MACs)  identical=%s" % (name, n, groups, us, flops / us / 1e6, same))
    os.environ.pop("VL_FLAT_GROUPS")
    us = t(lambda: K.conv_fwd_flat(s, x, w, b, out, relu=True))
    print("%-10s n=%d default   : %7.1f us  identical=%s" % (name, n, us, torch.equal(ref, out)))

print("---- input tile as several TMA boxes (conv1) ----")
s = sp["conv1_s2d"]
x = torch.randn(n, s.h, s.w, s.cin, device="cuda").to(torch.bfloat16)
w = (torch.randn(s.cout, s.k_packed, device="cuda") * 0.05).to(torch.bfloat16)
b = torch.randn(s.cout, device="cuda")
out = torch.empty(n, s.p, s.q, s.cout, device="cuda", dtype=torch.bfloat16)
K.conv_fwd_flat(s, x, w, b, out, relu=True)
ref = out.clone()
for rep in range(2):
    for split in ("1", "2", "3", "6"):
        os.environ["VL_FLAT_XSPLIT"] = split
        us = t(lambda: K.conv_fwd_flat(s, x, w, b, out, relu=True))
        print("conv1_s2d xsplit=%s : %7.1f us identical=%s" % (split, us, torch.equal(ref, out)))
os.environ.pop("VL_FLAT_XSPLIT")
for dbg, label in ((8, "no stores"), (64, "no epilogue at all"), (1, "no MMA"), (2, "no x loads")):
    os.environ["VL_GEMM_DBG"] = str(dbg)
    us = t(lambda: K.conv_fwd_flat(s, x, w, b, out, relu=True))
    print("conv1_s2d dbg=%d (%s): %7.1f us" % (dbg, label, us))
os.environ.pop("VL_GEMM_DBG")
