"""conv_flat on conv1 (s2d): stages switched off one by one / two by two (VL_GEMM_DBG bits: 1 no MMA, 2 no x loads,
8 no stores, 64 no epilogue)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import vlb200  # noqa
from vlb200 import kernels as K, engine as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
which = sys.argv[2] if len(sys.argv) > 2 else "conv1_s2d"
sp = E.encoder_specs(227, 227)
s = sp[which]

def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

x = torch.randn(n, s.h, s.w, s.cin, device="cuda").to(torch.bfloat16)
w = (torch.randn(s.cout, s.k_packed, device="cuda") * 0.05).to(torch.bfloat16)
b = torch.randn(s.cout, device="cuda")
out = torch.empty(n, s.p, s.q, s.cout, device="cuda", dtype=torch.bfloat16)
for rep in range(2):
    for dbg, label in ((0, "full"), (64, "MMA + loads"), (2, "MMA + epilogue"), (1, "loads + epilogue"), (2 | 64, "MMA only"),
                       (1 | 64, "loads only"), (1 | 2, "epilogue only"), (1 | 2 | 64, "nothing"), (8, "no global stores"),
                       (2 | 64 | 128, "MMA only, no tap shift"), (128, "full, no tap shift")):
        os.environ["VL_GEMM_DBG"] = str(dbg)
        us = t(lambda: K.conv_fwd_flat(s, x, w, b, out, relu=True))
        print("%s dbg=%3d %-18s: %7.1f us" % (which, dbg, label, us))
os.environ.pop("VL_GEMM_DBG")
