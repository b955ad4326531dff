"""Development probe: tap-shifted conv kernel, cost of misaligned (non multiple-of-8-row) operand starts."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import kernels as K
n, dev, bf = 1024, "cuda", torch.bfloat16
def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
specs = {"conv1": K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID"), "conv2": K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2)}
modes = [(0, "full"), (1, "noMMA"), (6, "MMA+epi"), (6 | 8, "MMA only"), (6 | 8 | 16, "MMA no col shift"), (6 | 8 | 32, "MMA no shift"), (1 | 8, "loads only"), (6 | 64, "MMA, no epilogue"), (1 | 64, "loads, no epi"), (1 | 6 | 64, "handshake only")]
print("%-8s" % "case" + "".join("%19s" % m[1] for m in modes))
for name, s in specs.items():
    x = torch.randn(n, s.h, s.w, s.cin, device=dev).to(bf)
    wp = K.pack_conv_weight_host(s, torch.randn(s.kh, s.kw, s.cin_g, s.cout, device=dev) * 0.05)
    b = torch.full((s.cout,), 0.1, device=dev)
    y = torch.empty(n, s.p, s.q, s.cout, device=dev, dtype=bf)
    row = []
    for dbg, _ in modes:
        os.environ["VL_GEMM_DBG"] = str(dbg)
        row.append(timed(lambda: K.conv_fwd_flat(s, x, wp, b, y, relu=True)))
    os.environ["VL_GEMM_DBG"] = "0"
    print("%-8s" % name + "".join("%16.1f us" % t for t in row), flush=True)
