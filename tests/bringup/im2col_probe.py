"""Development probe: does the im2col-mode TMA feed depend on channel padding / pixel pitch?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import kernels as K
n = 1024
dev, bf = "cuda", torch.bfloat16
def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
if "--lie" in sys.argv:
    os.environ["VL_IM2COL_LIE"] = "1"
specs = {"conv1 48ch": K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID"),
         "conv1 64ch": K.ConvSpec(59, 59, 64, 96, 3, 3, 1, 1, padding="VALID"),
         "conv2 96ch g2": K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2),
         "conv2 128ch g2": K.ConvSpec(28, 28, 128, 256, 5, 5, 1, 2),
         "conv2 64ch g1": K.ConvSpec(28, 28, 64, 128, 5, 5, 1, 1),
         "conv3": K.ConvSpec(13, 13, 256, 384, 3, 3, 1, 1),
         "conv5": K.ConvSpec(13, 13, 384, 256, 3, 3, 1, 2)}
modes = [(0, "full"), (1, "noMMA"), (1 | 4, "A only"), (1 | 2, "B only"), (1 | 2 | 4, "no loads"), (2 | 4, "MMA only")]
print("%-16s" % "case" + "".join("%12s" % m[1] for m in modes))
for name, s in specs.items():
    x = torch.randn(n, s.h, s.w, s.cin, device=dev).to(bf)
    w = torch.randn(s.kh, s.kw, s.cin_g, s.cout, device=dev) * 0.05
    wp = K.pack_conv_weight_host(s, w)
    b = torch.full((s.cout,), 0.1, device=dev)
    y = torch.empty(n, s.p, s.q, s.cout, device=dev, dtype=bf)
    for msub in (1, 2):
        if msub == 2 and s.cout_g > 128: continue
        row = []
        for dbg, _ in modes:
            os.environ["VL_GEMM_DBG"] = str(dbg)
            row.append(timed(lambda: K.conv_fwd(s, x, wp, b, y, relu=True, msub=msub)))
        os.environ["VL_GEMM_DBG"] = "0"
        print("%-16s" % (name + " m%d" % msub) + "".join("%9.1f us" % t for t in row), flush=True)
