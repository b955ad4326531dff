"""Development probe: decompose the contraction kernel's time on the train step's conv shapes into TMA-A feed,
TMA-B feed, MMA issue and epilogue, by disabling parts of the pipeline through VL_GEMM_DBG (results are garbage;
only the timings matter)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa
from vlb200 import kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


cases = []
specs = {"conv1": K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID"), "conv2": K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2), "conv3": K.ConvSpec(13, 13, 256, 384, 3, 3, 1, 1),
         "conv4": K.ConvSpec(13, 13, 384, 384, 3, 3, 1, 2), "conv5": K.ConvSpec(13, 13, 384, 256, 3, 3, 1, 2)}
for name, s in specs.items():
    x = torch.randn(n, s.h, s.w, s.cin, device=dev).to(bf)
    w = torch.randn(s.kh, s.kw, s.cin_g, s.cout, device=dev) * 0.05
    wp = K.pack_conv_weight_host(s, w)
    wd = w.to(bf).reshape(s.taps * s.cin_g, s.cout).contiguous()
    b = torch.full((s.cout,), 0.1, device=dev)
    y = torch.empty(n, s.p, s.q, s.cout, device=dev, dtype=bf)
    dy = torch.randn(n, s.p, s.q, s.cout, device=dev).to(bf)
    dx = torch.empty(n, s.h, s.w, s.cin, device=dev, dtype=bf) if name != "conv1" else None
    dw = torch.zeros(s.taps * s.cin_g, s.cout, device=dev)
    flops = 2.0 * n * s.p * s.q * s.taps * s.cin_g * s.cout
    cases.append((name + " fwd", flops, lambda s=s, x=x, wp=wp, b=b, y=y: K.conv_fwd(s, x, wp, b, y, relu=True)))
    if dx is not None:
        cases.append((name + " dgrad", flops, lambda s=s, dy=dy, wd=wd, dx=dx: K.conv_dgrad(s, dy, wd, dx)))
    if True:
        cases.append((name + " fwd FLAT", flops, lambda s=s, x=x, wp=wp, b=b, y=y: K.conv_fwd_flat(s, x, wp, b, y, relu=True)))
    if dx is not None:
        kpad = -(-s.cout_g // 64) * 64
        wdk = torch.randn(s.cin, s.taps * kpad, device=dev).to(bf)
        cases.append((name + " dgrad FLAT", flops, lambda s=s, dy=dy, wdk=wdk, dx=dx: K.conv_dgrad_flat(s, dy, wdk, dx)))
    cases.append((name + " wgrad", flops, lambda s=s, x=x, dy=dy, dw=dw: K.conv_wgrad(s, x, dy, dw)))
    cases.append((name + " wgrad T", flops, lambda s=s, x=x, dy=dy, dw=dw: K.conv_wgrad_t(s, x, dy, dw)))
# dense layers
for name, m, kk, nn in (("fc6", n, 9216, 4096), ("fc7", n, 4096, 4096)):
    x = torch.randn(m, kk, device=dev).to(bf)
    w = torch.randn(kk, nn, device=dev).to(bf)
    b = torch.zeros(nn, device=dev)
    y = torch.empty(m, nn, device=dev, dtype=bf)
    cases.append((name + " fwd", 2.0 * m * kk * nn, lambda x=x, w=w, b=b, y=y: K.linear_fwd(x, w, b, y, relu=True)))

modes = [(0, "full"), (1, "noMMA"), (1 | 4, "A only"), (1 | 2, "B only"), (1 | 2 | 4, "no loads"), (8, "no store"),
         (2 | 4, "MMA+epi only")]
if "--quick" in sys.argv:
    modes = [(0, "full"), (1, "noMMA"), (2 | 4, "MMA+epi only")]
print("%-14s" % "case" + "".join("%14s" % m[1] for m in modes) + "   TF/s(full)")
for name, flops, fn in cases:
    row = []
    for dbg, _ in modes:
        os.environ["VL_GEMM_DBG"] = str(dbg)
        row.append(timed(fn))
    os.environ["VL_GEMM_DBG"] = "0"
    print("%-16s" % name + "".join("%11.1f us" % t for t in row) + "   %8.1f" % (flops / row[0] / 1e6), flush=True)
