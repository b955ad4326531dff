"""Stage decomposition (VL_GEMM_DBG bits: 1 no MMA, 2 no A loads, 4 no B loads, 8 no stores) of the three contraction
launches furthest below the roofline at config-2 size: conv1 filter gradient (swapped, row-shift), conv2 filter gradient
(swapped) and conv2 data gradient (depth-to-space)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200  # noqa
from vlb200 import _native as nv, kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)


def timed(fn, it=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


cases = []
# conv1 filter gradient
s1 = K.ConvSpec(59, 59, 48, 96, 3, 3, 1, 1, padding="VALID")
x1 = torch.randn(n, 59, 59, 48, device=dev).to(bf)
dy1 = torch.randn(n, 57, 57, 96, device=dev).to(bf)
dw1 = torch.zeros(9 * 48, 96, dtype=torch.float32, device=dev)
cases.append(("conv1 wgrad row-shift", lambda: K.conv_wgrad_t(s1, x1, dy1.view(-1, 96), dw1, row_shift=True)))
# conv2 filter gradient
s2 = K.ConvSpec(28, 28, 96, 256, 5, 5, 1, 2)
x2 = torch.randn(n, 28, 28, 96, device=dev).to(bf)
dy2 = torch.randn(n, 28, 28, 256, device=dev).to(bf)
dw2 = torch.zeros(25 * 48, 256, dtype=torch.float32, device=dev)
cases.append(("conv2 wgrad swapped", lambda: K.conv_wgrad_t(s2, x2, dy2.view(-1, 256), dw2)))
# conv2 data gradient
w = (torch.randn(5, 5, 48, 256, device=dev) * 0.05)
rows, cols = K.d2s_filter_shape(s2, 2, 2)
wd = torch.empty(rows, cols, dtype=bf, device=dev)
nv.call("vl_pack_dgrad_d2s", w, wd, 5, 5, 48, 128, 2, 2, 2)
dx2 = torch.empty(n, 28, 28, 96, dtype=bf, device=dev)
cases.append(("conv2 dgrad d2s", lambda: K.conv_dgrad_d2s(s2, dy2, wd, dx2, sh=2, sw=2)))

for name, fn in cases:
    for dbg, label in ((0, "full"), (8, "no stores"), (1, "no MMA"), (2, "no A loads"), (4, "no B loads"), (6, "MMA + epilogue only"),
                       (1 | 8, "loads only"), (1 | 4 | 8, "A loads only"), (1 | 2 | 8, "B loads only"), (1 | 2 | 4 | 8, "hand-shakes only"),
                       (2 | 4 | 8, "MMA only")):
        os.environ["VL_GEMM_DBG"] = str(dbg)
        print("%-24s dbg=%2d %-22s %8.1f us" % (name, dbg, label, timed(fn)), flush=True)
os.environ.pop("VL_GEMM_DBG")

# one / two TMA producer warps (VL_GEMM_PRODUCERS)
for name, fn in cases:
    for prod in ("1", "2"):
        os.environ["VL_GEMM_PRODUCERS"] = prod
        print("%-24s producers=%s full %8.1f us" % (name, prod, timed(fn)), flush=True)
        os.environ["VL_GEMM_DBG"] = str(1 | 8)
        print("%-24s producers=%s loads only %8.1f us" % (name, prod, timed(fn)), flush=True)
        os.environ.pop("VL_GEMM_DBG")
os.environ.pop("VL_GEMM_PRODUCERS")
