"""Profiling target for the HBM-bound encoder kernels at config-2 size (1024 frames): each kernel twice."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vlb200
from vlb200 import _native as nv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
frames = torch.randint(0, 256, (n, 227, 227, 3), dtype=torch.uint8, device=dev)
mean = torch.tensor([99.2, 105.3, 109.5], device=dev)
xs = torch.empty(n, 59, 59, 48, dtype=bf, device=dev)
res = []
def timed(name, fn, bytes_):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-22s %8.1f us  %7.1f GB/s (algorithmic %.0f MB)" % (name, ms * 1e3, bytes_ / ms / 1e6, bytes_ / 1e6), flush=True)

timed("frames_s2d", lambda: nv.call("vl_frames_s2d", frames, 1, mean, xs, n, 227, 227, 4, 4, 4, 59, 59),
      frames.numel() + xs.numel() * 2)
for (h, c) in ((57, 96), (28, 256)):
    p = (h - 3) // 2 + 1
    x = torch.relu(torch.randn(n, h, h, c, device=dev) * 30).to(bf)
    y = torch.empty(n, p, p, c, dtype=bf, device=dev)
    arg = torch.empty(n, p, p, c, dtype=torch.uint8, device=dev)
    timed("lrn_pool_fwd %dx%d" % (h, c), lambda: nv.call("vl_lrn_pool_fwd", x, y, arg, n, h, h, c, 2, 2e-5, 0.75, 1.0),
          x.numel() * 2 + y.numel() * 3)
    dy = torch.randn(n, p, p, c, device=dev).to(bf)
    dx = torch.empty_like(x)
    db = torch.zeros(c, device=dev)
    timed("pool_lrn_bwd %dx%d" % (h, c), lambda: nv.call("vl_pool_lrn_bwd", x, dy, arg, dx, db, n, h, h, c, 2, 2e-5, 0.75, 1.0),
          x.numel() * 4 + y.numel() * 3)
x = torch.relu(torch.randn(n, 13, 13, 256, device=dev)).to(bf)
y = torch.empty(n, 6, 6, 256, dtype=bf, device=dev); arg = torch.empty(n, 6, 6, 256, dtype=torch.uint8, device=dev)
timed("maxpool_fwd 13x256", lambda: nv.call("vl_maxpool_fwd", x, y, arg, n, 13, 13, 256), x.numel() * 2 + y.numel() * 3)
