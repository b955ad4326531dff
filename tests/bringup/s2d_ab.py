import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import vlb200
from vlb200 import _native as nv
n=1024
frames = torch.randint(0, 256, (n, 227, 227, 3), dtype=torch.uint8, device="cuda")
mean = torch.tensor([99.2, 105.3, 109.5], device="cuda")
xs = torch.empty(n, 59, 59, 48, dtype=torch.bfloat16, device="cuda")
def t(it=50):
    for _ in range(5): nv.call("vl_frames_s2d", frames, 1, mean, xs, n, 227, 227, 4, 4, 4, 59, 59)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): nv.call("vl_frames_s2d", frames, 1, mean, xs, n, 227, 227, 4, 4, 4, 59, 59)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/it*1e3
for rep in range(3):
    for w in ("0","1"):
        os.environ["VL_S2D_WIDE"]=w
        print("wide=%s: %.1f us" % (w, t()))
